"""Generates tests/golden/dropin_cases.json.gz by running the UNMODIFIED reference (/root/reference/fslr/cluster.py
through oracle/ref_harness.run_reference_block, the restated main.py:209-257,334-352) in the build container:

    python tests/golden/make_dropin_golden.py

Unlike small_cases.npz (per-read ids), these fixtures keep what the reference's own BOUNDARY returns, so that the
drop-in functions of fslr_b200.cluster can be driven exactly as main.py:227-244 drives cluster.py and compared call by
call: the `list[set[str]]` of cluster.get_subgraphs, `network.number_of_nodes()`, the permutation of the reference's
unstable sort, and the bytes of `<base>.mappings.cluster.bed` / `<base>.mappings.representative.bed`.  Cases with
`filter_false=True` cover main.py:229-230 / cluster.py:80-86.
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                          # noqa: E402
from tests.golden import make_golden as mg                    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def with_false_names(df, rng, frac=0.3):
    """Some reads lose a primer: their qname carries `False` (find_reads_with_primers.py:61-62,109)."""
    names = df["qname"].unique()
    pick = set(n for n in names if rng.random() < frac)
    df = df.copy()
    df["qname"] = [q + ".0.9_0.9.21q1F_False" if q in pick else q + ".0.9_0.9.21q1F_21q1R" for q in df["qname"]]
    return df.sort_values(["n_alignments", "qname", "qstart"], ascending=[False, True, True]).reset_index(drop=True)


def all_clustered():
    """Every read ends in a cluster: the left merge of main.py:334 leaves no NaN, so `cluster` / `n_reads` stay integers."""
    rows = []
    for k in range(3):
        rows += mg.read_rows("fam1_%d" % k, [("chr1", 5_000_000 + k, 5_001_000 + k), ("chr2", 7_000_000 + k, 7_000_500 + k)])
        rows += mg.read_rows("fam2_%d" % k, [("chr3", 9_000_000 + k, 9_000_700 + k)])
    return mg.frame(rows), {"chr%d" % i: 100_000_000 for i in (1, 2, 3, 9, 10)}


def plan():
    out = []
    df, lens = mg.f1_adversarial24()
    out += [("F1_adversarial24_T10", df, lens, dict(edge_threshold=10)), ("F1_adversarial24_T3", df, lens, dict(edge_threshold=3))]
    df, lens = mg.f2_greedy()
    out.append(("F2_greedy", df, lens, dict(cluster_mask="")))
    df, lens = mg.f3_ties()
    out.append(("F3_ties", df, lens, dict()))
    df, lens = mg.f4_mask()
    for m in ("subtelomere", "subtelomere,L1_TALEN", "L1_TALEN,chr6,notachrom"):
        out.append(("F4_mask[%s]" % m, df, lens, dict(cluster_mask=m)))
    df, lens = mg.f5_float_boundaries()
    out.append(("F5_float", df, lens, dict()))
    df, lens = mg.f6_degenerate_noclusters()
    out.append(("F6_noclusters", df, lens, dict()))
    df, lens = mg.f6_false_names()
    out.append(("F6_false_kept", df, lens, dict()))
    out.append(("F6_false_filtered", df, lens, dict(filter_false=True)))
    df, lens = all_clustered()
    out.append(("all_clustered_int_columns", df, lens, dict()))
    rng = np.random.default_rng(20261020)
    cut_lists = ["1,1,0.66,0.66,0.66,0.5", "0.5", "1,0.5,0.34"]
    for i in range(24):
        df, lens = mg.random_table(rng)
        opts = dict(edge_threshold=int(rng.choice([1, 2, 3, 10])), jaccard_cutoffs=cut_lists[i % 3],
                    overlap=float(rng.choice([0.8, 0.5, 0.95])), cluster_mask="subtelomere" if i % 4 else "chrX",
                    qlen_diff=float(rng.choice([0.04, 0.2, 0.0])), n_alignment_diff=float(rng.choice([0.25, 0.0, 0.5])))
        if i % 2:
            df = with_false_names(df, rng)
            opts["filter_false"] = bool(i % 4 == 1)
        out.append(("rand%02d" % i, df, lens, opts))
    rng = np.random.default_rng(20261021)
    for i in range(8):
        df, lens = mg.wide_table(rng)
        opts = dict(edge_threshold=int(rng.choice([2, 10])), overlap=float(rng.choice([0.8, 0.5])))
        if i % 2:
            df = with_false_names(df, rng, 0.2)
            opts["filter_false"] = True
        out.append(("wide%02d" % i, df, lens, opts))
    return out


def main():
    cases = []
    for name, df, lens, opts in plan():
        exp = rh.run_reference_block(df, lens, **opts)
        exp["order"] = [int(x) for x in exp["order"]]
        cases.append({"name": name, "opts": opts, "chr_lengths": lens, "columns": list(df.columns),
                      "rows": df.values.tolist(), "expected": exp})
        print(name, len(df), "rows", len(exp["subgraphs"]), "components", flush=True)
    path = os.path.join(HERE, "dropin_cases.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(cases).encode())
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

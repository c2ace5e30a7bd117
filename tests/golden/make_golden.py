"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/fslr/cluster.py,
through oracle/ref_harness.py) in the build container.  Run from the repo root:

    python tests/golden/make_golden.py [--only small|wide|configs]

Each case stores its input columns, the option strings, the permutation the reference's own
unstable sort produced on the generating host (cluster.py:114 — host dependent, so it is part of
the fixture), and the reference's per-read `cluster` / `n_reads` (main.py:334-342).  The hand-built
cases are SURVEY.md §8c F1-F6.
"""
import argparse
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fslr_b200 import synth                                   # noqa: E402
from fslr_b200.table import ColumnarTable                     # noqa: E402
from oracle import oracle as orc, ref_harness as rh           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
COLS = ["chrom", "rstart", "rend", "qname", "n_alignments", "aln_size", "qstart", "qend"]


def frame(rows, names=None):
    """rows: list of (chrom str, rstart, rend, qname str, n_alignments, aln_size, qstart, qend)."""
    df = pd.DataFrame(rows, columns=COLS)
    df["alignment_score"] = df["aln_size"] * 2
    return df


def read_rows(qname, fillings, n_aln=None, bread=("chr9", 5_000_000, 5_000_300, "chr10", 7_000_000, 7_000_300)):
    """One read: bread, fillings [(chrom, start, end[, aln_size])], bread; qstart/qend cumulative."""
    n = len(fillings) + 2 if n_aln is None else n_aln
    rows, q = [], 0
    segs = [(bread[0], bread[1], bread[2], bread[2] - bread[1])]
    for f in fillings:
        segs.append((f[0], f[1], f[2], f[3] if len(f) > 3 else f[2] - f[1]))
    segs.append((bread[3], bread[4], bread[5], bread[5] - bread[4]))
    for c, s, e, a in segs:
        rows.append((c, s, e, qname, n, a, q, q + a))
        q += a + 1
    return rows


def f1_adversarial24():
    """the edge_threshold break changes the components."""
    rows = []
    def rd(name, i, j):
        return read_rows(name, [("chr1", 1_000_000 + i, 1_001_000 + i), ("chr2", 2_000_000 + j, 2_001_000 + j)])
    rows += rd("X", 0, 0) + rd("Y", 100, 200)
    for k in range(11):
        rows += rd("P%02d" % k, 200, -200)
    for k in range(10):
        rows += rd("Q%02d" % k, 150, 400)
    rows += read_rows("D", [("chr1", 4_000_000, 4_001_000), ("chr2", 2_000_200, 2_001_199, 999)])
    lens = {"chr%d" % i: 100_000_000 for i in (1, 2, 9, 10)}
    return frame(rows), lens


def f2_greedy():
    """greedy first-fit asymmetry (cluster.py:152-161)."""
    rows = read_rows("A", [("chr1", 1000, 2000), ("chr1", 1150, 2150)])
    rows += read_rows("B", [("chr1", 1100, 2100), ("chr1", 900, 1900)])
    rows += read_rows("C", [("chr1", 900, 1900), ("chr1", 1100, 2100)])
    rows += read_rows("E", [("chr1", 1149, 2149), ("chr1", 1001, 2001)])
    return frame(rows), {"chr1": 900_000, "chr9": 100_000_000, "chr10": 100_000_000}


def f3_ties(n=400, structures=25, seed=7):
    """every read shares an identical first filling (massive start ties)."""
    rng = np.random.default_rng(seed)
    rows = []
    pos = rng.integers(2_000_000, 50_000_000, size=structures)
    for r in range(n):
        s = r % structures
        rows += read_rows("T%04d" % rng.integers(0, 10**6) + "_%d" % r,
                          [("chr3", 3_000_000, 3_001_000), ("chr4", int(pos[s]), int(pos[s]) + 900)])
    return frame(rows), {"chr3": 100_000_000, "chr4": 100_000_000, "chr9": 100_000_000, "chr10": 100_000_000}


def f4_mask():
    """subtelomere rule, <=1 Mb contig exemption, chromosome-name mask."""
    rows = []
    for k in range(3):
        rows += read_rows("near%d" % k, [("chr1", 100 + k, 200 + k), ("chr5", 30_000_000 + k, 30_000_800 + k)])
        rows += read_rows("contig%d" % k, [("ctgA", 100 + k, 200 + k), ("chr5", 31_000_000 + k, 31_000_800 + k)])
        rows += read_rows("far%d" % k, [("chr6", 600_000 + k, 600_100 + k), ("chr5", 32_000_000 + k, 32_000_800 + k)])
        rows += read_rows("onlymasked%d" % k, [("chr1", 300 + k, 900 + k)])
        rows += read_rows("l1_%d" % k, [("L1_TALEN", 500 + k, 1500 + k), ("chr5", 33_000_000 + k, 33_000_800 + k)])
        rows += read_rows("l1only_%d" % k, [("L1_TALEN", 2500 + k, 3500 + k)])
    lens = {"chr1": 50_000_000, "ctgA": 900_000, "chr6": 1_000_050, "chr5": 100_000_000,
            "chr9": 100_000_000, "chr10": 100_000_000, "L1_TALEN": 8000}
    return frame(rows), lens


def f5_float_boundaries():
    """800/1000 vs 799/999 at p=0.8; 2/3, 3/5 at 0.66; 24/25 at 1-0.04; 3/4 at 1-0.25."""
    rows = []
    # reciprocal overlap exactly 0.8 (passes) and 799/999 (fails)
    rows += read_rows("o1", [("chr1", 10_000, 11_000, 1000)]) + read_rows("o2", [("chr1", 10_200, 11_000, 1000)])
    rows += read_rows("o3", [("chr2", 10_000, 10_999, 999)]) + read_rows("o4", [("chr2", 10_200, 10_999, 999)])
    # jaccard 2/3 >= 0.66 (edge) : 2 matched of (2,3) fillings ; 3/5 < 0.66 : 3 matched of (4,4)
    rows += read_rows("j1", [("chr3", 1_000_000, 1_001_000), ("chr3", 2_000_000, 2_001_000)], n_aln=5)
    rows += read_rows("j2", [("chr3", 1_000_001, 1_001_001), ("chr3", 2_000_001, 2_001_001), ("chr4", 900_000, 901_000)], n_aln=5)
    rows += read_rows("j3", [("chr5", 1_000_000 + 2000 * k, 1_001_000 + 2000 * k) for k in range(3)] + [("chr6", 5_000_000, 5_001_000)], n_aln=6)
    rows += read_rows("j4", [("chr5", 1_000_000 + 2000 * k, 1_001_000 + 2000 * k) for k in range(3)] + [("chr7", 5_000_000, 5_001_000)], n_aln=6)
    # qlen ratio 24/25 = 0.96 with n_alignments ratio failing (3 vs 5 -> 0.6)
    rows += read_rows("q1", [("chr8", 3_000_000, 3_002_400, 2400)], n_aln=3)
    rows += read_rows("q2", [("chr8", 3_000_000, 3_002_500, 2500)], n_aln=5)
    rows += read_rows("q3", [("chr8", 4_000_000, 4_002_399, 2399)], n_aln=3)
    rows += read_rows("q4", [("chr8", 4_000_000, 4_002_500, 2500)], n_aln=5)
    lens = {"chr%d" % i: 100_000_000 for i in range(1, 11)}
    return frame(rows), lens


def f6_degenerate_noclusters():
    rows = []
    for k in range(5):
        rows += read_rows("two%d" % k, [])                        # 2 alignments -> vanish in keep_fillings
    rows += read_rows("solo", [("chr1", 5_000_000, 5_001_000)])
    return frame(rows), {"chr1": 100_000_000, "chr9": 100_000_000, "chr10": 100_000_000}


def f6_false_names():
    rows = []
    for k in range(4):
        rows += read_rows("r%d.0.9_0.9.21q1F_False" % k, [("chr1", 5_000_000 + k, 5_001_000 + k)])
        rows += read_rows("s%d.0.9_0.9.21q1F_21q1R" % k, [("chr1", 5_000_000 + k, 5_001_000 + k)])
    return frame(rows), {"chr1": 100_000_000, "chr9": 100_000_000, "chr10": 100_000_000}


def random_table(rng):
    """Dense random geometry: few chromosomes, small coordinate range, duplicates with jitter, start ties."""
    n_reads = int(rng.integers(5, 46))
    n_struct = int(rng.integers(1, 6))
    chroms = ["chr1", "chr2"] if rng.random() < 0.7 else ["chr1", "chr2", "chrX"]
    structs = []
    for _ in range(n_struct):
        L = int(rng.integers(1, 4))
        structs.append([(chroms[int(rng.integers(0, len(chroms)))], int(rng.integers(600_000, 600_400)) if rng.random() < 0.8
                         else int(rng.integers(600_000, 640_000)), int(rng.integers(80, 400))) for _ in range(L)])
    rows = []
    jit = int(rng.integers(0, 30))
    for r in range(n_reads):
        st = structs[int(rng.integers(0, n_struct))]
        fl = []
        for c, s, ln in st:
            s2 = s + int(rng.integers(-jit, jit + 1))
            e2 = s2 + ln + int(rng.integers(-jit, jit + 1))
            if e2 <= s2:
                e2 = s2 + 5
            if rng.random() < 0.1:
                s2, e2 = e2, s2                                   # rstart > rend rows (cluster.py:111-112)
            fl.append((c, s2, e2, max(1, abs(e2 - s2) + int(rng.integers(-2, 3)))))
        if rng.random() < 0.15:
            fl = fl[: max(1, len(fl) - 1)]
        n_aln = len(fl) + 2 + (int(rng.integers(0, 2)) if rng.random() < 0.2 else 0)
        rows += read_rows("%06d" % rng.integers(0, 10**6) + "r%d" % r, fl, n_aln=n_aln)
    df = frame(rows)
    # table order of the producer (collect_mapping_info.py:174)
    df = df.sort_values(["n_alignments", "qname", "qstart"], ascending=[False, True, True]).reset_index(drop=True)
    lens = {"chr1": 100_000_000, "chr2": 1_200_000 if rng.random() < 0.3 else 90_000_000, "chrX": 900_000,
            "chr9": 100_000_000, "chr10": 100_000_000}
    return df, lens


def run_case(name, df, lens, opts):
    ct = ColumnarTable.from_dataframe(df, lens)
    order = rh.reference_sort_order(orc.fillings_start_column(ct)) if not opts.get("filter_false") else None
    out = rh.run_reference(df, lens, **opts)
    case = {"name": name, "opts": opts, "chrom_names": [str(c) for c in ct.chrom_names],
            "chrom_len": ct.chrom_len.tolist(), "n_reads": ct.n_reads, "no_clusters": out is None}
    arrays = {k: getattr(ct, k) for k in ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments")}
    if order is not None:
        arrays["order"] = order.astype(np.int32)
    if out is not None:
        qn, cl, nr = out
        assert list(qn) == list(ct.qnames)
        arrays["cluster"] = cl.astype(np.int32)
        arrays["n_reads"] = nr.astype(np.int32)
    return case, arrays


def save(path, cases):
    meta, blob = [], {}
    for i, (c, arrs) in enumerate(cases):
        meta.append(c)
        for k, v in arrs.items():
            blob["%d/%s" % (i, k)] = v
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **blob)
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")


def small_cases():
    cases = []
    df, lens = f1_adversarial24()
    for T in (10, 10**9, 3):
        cases.append(run_case("F1_adversarial24_T%d" % T, df, lens, dict(edge_threshold=T)))
    df, lens = f2_greedy()
    cases.append(run_case("F2_greedy", df, lens, dict(cluster_mask="")))
    cases.append(run_case("F2_greedy_cut05", df, lens, dict(cluster_mask="", jaccard_cutoffs="0.3")))
    df, lens = f3_ties()
    cases.append(run_case("F3_ties", df, lens, dict()))
    df, lens = f4_mask()
    for m in ("subtelomere", "subtelomere,L1_TALEN", "L1_TALEN,chr6,notachrom", ""):
        cases.append(run_case("F4_mask[%s]" % m, df, lens, dict(cluster_mask=m)))
    df, lens = f5_float_boundaries()
    cases.append(run_case("F5_float", df, lens, dict()))
    cases.append(run_case("F5_float_p079", df, lens, dict(overlap=0.79, qlen_diff=0.05, n_alignment_diff=0.5)))
    df, lens = f6_degenerate_noclusters()
    cases.append(run_case("F6_noclusters", df, lens, dict()))
    df, lens = f6_false_names()
    cases.append(run_case("F6_false_kept", df, lens, dict()))
    cases.append(run_case("F6_false_single_cutoff", df, lens, dict(jaccard_cutoffs="0.5")))
    rng = np.random.default_rng(20261018)
    cut_lists = ["1,1,0.66,0.66,0.66,0.5", "0.5", "1,0.5,0.34"]
    for i in range(400):
        df, lens = random_table(rng)
        opts = dict(edge_threshold=int(rng.choice([1, 2, 3, 10])), jaccard_cutoffs=cut_lists[i % 3],
                    overlap=float(rng.choice([0.8, 0.5, 0.95])), cluster_mask="subtelomere" if i % 4 else "chrX",
                    qlen_diff=float(rng.choice([0.04, 0.2, 0.0])), n_alignment_diff=float(rng.choice([0.25, 0.0, 0.5])))
        cases.append(run_case("rand%03d" % i, df, lens, opts))
    save(os.path.join(HERE, "small_cases.npz"), cases)


def wide_table(rng):
    """Reads with up to 9 fillings (device paths for > 4 fillings), bigger duplicate families (many saturating reads),
    fillings of one read that overlap each other (one partner hit by several fillings) and partial structure sharing."""
    n_reads = int(rng.integers(30, 161))
    n_struct = int(rng.integers(1, 5))
    chroms = ["chr1", "chr2", "chrX"]
    structs = []
    for _ in range(n_struct):
        L = int(rng.integers(1, 10))
        st = []
        for _ in range(L):
            c = chroms[int(rng.integers(0, 3))]
            s = int(rng.integers(600_000, 600_300)) if rng.random() < 0.6 else int(rng.integers(600_000, 700_000))
            st.append((c, s, int(rng.integers(80, 500))))
        structs.append(st)
    if n_struct > 1 and rng.random() < 0.5:                       # a structure sharing a prefix with another one
        structs[-1] = structs[0][: max(1, len(structs[0]) // 2)] + structs[-1][:3]
    rows = []
    jit = int(rng.integers(0, 12))
    for r in range(n_reads):
        st = list(structs[int(rng.integers(0, n_struct))])
        if rng.random() < 0.1 and len(st) > 1:
            st.pop(int(rng.integers(0, len(st))))
        fl = []
        for c, s, ln in st[:9]:
            s2 = s + int(rng.integers(-jit, jit + 1))
            e2 = s2 + ln + int(rng.integers(-jit, jit + 1))
            if e2 <= s2:
                e2 = s2 + 5
            fl.append((c, s2, e2, max(1, abs(e2 - s2) + int(rng.integers(-2, 3)))))
        rows += read_rows("%06d" % rng.integers(0, 10**6) + "w%d" % r, fl, n_aln=len(fl) + 2)
    df = frame(rows)
    df = df.sort_values(["n_alignments", "qname", "qstart"], ascending=[False, True, True]).reset_index(drop=True)
    lens = {"chr1": 100_000_000, "chr2": 90_000_000, "chrX": 900_000, "chr9": 100_000_000, "chr10": 100_000_000}
    return df, lens


def wide_cases():
    cases = []
    rng = np.random.default_rng(20261019)
    cut_lists = ["1,1,0.66,0.66,0.66,0.5", "0.5", "1,0.5,0.34", "0.2"]
    for i in range(120):
        df, lens = wide_table(rng)
        opts = dict(edge_threshold=int(rng.choice([1, 2, 3, 10, 10])), jaccard_cutoffs=cut_lists[i % 4],
                    overlap=float(rng.choice([0.8, 0.5, 0.95, 0.0, -0.5])), cluster_mask="subtelomere" if i % 5 else "",
                    qlen_diff=float(rng.choice([0.04, 0.2])), n_alignment_diff=float(rng.choice([0.25, 0.5])))
        cases.append(run_case("wide%03d" % i, df, lens, opts))
    save(os.path.join(HERE, "wide_cases.npz"), cases)


def config_cases():
    """Named configs at the sizes the Python reference finishes in minutes; inputs are regenerated
    from the seed at test time, so only the sort permutation and the expected outputs are stored."""
    cases = []
    plan = [("C1", 1.0, dict()), ("C2", 1.0, dict()),
            ("C5", 0.004, dict())]
    for cut in synth.C3_CUTOFF_SWEEP:
        plan.append(("C3", 0.03, dict(jaccard_cutoffs=cut)))
    for name, scale, extra in plan:
        t = synth.make_config(name, scale)
        df = t.to_dataframe()
        opts = dict(cluster_mask=synth.CONFIG_MASK[name], **extra)
        case, arrs = run_case("%s@%g[%s]" % (name, scale, extra.get("jaccard_cutoffs", "default")), df, t.chr_lengths, opts)
        case["synth"] = {"config": name, "scale": scale}
        # inputs are reproducible from the seed: keep a checksum instead of the columns
        chk = int(sum(int(np.asarray(arrs[k], dtype=np.int64).sum()) * (i + 1) for i, k in enumerate(
            ("read_id", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments"))))     # chrom ids are a labelling
        case["input_checksum"] = chk
        keep = {k: arrs[k] for k in ("order", "cluster", "n_reads") if k in arrs}
        cases.append((case, keep))
        print(case["name"], "done", flush=True)
    save(os.path.join(HERE, "config_cases.npz"), cases)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "small"):
        small_cases()
    if a.only in ("", "wide"):
        wide_cases()
    if a.only in ("", "configs"):
        config_cases()

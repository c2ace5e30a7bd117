"""The reference's own boundary: the six calls main.py:227-244 makes into cluster.py, made into
fslr_b200.cluster instead, in the same order with the same arguments, followed by the pandas glue of main.py:251-257,
334-349 — compared call by call with what the UNMODIFIED reference returned for the same table
(tests/golden/dropin_cases.json.gz, written by tests/golden/make_dropin_golden.py).  Includes --filter-false
(main.py:229-230) and the default TIE_ORDER="reference"."""
import gzip
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    with gzip.open(os.path.join(HERE, "dropin_cases.json.gz")) as f:
        return json.loads(f.read())


def frame_of(case):
    return pd.DataFrame(case["rows"], columns=case["columns"])


def main_block(cluster, bed_file, chr_lengths, cluster_mask="subtelomere", jaccard_cutoffs="1,1,0.66,0.66,0.66,0.5", overlap=0.8,
               n_alignment_diff=0.25, qlen_diff=0.04, filter_false=False, edge_threshold=10):
    """main.py:209-257,334-349 with `cluster` = the module under test.  Returns (subgraphs, n_nodes, cluster_bed text | None)."""
    chromosome_mask = set([])                                            # main.py:211-216
    if cluster_mask:
        allowed = set(bed_file["chrom"])
        for item in cluster_mask.split(","):
            if item in allowed or item == "subtelomere":
                chromosome_mask.add(item)
    cutoffs = [float(i) for i in jaccard_cutoffs.split(",")]             # main.py:219
    bed_file, chr_lengths, chromosome_mask, chrom_to_num_map = cluster.rename_chromosomes(bed_file, chr_lengths, chromosome_mask)
    if filter_false:
        bed_file = cluster.delete_false(bed_file)                        # main.py:229-230
    fillings = cluster.keep_fillings(bed_file)                           # main.py:233
    data = cluster.prepare_data(fillings, chromosome_mask, chr_lengths, threshold=500_000)   # main.py:237
    interval_tree = cluster.build_interval_trees(data)                   # main.py:240
    match_data, network = cluster.query_interval_trees(interval_tree, data, overlap, cutoffs, edge_threshold, qlen_diff,
                                                       n_alignment_diff)   # main.py:242
    subgraphs = cluster.get_subgraphs(network)                           # main.py:244
    n_nodes = network.number_of_nodes()
    if len(list(subgraphs)) == n_nodes:                                  # main.py:247-249
        return subgraphs, n_nodes, None
    subg_df = pd.DataFrame(subgraphs).T                                  # main.py:251-257
    subg_long = pd.melt(subg_df, var_name="cluster", value_name="qname").dropna()
    subg_long["cluster"] = pd.to_numeric(subg_long["cluster"], errors="coerce")
    n_reads = subg_long["cluster"].value_counts().rename("n_reads")
    subg_long_reads = pd.merge(subg_long, n_reads, on="cluster")
    bed_file = bed_file.merge(subg_long_reads, on="qname", how="left")    # main.py:334-342
    n_cluster = max(subg_long_reads["cluster"]) + 1
    single = ~bed_file["qname"].isin(subg_long_reads["qname"])
    all_reads = n_cluster + len(bed_file[single]["qname"].unique())
    qname_single = bed_file[single]["qname"].unique().tolist()
    singleton_cluster_id2 = pd.DataFrame({"qname": qname_single, "cluster": range(n_cluster, all_reads)})
    bed_file["cluster"] = bed_file["cluster"].fillna(bed_file["qname"].map(singleton_cluster_id2.set_index("qname")["cluster"]))
    bed_file["n_reads"] = bed_file["n_reads"].fillna(1)
    bed_file = cluster.chrom_to_str(bed_file, chrom_to_num_map)          # main.py:344
    return subgraphs, n_nodes, bed_file.to_csv(index=False, sep="\t")    # main.py:349


CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_six_calls_like_main(case):
    from fslr_b200 import cluster
    assert cluster.TIE_ORDER == "reference"
    exp = case["expected"]
    subgraphs, n_nodes, bed = main_block(cluster, frame_of(case), dict(case["chr_lengths"]), **case["opts"])
    assert isinstance(subgraphs, list) and all(isinstance(s, set) for s in subgraphs)
    assert [sorted(s) for s in subgraphs] == exp["subgraphs"]            # same sets in the same (networkx) order
    assert n_nodes == exp["n_nodes"]
    assert bed == exp["cluster_bed"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cluster_step_writes_the_reference_files(case, tmp_path):
    """pipeline.cluster_step(out_base=...) = the same block behind one call: byte-identical mappings.cluster.bed and
    mappings.representative.bed (main.py:349-352)."""
    from fslr_b200 import pipeline
    exp = case["expected"]
    o = {k: v for k, v in case["opts"].items() if k != "edge_threshold"}
    if "edge_threshold" in case["opts"] and case["opts"]["edge_threshold"] != 10:
        pytest.skip("cluster_step fixes edge_threshold = 10 like main.py:221")
    base = str(tmp_path / "s")
    out = pipeline.cluster_step(frame_of(case), dict(case["chr_lengths"]), out_base=base, **o)
    if exp["cluster_bed"] is None:
        assert out is None and not os.path.exists(base + ".mappings.cluster.bed")
        return
    assert open(base + ".mappings.cluster.bed").read() == exp["cluster_bed"]
    assert open(base + ".mappings.representative.bed").read() == exp["representative_bed"]


def test_fast_io_integer_columns_when_every_read_is_clustered(tmp_path):
    """The GPU renderer writes integer `cluster` / `n_reads` when no read is a singleton (no NaN in main.py:334's merge)."""
    from fslr_b200 import pipeline
    case = next(c for c in CASES if c["name"] == "all_clustered_int_columns")
    bed = tmp_path / "in.mappings.bed"
    frame_of(case).to_csv(bed, index=False, sep="\t")
    base = str(tmp_path / "o")
    res = pipeline.cluster_step_fast(str(bed), dict(case["chr_lengths"]), base, representative=False)
    assert res is not None
    assert open(base + ".mappings.cluster.bed").read() == case["expected"]["cluster_bed"]


def test_zero_divisor_raises_like_the_reference():
    """aln_size 0 on a filling: cluster.py:135 divides by it -> ZeroDivisionError through the drop-in, too."""
    from fslr_b200 import cluster
    case = next(c for c in CASES if c["name"] == "F2_greedy")
    df = frame_of(case)
    df.loc[1, "aln_size"] = 0
    with pytest.raises(ZeroDivisionError):
        main_block(cluster, df, dict(case["chr_lengths"]), **case["opts"])

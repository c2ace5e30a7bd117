"""TEST INFRASTRUCTURE ONLY — drives the UNMODIFIED reference
(/root/reference/fslr/cluster.py) the way /root/reference/fslr/main.py:209-257,
334-342 does, so fixtures under tests/golden/ can be generated in the build
container.  `main.py` itself cannot be imported (needs skbio and an installed
`fslr` distribution, main.py:4,15), hence the restated glue below; every block
cites the main.py lines it follows.  Nothing on the product path may import
this module; /root/reference does not exist on the GPU box.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get("FSLR_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")

DEFAULT_JACCARD = "1,1,0.66,0.66,0.66,0.5"     # main.py:33


def reference_available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "fslr", "cluster.py"))


def import_reference_cluster():
    """Import fslr.cluster from the read-only reference with stub pysam/superintervals."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fslr import cluster  # noqa
    return cluster


def run_reference(bed_file, chr_lengths, cluster_mask="subtelomere", jaccard_cutoffs=DEFAULT_JACCARD,
                  overlap=0.8, n_alignment_diff=0.25, qlen_diff=0.04, filter_false=False,
                  edge_threshold=10, subtel=500_000, counters=None):
    """Restates main.py:209-257,334-342 around the imported cluster.py.

    bed_file : DataFrame as read by main.py:209 (string chrom / qname columns).
    Returns (qnames, cluster, n_reads) with one entry per distinct qname in order of first
    appearance in the bed table, or None when main.py:247-249 takes the early return.
    """
    cluster = import_reference_cluster()
    warnings.filterwarnings("ignore")                                    # main.py:13
    bed_file = bed_file.copy()

    chromosome_mask = set([])                                            # main.py:211-216
    if cluster_mask:
        allowed = set(bed_file["chrom"])
        for item in cluster_mask.split(","):
            if item in allowed or item == "subtelomere":
                chromosome_mask.add(item)
    cutoffs = [float(i) for i in jaccard_cutoffs.split(",")]             # main.py:219

    bed_file, chr_len, chromosome_mask, chrom_map = cluster.rename_chromosomes(
        bed_file, dict(chr_lengths), chromosome_mask)                    # main.py:227
    if filter_false:
        bed_file = cluster.delete_false(bed_file)                        # main.py:229-230
    fillings = cluster.keep_fillings(bed_file)                           # main.py:233
    data = cluster.prepare_data(fillings, chromosome_mask, chr_len, threshold=subtel)  # main.py:237
    trees = cluster.build_interval_trees(data)                           # main.py:240

    if counters is not None:                                             # optional call counting
        orig = cluster.overall_jaccard_similarity
        def counted(*a, **k):
            counters["pair_tests"] = counters.get("pair_tests", 0) + 1
            return orig(*a, **k)
        cluster.overall_jaccard_similarity = counted
    try:
        match_data, network = cluster.query_interval_trees(
            trees, data, overlap, cutoffs, edge_threshold, qlen_diff, n_alignment_diff)  # main.py:242
    finally:
        if counters is not None:
            cluster.overall_jaccard_similarity = orig
    subgraphs = cluster.get_subgraphs(network)                           # main.py:244
    if counters is not None:
        counters["edges"] = network.number_of_edges()
        counters["components"] = len(subgraphs)
    if len(list(subgraphs)) == network.number_of_nodes():                # main.py:247-249
        return None

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
    bed_file["cluster"] = bed_file["cluster"].fillna(
        bed_file["qname"].map(singleton_cluster_id2.set_index("qname")["cluster"]))
    bed_file["n_reads"] = bed_file["n_reads"].fillna(1)

    per_read = bed_file.drop_duplicates("qname", keep="first")
    return (per_read["qname"].to_numpy(),
            per_read["cluster"].to_numpy().astype(np.int64),
            per_read["n_reads"].to_numpy().astype(np.int64))


def reference_sort_order(starts):
    """The permutation `DataFrame.sort_values('start')` applies (cluster.py:114): pandas sorts an
    int64 column with numpy's default (unstable) quicksort; same array + same numpy → same order."""
    df = pd.DataFrame({"start": np.asarray(starts, dtype=np.int64)})
    return df.sort_values("start").index.to_numpy()


FSLR_VERSION_FOR_TESTS = "0.0.test"


def run_reference_mapping_info(bam_path, out_path, regions_path=None, primers=None):
    """Runs the UNMODIFIED collect_mapping_info.mapping_info (collect_mapping_info.py:19-181) on a BAM file through the
    stub pysam; `importlib.metadata.version("fslr")` (no installed distribution here) is answered with a constant.
    Returns None, or the SystemExit / exception the reference ended with (its `quit()` calls, KeyError on a missing tag)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import contextlib
    import io
    from fslr import collect_mapping_info as cmi
    cmi.version = lambda name: FSLR_VERSION_FOR_TESTS
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            cmi.mapping_info(bam_path, out_path, regions_path, primers or {})
    except SystemExit as e:
        return e
    except Exception as e:          # noqa: BLE001 - the reference's own failure mode is the thing recorded
        return e
    return None


def run_reference_block(bed_file, chr_lengths, cluster_mask="subtelomere", jaccard_cutoffs=DEFAULT_JACCARD,
                        overlap=0.8, n_alignment_diff=0.25, qlen_diff=0.04, filter_false=False,
                        edge_threshold=10, subtel=500_000):
    """The whole clustering block main.py:209-257,334-352 around the imported cluster.py, keeping what the drop-in tests
    compare: the `list[set[str]]` cluster.get_subgraphs returned (main.py:244), the permutation the reference's own
    unstable sort applied to the fillings (cluster.py:114), and the text of `<base>.mappings.cluster.bed` (main.py:349)
    and `<base>.mappings.representative.bed` (main.py:351-352).  Same statements as run_reference above, nothing dropped.
    Returns a dict; "cluster_bed" / "representative_bed" are None when main.py:247-249 returns early."""
    cluster = import_reference_cluster()
    warnings.filterwarnings("ignore")                                    # main.py:13
    bed_file = bed_file.copy()
    chromosome_mask = set([])                                            # main.py:211-216
    if cluster_mask:
        allowed = set(bed_file["chrom"])
        for item in cluster_mask.split(","):
            if item in allowed or item == "subtelomere":
                chromosome_mask.add(item)
    cutoffs = [float(i) for i in jaccard_cutoffs.split(",")]             # main.py:219
    bed_file, chr_len, chromosome_mask, chrom_map = cluster.rename_chromosomes(
        bed_file, dict(chr_lengths), chromosome_mask)                    # main.py:227
    if filter_false:
        bed_file = cluster.delete_false(bed_file)                        # main.py:229-230
    fillings = cluster.keep_fillings(bed_file)                           # main.py:233
    # the permutation sort_values('start') is about to apply inside prepare_data (same column, same routine)
    starts = np.minimum(fillings["rstart"].to_numpy(), fillings["rend"].to_numpy())
    order = reference_sort_order(starts)
    data = cluster.prepare_data(fillings, chromosome_mask, chr_len, threshold=subtel)  # main.py:237
    trees = cluster.build_interval_trees(data)                           # main.py:240
    match_data, network = cluster.query_interval_trees(
        trees, data, overlap, cutoffs, edge_threshold, qlen_diff, n_alignment_diff)  # main.py:242
    subgraphs = cluster.get_subgraphs(network)                           # main.py:244
    out = {"subgraphs": [sorted(s) for s in subgraphs], "n_nodes": network.number_of_nodes(), "order": order,
           "cluster_bed": None, "representative_bed": None}
    if len(list(subgraphs)) == network.number_of_nodes():                # main.py:247-249
        return out
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
    bed_file["cluster"] = bed_file["cluster"].fillna(
        bed_file["qname"].map(singleton_cluster_id2.set_index("qname")["cluster"]))
    bed_file["n_reads"] = bed_file["n_reads"].fillna(1)
    bed_file = cluster.chrom_to_str(bed_file, chrom_map)                 # main.py:344
    out["cluster_bed"] = bed_file.to_csv(index=False, sep="\t")          # main.py:349
    rep = cluster.choose_alignment(bed_file)                             # main.py:351-352
    out["representative_bed"] = rep.to_csv(index=False, sep="\t")
    return out

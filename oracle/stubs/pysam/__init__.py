"""TEST INFRASTRUCTURE ONLY — stand-in for `pysam` so that the unmodified
/root/reference/fslr/cluster.py can be imported (cluster.py:4).  The only use on
the clustering path is `get_chromosome_lengths` (cluster.py:173-175), which the
harness bypasses by passing the {chrom: length} dict directly."""


class AlignmentFile:  # pragma: no cover - never constructed by the harness
    def __init__(self, *a, **k):
        raise RuntimeError("pysam stub: BAM access is not available; pass chr_lengths as a dict")

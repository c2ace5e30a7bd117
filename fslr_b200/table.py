"""Host-side columnar (SoA) form of `<name>.mappings.bed` and of the clustering options.

This is the layout the C ABI (include/fslr_b200.h) consumes: one int32 column per field the
clustering step reads — /root/reference/fslr/cluster.py touches only chrom, rstart, rend, qname,
n_alignments, aln_size, qstart, qend — with qname replaced by a dense read id in
order of first appearance (that order is also the singleton numbering order of
/root/reference/fslr/main.py:336-341) and chrom by a small integer id
(cluster.rename_chromosomes, cluster.py:34-43; only equality of ids is ever used).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

DEFAULT_JACCARD = "1,1,0.66,0.66,0.66,0.5"      # main.py:33
SUBTEL_DEFAULT = 500_000                        # main.py:237
EDGE_THRESHOLD_DEFAULT = 10                     # main.py:221


@dataclass
class ColumnarTable:
    read_id: np.ndarray
    chrom: np.ndarray
    rstart: np.ndarray
    rend: np.ndarray
    aln_size: np.ndarray
    qstart: np.ndarray
    qend: np.ndarray
    n_alignments: np.ndarray
    n_reads: int
    chrom_names: List[str]
    chrom_len: np.ndarray                       # int64 [n_chrom], 0 = not in the BAM header
    strand: Optional[np.ndarray] = None         # carried for completeness; never read by the reference
    qnames: Optional[np.ndarray] = None         # [n_reads] read id -> qname
    alignment_score: Optional[np.ndarray] = None

    @property
    def n_rows(self):
        return int(self.read_id.shape[0])

    @property
    def n_chrom(self):
        return len(self.chrom_names)

    @staticmethod
    def _i32(x):
        a = np.asarray(x)
        if a.size and (a.min() < -2**31 or a.max() >= 2**31):
            raise ValueError("column does not fit int32")
        return np.ascontiguousarray(a, dtype=np.int32)

    @classmethod
    def from_dataframe(cls, bed_df, chr_lengths):
        """bed_df as read at main.py:209 (string or already-renamed integer `chrom`)."""
        import pandas as pd
        rid, qn = pd.factorize(bed_df["qname"], sort=False)
        cid, cn = pd.factorize(bed_df["chrom"], sort=False)
        names = [c for c in cn.tolist()]
        clen = np.array([int(chr_lengths.get(c, 0) or 0) for c in names], dtype=np.int64)
        return cls(read_id=cls._i32(rid), chrom=cls._i32(cid), rstart=cls._i32(bed_df["rstart"]),
                   rend=cls._i32(bed_df["rend"]), aln_size=cls._i32(bed_df["aln_size"]),
                   qstart=cls._i32(bed_df["qstart"]), qend=cls._i32(bed_df["qend"]),
                   n_alignments=cls._i32(bed_df["n_alignments"]), n_reads=int(len(qn)),
                   chrom_names=names, chrom_len=clen, qnames=np.asarray(qn, dtype=object),
                   alignment_score=(np.asarray(bed_df["alignment_score"], dtype=np.float64)
                                    if "alignment_score" in bed_df else None))

    @classmethod
    def from_synth(cls, t):
        """From fslr_b200.synth.MappingsTable without materialising strings."""
        rid, n = t.read_ids()
        clen = np.array([t.chr_lengths.get(c, 0) for c in t.chrom_names], dtype=np.int64)
        return cls(read_id=rid, chrom=t.chrom.astype(np.int32), rstart=t.rstart, rend=t.rend,
                   aln_size=t.aln_size, qstart=t.qstart, qend=t.qend, n_alignments=t.n_alignments,
                   n_reads=n, chrom_names=list(t.chrom_names), chrom_len=clen, strand=t.strand,
                   alignment_score=t.alignment_score.astype(np.float64))


@dataclass
class ClusterParams:
    """The clustering options of main.py:33-37,219-223,237 in numeric form."""
    jaccard_cutoffs: List[float] = field(default_factory=lambda: [float(i) for i in DEFAULT_JACCARD.split(",")])
    overlap: float = 0.8
    n_alignment_diff: float = 0.25
    qlen_diff: float = 0.04
    chrom_masked: Optional[np.ndarray] = None   # uint8 [n_chrom]
    mask_subtelomere: bool = False
    subtel: int = SUBTEL_DEFAULT
    edge_threshold: int = EDGE_THRESHOLD_DEFAULT

    @classmethod
    def from_options(cls, table, cluster_mask="subtelomere", jaccard_cutoffs=DEFAULT_JACCARD, overlap=0.8,
                     n_alignment_diff=0.25, qlen_diff=0.04, edge_threshold=EDGE_THRESHOLD_DEFAULT,
                     subtel=SUBTEL_DEFAULT):
        """Option strings exactly as main.py parses them (main.py:211-223)."""
        masked = np.zeros(table.n_chrom, dtype=np.uint8)
        sub = False
        if cluster_mask:
            allowed = {str(c): i for i, c in enumerate(table.chrom_names)}
            for item in str(cluster_mask).split(","):
                if item == "subtelomere":
                    sub = True
                elif item in allowed:
                    masked[allowed[item]] = 1
        if isinstance(jaccard_cutoffs, str):
            cut = [float(i) for i in jaccard_cutoffs.split(",")]
        else:
            cut = [float(i) for i in jaccard_cutoffs]
        if not cut:
            raise ValueError("--jaccard-cutoffs must name at least one value")
        return cls(jaccard_cutoffs=cut, overlap=float(overlap), n_alignment_diff=float(n_alignment_diff),
                   qlen_diff=float(qlen_diff), chrom_masked=masked, mask_subtelomere=sub, subtel=int(subtel),
                   edge_threshold=int(edge_threshold))

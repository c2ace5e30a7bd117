"""Deterministic synthetic `mappings.bed` tables for the five BASELINE.json configs.

The generator writes the table schema of the reference's producer
(/root/reference/fslr/collect_mapping_info.py:79-94,174-181): PCR-duplicate *families* of split-read
amplicons that share a junction structure, first/last alignment ("bread") at the primer locus,
`n_alignments - 2` fillings elsewhere, rows sorted by (n_alignments desc, qname, qstart).
Everything is vectorised numpy so the 10M-read config is generated in seconds; strings (qname,
chrom names) are only materialised by `to_dataframe()` for the sizes the Python reference can run.
"""
from dataclasses import dataclass, field

import numpy as np

SEED_BASE = 20261018

# T2T-CHM13v2.0 chromosome lengths + one short contig that can only be masked by name
GENOME = [
    ("chr1", 248387328), ("chr2", 242696752), ("chr3", 201105948), ("chr4", 193574945),
    ("chr5", 182045439), ("chr6", 172126628), ("chr7", 160567428), ("chr8", 146259331),
    ("chr9", 150617247), ("chr10", 134758134), ("chr11", 135127769), ("chr12", 133324548),
    ("chr13", 113566686), ("chr14", 101161492), ("chr15", 99753195), ("chr16", 96330374),
    ("chr17", 84276897), ("chr18", 80542538), ("chr19", 61707364), ("chr20", 66210255),
    ("chr21", 45090682), ("chr22", 51324926), ("chrX", 154259566), ("chrY", 62460029),
    ("L1_TALEN", 8000),
]
CHROM_NAMES = [g[0] for g in GENOME]
CHROM_LEN = np.array([g[1] for g in GENOME], dtype=np.int64)
N_NUCLEAR = 24
L1_IDX = 24
# primer name -> (chrom index, True if the locus is at the q (far) end)
PRIMER_LOCUS = {"21q1": (20, True), "17p6": (16, False)}
SUBTEL = 500_000


@dataclass
class MappingsTable:
    """Columnar mappings.bed (one entry per alignment row, already in table order)."""
    chrom: np.ndarray          # int32 index into chrom_names
    rstart: np.ndarray         # int32, 1-based (collect_mapping_info.py:71)
    rend: np.ndarray           # int32, inclusive
    read_key: np.ndarray       # int64, the hex prefix of qname
    n_alignments: np.ndarray   # int32
    aln_size: np.ndarray       # int32, qend - qstart
    qstart: np.ndarray         # int32
    qend: np.ndarray           # int32
    strand: np.ndarray         # int8, 0 '+', 1 '-'
    qlen: np.ndarray           # int32
    alignment_score: np.ndarray  # int32
    primer: np.ndarray         # int8 per row, index into primers
    primers: list = field(default_factory=list)
    chrom_names: list = field(default_factory=lambda: list(CHROM_NAMES))
    chr_lengths: dict = field(default_factory=lambda: {n: int(l) for n, l in GENOME})
    name: str = ""
    read_id: np.ndarray = None  # int32 dense read ids in order of first appearance (== pandas.factorize(qname))
    n_reads: int = 0

    @property
    def n_rows(self):
        return int(self.chrom.shape[0])

    def read_ids(self):
        """Dense read ids in order of first appearance (what pandas.factorize(qname) yields)."""
        return self.read_id, self.n_reads

    def qnames(self):
        pn = np.array(self.primers)[self.primer]
        return np.array(["%08x-read.0.9_0.9.%sF_%sR" % (k, p, p) for k, p in zip(self.read_key.tolist(), pn.tolist())],
                        dtype=object)

    def to_dataframe(self):
        import pandas as pd
        names = np.array(self.chrom_names, dtype=object)
        return pd.DataFrame({
            "chrom": names[self.chrom],
            "rstart": self.rstart.astype(np.int64), "rend": self.rend.astype(np.int64),
            "qname": self.qnames(),
            "n_alignments": self.n_alignments.astype(np.int64),
            "aln_size": self.aln_size.astype(np.int64),
            "qstart": self.qstart.astype(np.int64), "qend": self.qend.astype(np.int64),
            "strand": np.where(self.strand == 0, "+", "-"),
            "mapq": np.full(self.n_rows, 60, dtype=np.int64),
            "qlen": self.qlen.astype(np.int64),
            "alignment_score": self.alignment_score.astype(np.int64),
            "short_anchor<50bp": np.zeros(self.n_rows, dtype=np.int64),
            "fslr_version": np.full(self.n_rows, "0.3.10", dtype=object),
            "inferred_by_primer": np.zeros(self.n_rows, dtype=np.int64),
            "seq": np.full(self.n_rows, "", dtype=object),
        })

    def write_bed(self, path):
        self.to_dataframe().to_csv(path, index=False, sep="\t")


def _family_sizes(rng, n_reads, mean=4.0):
    sizes = []
    total = 0
    while total < n_reads:
        s = rng.geometric(1.0 / mean, size=max(1024, int(n_reads / mean * 1.1)))
        sizes.append(s)
        total += int(s.sum())
    s = np.concatenate(sizes)
    c = np.cumsum(s)
    k = int(np.searchsorted(c, n_reads))
    s = s[:k + 1].copy()
    s[k] -= int(c[k] - n_reads)
    return s[s > 0]


def make_table(n_reads, primers=("21q1",), seed=SEED_BASE, subtel_frac=0.0, l1_frac=0.0,
               hotspot_reads=0, hotspot_fillings=2, name="", naln_range=(2, 7), genome_scale=1.0):
    """Generate `n_reads` reads (`hotspot_reads` of them one giant family).  `genome_scale` < 1 shortens every nuclear
    chromosome and the subtelomere window by that factor: n_reads * genome_scale reads on such a genome have the filling
    density (chance overlaps per filling) of the full-size table — the density-preserving sample bench.py times on the CPU."""
    rng = np.random.default_rng(seed)
    CHROM_LEN = globals()["CHROM_LEN"].copy()
    CHROM_LEN[:N_NUCLEAR] = (CHROM_LEN[:N_NUCLEAR] * genome_scale).astype(np.int64)
    SUBTEL = max(8000, int(globals()["SUBTEL"] * genome_scale))
    primers = list(primers)
    n_bg = n_reads - hotspot_reads
    fam_size = _family_sizes(rng, n_bg) if n_bg > 0 else np.zeros(0, dtype=np.int64)
    n_fam = fam_size.shape[0]
    fam_naln = rng.integers(naln_range[0], naln_range[1], size=n_fam)   # U{2..6} by default
    if hotspot_reads:
        fam_size = np.concatenate([fam_size, [hotspot_reads]])
        fam_naln = np.concatenate([fam_naln, [hotspot_fillings + 2]])
        n_fam += 1
    fam_primer = rng.integers(0, len(primers), size=n_fam)

    # ---- family-level alignment templates (one row per alignment of the family structure)
    t_fam = np.repeat(np.arange(n_fam), fam_naln)                    # template row -> family
    t_first = np.cumsum(fam_naln) - fam_naln
    t_k = np.arange(t_fam.shape[0]) - t_first[t_fam]                 # alignment index within read
    t_bread = (t_k == 0) | (t_k == fam_naln[t_fam] - 1)
    nt = t_fam.shape[0]
    t_len = rng.integers(80, 1501, size=nt)
    # fillings: uniform over the nuclear genome, optionally forced into subtelomeres / the contig
    w = CHROM_LEN[:N_NUCLEAR] / CHROM_LEN[:N_NUCLEAR].sum()
    t_chrom = rng.choice(N_NUCLEAR, size=nt, p=w)
    u = rng.random(nt)
    t_pos = (1 + u * (CHROM_LEN[t_chrom] - 4000)).astype(np.int64)
    mode = rng.random(nt)
    in_sub = (mode < subtel_frac) & ~t_bread
    side = rng.random(nt) < 0.5
    sub_pos = (1 + rng.random(nt) * (SUBTEL - 4000)).astype(np.int64)
    t_pos = np.where(in_sub, np.where(side, sub_pos, CHROM_LEN[t_chrom] - 3000 - sub_pos), t_pos)
    in_l1 = (mode >= subtel_frac) & (mode < subtel_frac + l1_frac) & ~t_bread
    t_chrom = np.where(in_l1, L1_IDX, t_chrom)
    t_pos = np.where(in_l1, (10 + rng.random(nt) * (CHROM_LEN[L1_IDX] - 1700)).astype(np.int64), t_pos)
    t_len = np.where(in_l1, np.minimum(t_len, 1500), t_len)
    # breads at the primer locus (within 500 kb of the named chromosome end)
    p_chr = np.array([PRIMER_LOCUS[p][0] for p in primers])[fam_primer[t_fam]]
    p_far = np.array([PRIMER_LOCUS[p][1] for p in primers])[fam_primer[t_fam]]
    off = (2000 + rng.random(nt) * (SUBTEL - 6000)).astype(np.int64)
    b_pos = np.where(p_far, CHROM_LEN[p_chr] - off, off)
    t_chrom = np.where(t_bread, p_chr, t_chrom)
    t_pos = np.where(t_bread, b_pos, t_pos)
    t_len = np.where(t_bread, rng.integers(100, 400, size=nt), t_len)
    t_strand = rng.integers(0, 2, size=nt)

    # ---- expand templates to reads (family copies) and jitter
    fam_rows0 = t_first                                               # first template row per family
    read_fam = np.repeat(np.arange(n_fam), fam_size)
    R = read_fam.shape[0]
    read_naln = fam_naln[read_fam]
    row_read = np.repeat(np.arange(R), read_naln)
    r_first = np.cumsum(read_naln) - read_naln
    row_k = np.arange(row_read.shape[0]) - r_first[row_read]
    row_t = fam_rows0[read_fam[row_read]] + row_k
    A = row_read.shape[0]
    rstart = t_pos[row_t] + rng.integers(-3, 4, size=A)
    rend = t_pos[row_t] + t_len[row_t] + rng.integers(-3, 4, size=A)
    rstart = np.maximum(rstart, 1)
    span = rend - rstart
    aln = np.maximum(span + rng.integers(-2, 3, size=A), 1)
    gap = rng.integers(0, 5, size=A)
    gap[r_first] = 0
    cum = np.cumsum(aln + gap)
    base = (cum - aln - gap)[r_first][row_read] if A else cum
    qend = cum - base
    qstart = qend - aln
    qlen = qend[r_first + read_naln - 1][row_read] if A else qend

    # ---- qname keys: random permutation so name order is unrelated to family membership
    key = rng.permutation(R).astype(np.int64)
    # table order of collect_mapping_info.py:174: n_alignments desc, qname asc, qstart asc.  Rows of a read are
    # generated in qstart order and keys are unique, so ordering the READS and expanding is the same permutation.
    ro = np.argsort(((max(6, naln_range[1]) - read_naln).astype(np.int64) << 40) | key)
    na_s = read_naln[ro]
    new_first = np.cumsum(na_s) - na_s
    order = np.repeat(r_first[ro] - new_first, na_s) + np.arange(A)
    read_id = np.repeat(np.arange(R, dtype=np.int32), na_s)          # dense ids in order of first appearance
    row_key = key[row_read]
    return MappingsTable(
        chrom=t_chrom[row_t][order].astype(np.int32),
        rstart=rstart[order].astype(np.int32), rend=rend[order].astype(np.int32),
        read_key=row_key[order],
        n_alignments=read_naln[row_read][order].astype(np.int32),
        aln_size=aln[order].astype(np.int32),
        qstart=qstart[order].astype(np.int32), qend=qend[order].astype(np.int32),
        strand=t_strand[row_t][order].astype(np.int8),
        qlen=qlen[order].astype(np.int32),
        alignment_score=(aln[order] * 9 // 5).astype(np.int32),
        primer=fam_primer[read_fam[row_read]][order].astype(np.int8),
        primers=primers, name=name, read_id=read_id, n_reads=R,
        chr_lengths={n: int(l) for n, l in zip(CHROM_NAMES, CHROM_LEN)})


# name -> (generator kwargs, clustering parameters handed to main.py's options)
CONFIGS = {
    "C1": dict(n_reads=5_000, primers=("21q1",), seed=SEED_BASE + 1),
    "C2": dict(n_reads=100_000, primers=("21q1", "17p6"), seed=SEED_BASE + 2),
    "C3": dict(n_reads=1_000_000, primers=("21q1", "17p6"), seed=SEED_BASE + 3, subtel_frac=0.15, l1_frac=0.05),
    "C4": dict(n_reads=10_000_000, primers=("21q1", "17p6"), seed=SEED_BASE + 4),
    "C5": dict(n_reads=2_000_000, primers=("21q1", "17p6"), seed=SEED_BASE + 5, hotspot_reads=500_000),
}
CONFIG_MASK = {"C1": "subtelomere", "C2": "subtelomere", "C3": "subtelomere,L1_TALEN",
               "C4": "subtelomere", "C5": "subtelomere"}
C3_CUTOFF_SWEEP = ["1,1,0.66,0.66,0.66,0.5", "1,1,1,1,1,1", "1,0.5,0.5,0.5,0.5,0.5", "0.5", "0.34"]


def make_config(name, scale=1.0, genome_scale=1.0):
    """Table for a named config; `scale` < 1 shrinks the read counts proportionally (tests); `genome_scale`: see make_table."""
    kw = dict(CONFIGS[name], genome_scale=genome_scale)
    kw["n_reads"] = max(16, int(round(kw["n_reads"] * scale)))
    if "hotspot_reads" in kw:
        kw["hotspot_reads"] = int(round(kw["hotspot_reads"] * scale))
    return make_table(name=name if scale == 1.0 else "%s@%g" % (name, scale), **kw)

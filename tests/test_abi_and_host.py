"""CPU tests: the C-ABI library builds/loads and exports every symbol include/fslr_b200.h declares (no compute
without a GPU), and the host-side logic (threshold tables, option parsing, reference-compatible helpers)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from fslr_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "fslr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fslrc_\w+)\s*\(", hdr))
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.fslrc_version() >= 1
    lib.fslrc_stage_name.restype = ctypes.c_char_p
    assert lib.fslrc_stage_name(7) == b"pair_kernel"


def test_struct_layout_matches_header():
    from fslr_b200 import _native
    assert ctypes.sizeof(_native.Table) == 8 * 2 + 8 * 8 + 8 + 8 + 8 * 3 + 8 + 8 * 3
    assert ctypes.sizeof(_native.Stats) == 8 * 11 + 4 + 4 + 4 * _native.N_STAGES
    assert _native.Params.umax.offset == 24 and _native.Params.edge_threshold.offset == 24 + 4 * 65 + 4
    assert ctypes.sizeof(_native.BamInfo) == 8 * 4 + 4 * 2 + 8 * 15 + 4 + 4
    # the field order of the pointer block of fslrc_bam_info is the one the header declares
    hdr = open(os.path.join(ROOT, "include", "fslr_b200.h")).read()
    block = hdr[hdr.index("int32_t overlaps_as_float;"):hdr.index("} fslrc_bam_info;")]
    names = re.findall(r"\*(\w+)", re.sub(r"/\*.*?\*/", "", block, flags=re.S))
    assert tuple(n.replace("n_alignments", "n_alignments") for n in names) == _native.BAM_COLUMNS


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fslr_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_umax_table_is_the_reference_comparison():
    """umax[n] is the largest union with n/union >= cutoff(n) in float64 (cluster.py:165-170,218-219)."""
    from fslr_b200.engine import umax_table
    for cut in ([1, 1, 0.66, 0.66, 0.66, 0.5], [0.5], [0.34], [1.0], [0.0], [1, 0.5, 0.34], [1.5]):
        um = umax_table([float(c) for c in cut], 12)
        for n in range(1, 13):
            t = cut[n - 1] if n - 1 < len(cut) else cut[-1]
            for u in range(n, 4 * n + 8):
                assert ((n / u) >= t) == (u <= um[n]), (cut, n, u)
    # default cutoffs -> max len1+len2 = umax[n] + n = [2, 4, 7, 10, 12, 18, 21, 24]
    um = umax_table([1, 1, 0.66, 0.66, 0.66, 0.5], 8)
    assert [um[n] + n for n in range(1, 9)] == [2, 4, 7, 10, 12, 18, 21, 24]


def test_integer_thresholds_equal_float_tests():
    """T(a) = min{o : o/a >= p} reproduces min(ov/a1, ov/a2) >= p (cluster.py:133-136,157) exactly."""
    from tests.proto_model import _thr
    rng = np.random.default_rng(3)
    for p in (0.8, 0.5, 0.79, 0.95, 1.0, 1 / 3, 0.0):
        a = rng.integers(1, 4000, size=3000)
        b = rng.integers(1, 4000, size=3000)
        ov = rng.integers(0, 4200, size=3000)
        ref = np.minimum(ov / a, ov / b) >= p
        T = np.array([max(_thr(int(x), p), _thr(int(y), p)) for x, y in zip(a, b)])
        assert np.array_equal(ref, ov >= T)
    assert _thr(1000, 0.8) == 800 and _thr(999, 0.8) == 800       # 800/1000 passes, 799/999 fails


def test_option_parsing_like_main():
    from fslr_b200.table import ClusterParams, ColumnarTable
    df = pd.DataFrame({"chrom": ["chr1", "L1_TALEN", "chr1"], "rstart": [1, 2, 3], "rend": [5, 6, 7], "qname": ["a", "a", "a"],
                       "n_alignments": [3, 3, 3], "aln_size": [4, 4, 4], "qstart": [0, 4, 8], "qend": [4, 8, 12]})
    t = ColumnarTable.from_dataframe(df, {"chr1": 1000})
    p = ClusterParams.from_options(t, cluster_mask="subtelomere,L1_TALEN,chr9", jaccard_cutoffs="1,0.5")
    assert p.mask_subtelomere and p.chrom_masked.tolist() == [0, 1]          # chr9 is not in the table: ignored (main.py:213-216)
    assert p.jaccard_cutoffs == [1.0, 0.5]
    assert t.chrom_len.tolist() == [1000, 0]


def test_rename_chromosomes_like_reference():
    from fslr_b200 import cluster
    df = pd.DataFrame({"chrom": ["chr10", "chr2", "chrX", "chr2"]})
    out, lens, mask, m = cluster.rename_chromosomes(df, {"chr2": 5, "chr10": 7, "chrM": 1}, {"subtelomere", "chrX"})
    assert m["chr2"] == 1 and m["chr10"] == 2 and m["chrX"] == 3
    assert out["chrom"].tolist() == [2, 1, 3, 1]
    assert lens[1] == 5 and lens[2] == 7 and None in lens
    assert sorted(map(str, mask)) == ["3", "subtelomere"]


def test_synth_tables_are_well_formed():
    from fslr_b200 import synth
    t = synth.make_config("C2", 0.05)
    df = t.to_dataframe()
    srt = df.sort_values(["n_alignments", "qname", "qstart"], ascending=[False, True, True], kind="stable")
    assert srt.index.tolist() == df.index.tolist()                          # collect_mapping_info.py:174 order
    assert (df.groupby("qname")["n_alignments"].nunique() == 1).all()
    assert (df.groupby("qname").size() == df.groupby("qname")["n_alignments"].first()).all()
    assert (df["aln_size"] > 0).all() and (df["qend"] - df["qstart"] == df["aln_size"]).all()
    rid, n = t.read_ids()
    assert np.array_equal(rid, pd.factorize(df["qname"])[0]) and n == df["qname"].nunique()


def test_get_chromosome_lengths_reads_bam_header(tmp_path):
    """cluster.get_chromosome_lengths (cluster.py:173-175) without pysam: a BAM header split over several BGZF blocks."""
    import struct
    import zlib
    from fslr_b200 import cluster
    refs = [("chr%d" % i, 1_000_000 + 37 * i) for i in range(1, 60)] + [("L1_TALEN", 8000)]
    text = b"@HD\tVN:1.6\tSO:unsorted\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), l) for n, l in refs)
    raw = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for n, l in refs:
        raw += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    raw += b"\x00" * 100                                                       # (first bytes of the alignment section)

    def bgzf(chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        data = c.compress(chunk) + c.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(data) + 25) + data +
                struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    path = tmp_path / "t.bam"
    path.write_bytes(b"".join(bgzf(raw[i:i + 300]) for i in range(0, len(raw), 300)))
    assert cluster.get_chromosome_lengths(str(path)) == dict(refs)
    (tmp_path / "bad.bam").write_bytes(b"not a bam file at all, not even gzip")
    with pytest.raises(ValueError):
        cluster.get_chromosome_lengths(str(tmp_path / "bad.bam"))

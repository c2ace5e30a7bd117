"""GPU parity of the table-producer row through the C ABI (fslrc_bam_*): the TSV rendered on the
device must equal, byte for byte, the file the unmodified reference wrote (tests/golden/bam_cases) and, on larger random
inputs, the oracle restatement; inputs the reference stops on must raise."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES_DIR = os.path.join(ROOT, "tests", "golden", "bam_cases")
CASES = json.load(open(os.path.join(CASES_DIR, "cases.json")))


def _first_diff(got, want):
    g, w = got.split(b"\n"), want.split(b"\n")
    for i, (a, b) in enumerate(zip(g, w)):
        if a != b:
            return "line %d:\n got %r\nwant %r" % (i, a[:260], b[:260])
    return "line counts %d / %d" % (len(g), len(w))


@pytest.mark.parametrize("device_inflate", [True, False])
@pytest.mark.parametrize("name", sorted(n for n in CASES if not CASES[n]["reference_error"]))
def test_fixture_tsv_is_byte_identical(name, device_inflate):
    from fslr_b200 import mapping_info as mi
    c = CASES[name]
    reg = os.path.join(CASES_DIR, name + ".regions.bed") if c["regions"] else None
    t = mi.read_bam_table(os.path.join(CASES_DIR, name + ".bam"), reg, c["primers"], device_inflate=device_inflate)
    try:
        got = t.mappings_bed_bytes(c["fslr_version"]).tobytes()
        want = open(os.path.join(CASES_DIR, name + ".mappings.bed"), "rb").read()
        assert got == want, _first_diff(got, want)
        df = t.to_dataframe()
        assert len(df) == t.n_rows and df["qname"].nunique() == t.n_reads
        rid = t.column("read_id")
        assert rid[0] == 0 and (np.diff(rid) >= 0).all() and (np.diff(rid) <= 1).all()      # numbered in output order
    finally:
        t.close()


@pytest.mark.parametrize("name", sorted(n for n in CASES if CASES[n]["reference_error"]))
def test_inputs_the_reference_stops_on_raise(name):
    from fslr_b200 import _native, mapping_info as mi
    for device_inflate in (True, False):
        with pytest.raises(_native.FslrError):
            mi.read_bam_table(os.path.join(CASES_DIR, name + ".bam"), None, CASES[name]["primers"], device_inflate=device_inflate)


@pytest.mark.parametrize("seed,kw", [(21, dict(n_reads=4000)), (22, dict(n_reads=3000, name_style="prefix", p_single=0.6)),
                                     (23, dict(n_reads=2500, with_seq_on_supp=True, max_aln=12))])
def test_random_bam_against_oracle(tmp_path, seed, kw):
    from fslr_b200 import mapping_info as mi, synth_bam as sb
    from oracle import mapping_info_oracle as mo
    refs, recs, primers = sb.make_alignments(seed=seed, **kw)
    bam = str(tmp_path / "t.bam")
    sb.write_bam(bam, refs, [sb.encode_record(*r) for r in recs])
    reg = str(tmp_path / "r.bed")
    open(reg, "w").write("chr1\t1000\t90000000\nchr21\t5\t20000000\nL1_TALEN\t100\t4000\nchr9\t1\t2\n")
    want = mo.mapping_tsv(mo.mapping_rows(bam, mo.read_regions(reg), primers, "9.9"), True).encode()
    t = mi.mapping_info(bam, str(tmp_path / "out.bed"), reg, primers, fslr_version="9.9")
    try:
        got = open(str(tmp_path / "out.bed"), "rb").read()
        assert got == want, _first_diff(got, want)
    finally:
        t.close()


def test_bam_table_feeds_the_clustering_step(tmp_path):
    """BAM -> device table -> clusters equals BAM -> mappings.bed file -> GPU parse -> clusters."""
    from fslr_b200 import mapping_info as mi, synth_bam as sb, tsv
    refs, recs, primers = sb.make_alignments(3000, seed=31, refs=[("chr1", 3_000_000), ("chr2", 2_500_000), ("chr21", 2_000_000)],
                                             p_single=0.1)
    bam = str(tmp_path / "t.bam")
    sb.write_bam(bam, refs, [sb.encode_record(*r) for r in recs])
    t = mi.mapping_info(bam, str(tmp_path / "m.bed"), None, primers, fslr_version="9.9")
    lens = dict(refs)
    try:
        a = t.cluster(cluster_mask="subtelomere")
        names_a = t.qnames()
    finally:
        t.close()
    p = tsv.read_mappings_bed(str(tmp_path / "m.bed"), lens)
    try:
        b = p.cluster(cluster_mask="subtelomere")
        names_b = p.qnames()
    finally:
        p.close()
    assert list(names_a) == list(names_b)
    assert np.array_equal(a.cluster, b.cluster) and np.array_equal(a.n_reads, b.n_reads)


def test_bam_to_clusters_writes_the_three_files(tmp_path):
    import pandas as pd
    from fslr_b200 import pipeline, synth_bam as sb
    refs, recs, primers = sb.make_alignments(1500, seed=41, refs=[("chr1", 3_000_000), ("chr2", 2_500_000)], p_single=0.1)
    bam = str(tmp_path / "t.bam")
    sb.write_bam(bam, refs, [sb.encode_record(*r) for r in recs])
    base = str(tmp_path / "s")
    res = pipeline.bam_to_clusters(bam, primers, base, fslr_version="9.9")
    bed = pd.read_csv(base + ".mappings.bed", sep="\t")
    if res is None:
        return
    cl = pd.read_csv(base + ".mappings.cluster.bed", sep="\t")
    assert list(cl.columns) == list(bed.columns) + ["cluster", "n_reads"] and len(cl) == len(bed)
    assert cl["cluster"].dtype == float and (cl.groupby("qname")["cluster"].nunique() == 1).all()
    per_read = cl.drop_duplicates("qname")
    assert (per_read.groupby("cluster").size() == per_read.groupby("cluster")["n_reads"].first()).all()
    assert os.path.exists(base + ".mappings.representative.bed")


@pytest.mark.parametrize("level,block,force_host_walk", [(0, 60000, False), (1, 5000, False), (9, 65280, False), (6, 60000, True)])
def test_device_inflate_and_record_finder(tmp_path, monkeypatch, level, block, force_host_walk):
    """Stored / fast / best DEFLATE blocks of several sizes through the device decoder and the device record-boundary
    finder give the table the host-inflated path gives; the host-walk fallback of a failed chain check is forced once."""
    from fslr_b200 import mapping_info as mi, synth_bam as sb
    refs, recs, primers = sb.make_alignments(2500, seed=50 + level, p_unmapped=0.2)
    raw = sb.bam_bytes(refs, [sb.encode_record(*r) for r in recs], block=block, level=level)
    a = mi.read_bam_table(raw, None, primers, device_inflate=False)
    want = a.mappings_bed_bytes("9.9").tobytes()
    names = list(a.qnames())
    a.close()
    if force_host_walk:
        monkeypatch.setenv("FSLRC_BAM_FORCE_HOST_WALK", "1")
    b = mi.read_bam_table(raw, None, primers, device_inflate=True)
    try:
        assert int(b.info.reserved) == int(force_host_walk)
        assert (b.n_records, b.n_mapped) == (len(recs), sum(1 for r in recs if not r[1] & 4))
        got = b.mappings_bed_bytes("9.9").tobytes()
        assert got == want, _first_diff(got, want)
        assert list(b.qnames()) == names
    finally:
        b.close()


def test_corrupt_bgzf_payload_is_reported():
    from fslr_b200 import _native, mapping_info as mi, synth_bam as sb
    refs, recs, primers = sb.make_alignments(200, seed=60)
    raw = bytearray(sb.bam_bytes(refs, [sb.encode_record(*r) for r in recs], level=6))
    for k in range(len(raw) // 2, len(raw) // 2 + 40):
        raw[k] ^= 0x5a
    with pytest.raises((_native.FslrError, ValueError)):
        mi.read_bam_table(bytes(raw), None, primers, device_inflate=True)

"""CPU tests of the table-producer row: the oracle restatement of collect_mapping_info.mapping_info
against the fixtures the unmodified reference produced (tests/golden/make_bam_golden.py), and the host-side BGZF /
header readers of fslr_b200.mapping_info."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import mapping_info_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES_DIR = os.path.join(ROOT, "tests", "golden", "bam_cases")
CASES = json.load(open(os.path.join(CASES_DIR, "cases.json")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_fixture(name):
    c = CASES[name]
    bam = os.path.join(CASES_DIR, name + ".bam")
    regions = mo.read_regions(os.path.join(CASES_DIR, name + ".regions.bed")) if c["regions"] else None
    if c["reference_error"]:
        # err_empty: the reference's KeyError on an empty frame; err_bad_refid: the reader's IndexError on a reference id
        # that is not in the header
        with pytest.raises((mo.MappingInfoError, KeyError, IndexError)):
            rows = mo.mapping_rows(bam, regions, c["primers"], c["fslr_version"])
            if not rows:
                raise KeyError("empty table")
        return
    rows = mo.mapping_rows(bam, regions, c["primers"], c["fslr_version"])
    want = open(os.path.join(CASES_DIR, name + ".mappings.bed")).read()
    assert mo.mapping_tsv(rows, c["regions"]) == want


def test_bgzf_inflate_and_header():
    from fslr_b200 import mapping_info as mi, synth_bam as sb
    refs, recs, _ = sb.make_alignments(60, seed=11)
    b = sb.bam_bytes(refs, [sb.encode_record(*r) for r in recs], block=777)       # many small blocks, one EOF block
    u = mi.inflate_bgzf(b, threads=4)
    assert bytes(u) == gzip.decompress(b)
    got, first = mi.parse_bam_header(u)
    assert got == list(refs)
    assert int(np.frombuffer(u[first:first + 4].tobytes(), dtype="<i4")[0]) >= 32     # block_size of the first record
    with pytest.raises(ValueError):
        mi.inflate_bgzf(b"not a bam file at all, not even gzip")
    with pytest.raises(ValueError):
        mi.inflate_bgzf(b[:len(b) // 2])


def test_reference_end_of_an_alignment_without_reference_bases():
    """htslib's bam_endpos counts a zero reference span as 1 (pysam reference_end); the stub, the oracle and the kernel agree."""
    rows = mo.mapping_rows(os.path.join(CASES_DIR, "edge.bam"), None, CASES["edge"]["primers"])
    r = [x for x in rows if x["qname"].startswith("cccc") and not x["inferred_by_primer"]][0]
    assert (r["rstart"], r["rend"]) == (78, 78)


@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_oracle_against_the_reference_itself_on_random_bams(tmp_path, seed):
    """Only where /root/reference exists (the build container): fresh random BAMs through the unmodified
    collect_mapping_info.mapping_info and through the restatement must give the same file."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference not present")
    from fslr_b200 import synth_bam as sb
    kw = [dict(), dict(name_style="prefix", with_seq_on_supp=True), dict(p_single=0.7, p_false=0.5), dict(max_aln=12, p_two_primary=0.3)][seed % 4]
    refs, recs, primers = sb.make_alignments(150, seed=seed, **kw)
    bam = str(tmp_path / "t.bam")
    sb.write_bam(bam, refs, [sb.encode_record(*r) for r in recs])
    reg = None
    if seed % 2:
        reg = str(tmp_path / "r.bed")
        open(reg, "w").write("chr1\t1000\t90000000\nchr21\t5\t20000000\nL1_TALEN\t100\t4000\n")
    out = str(tmp_path / "ref.bed")
    assert rh.run_reference_mapping_info(bam, out, reg, primers) is None
    rows = mo.mapping_rows(bam, mo.read_regions(reg) if reg else None, primers, rh.FSLR_VERSION_FOR_TESTS)
    assert mo.mapping_tsv(rows, bool(reg)) == open(out).read()

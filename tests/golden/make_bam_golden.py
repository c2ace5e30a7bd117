"""Generates tests/golden/bam_cases/: small synthetic BAM files plus the `mappings.bed` the UNMODIFIED reference
(collect_mapping_info.mapping_info, run through oracle/ref_harness.py with the stub pysam) writes for each, and the
inputs it stops on.  Run in the build container (needs /root/reference):  python tests/golden/make_bam_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fslr_b200 import synth_bam as sb          # noqa: E402
from oracle import ref_harness as rh           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bam_cases")
REGIONS = "chr1\t1000\t90000000\nchr21\t5\t20000000\nL1_TALEN\t100\t4000\nchr9\t1\t2\n"


def edge_records():
    """Hand-made records for the branches random data rarely reaches."""
    M, I, S, H = 0, 1, 4, 5
    seq = "ACGTTGCANN" * 12                                                  # 120 nt
    r = []
    # two would-be primaries with the same AS: the first wins; records of the read interleaved with another read's
    r.append(("bbbb.1.21q1F_17p6R", 0, 0, 1000, 60, [(M, 60), (S, 60)], seq, [("AS", "C", 50)]))
    r.append(("aaaa.9.False_False", 16, 1, 500, 3, [(S, 20), (M, 100)], seq, [("AS", "c", -7)]))
    r.append(("bbbb.1.21q1F_17p6R", 16, 2, 2000, 60, [(M, 60), (S, 60)], sb._revcomp(seq), [("AS", "s", 50)]))
    r.append(("bbbb.1.21q1F_17p6R", 2048, 3, 3000, 0, [(H, 30), (M, 40), (I, 5), (M, 15), (H, 30)], "", [("AS", "S", 40000)]))
    r.append(("bbbb.1.21q1F_17p6R", 2064, 3, 3000, 0, [(H, 30), (M, 60), (H, 30)], "", [("AS", "i", -40000)]))   # same qstart after the flip
    # no reference-consuming operation: htslib's end = pos + 1
    r.append(("cccc.x.y.21q1R_False", 0, 4, 77, 9, [(S, 10), (I, 100), (S, 10)], seq, [("AS", "I", 12)]))
    # single alignment, primer 2 inferred, read shorter than the primer: negative qstart
    r.append(("dddd.False_16p1R", 0, 5, 10, 1, [(M, 12)], "ACGTACGTACGT", [("XX", "Z", "AS"), ("AS", "C", 1)]))
    # single alignment, gaps at both ends > 5: left alone; name without '.'
    r.append(("eeee_17p6F", 0, 0, 5000, 30, [(S, 6), (M, 100), (S, 14)], seq, [("AS", "C", 9)]))
    # single alignment, gap exactly 5 at one end
    r.append(("ffff.21q1F_17p6F", 16, 0, 6000, 30, [(S, 5), (M, 100), (S, 15)], seq, [("AS", "C", 9)]))
    # primer token with several trailing F/R characters
    r.append(("gggg.16p1FRRF_False", 0, 0, 7000, 30, [(M, 120)], seq, [("ZB", "B", ("C", [1, 2, 3])), ("AS", "C", 9)]))
    # unmapped record that shares a name with a mapped one
    r.append(("gggg.16p1FRRF_False", 4, -1, -1, 0, [], "ACGT", []))
    # aln_size straddling the 50 bp anchor limit (49 / 50) at either end
    r.append(("hhhh.False_False", 0, 0, 8000, 30, [(M, 49), (S, 171)], seq + seq[:100], [("AS", "C", 9)]))
    r.append(("hhhh.False_False", 2048, 1, 8000, 30, [(H, 49), (M, 121), (H, 50)], "", [("AS", "C", 9)]))
    r.append(("hhhh.False_False", 2048, 2, 8000, 30, [(H, 170), (M, 50)], "", [("AS", "C", 9)]))
    r.append(("iiii.False_False", 0, 0, 8000, 30, [(M, 50), (S, 170)], seq + seq[:100], [("AS", "C", 9)]))
    r.append(("iiii.False_False", 2048, 2, 8000, 30, [(H, 170), (M, 50)], "", [("AS", "C", 9)]))
    return r


def error_cases():
    M, S = 0, 4
    seq = "ACGT" * 10
    ok = ("ok.1.False_False", 0, 0, 100, 60, [(M, 40)], seq, [("AS", "C", 5)])
    return {
        "err_no_as": [ok, ("r.1.False_False", 0, 0, 100, 60, [(M, 40)], seq, [("NM", "C", 5)])],
        "err_no_primary": [ok, ("r.1.False_False", 2048, 0, 100, 60, [(M, 40)], seq, [("AS", "C", 5)])],
        "err_no_seq": [ok, ("r.1.False_False", 0, 0, 100, 60, [(M, 40)], "", [("AS", "C", 5)])],
        "err_bad_name": [ok, ("r.1.A_B_C", 0, 0, 100, 60, [(S, 10), (M, 20), (S, 10)], seq, [("AS", "C", 5)])],
        "err_unknown_primer": [ok, ("r.1.9z9F_False", 0, 0, 100, 60, [(M, 40)], seq, [("AS", "C", 5)])],
        "err_no_cigar": [ok, ("r.1.False_False", 0, 0, 100, 60, [], seq, [("AS", "C", 5)])],
        "err_empty": [("r.1.False_False", 4, -1, -1, 0, [], seq, [])],
        "err_bad_refid": [ok, ("r.1.False_False", 0, 1000, 100, 60, [(M, 40)], seq, [("AS", "C", 5)])],   # mapped, but no such reference
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {}
    specs = {
        "basic": dict(n_reads=110, seed=1),
        "regions_float": dict(n_reads=90, seed=2, regions=True),
        "regions_int": dict(n_reads=80, seed=3, regions=True, p_single=0.0),
        "prefix_names": dict(n_reads=90, seed=4, name_style="prefix", with_seq_on_supp=True),
        "mostly_single": dict(n_reads=150, seed=5, p_single=0.85),
        "one_read": dict(n_reads=1, seed=6, p_unmapped=0.0),
    }
    for name, kw in specs.items():
        regions = kw.pop("regions", False)
        refs, recs, primers = sb.make_alignments(**kw)
        cases[name] = (refs, recs, primers, regions)
    cases["edge"] = (sb.DEFAULT_REFS, edge_records(), sb.DEFAULT_PRIMERS, False)
    cases["edge_regions"] = (sb.DEFAULT_REFS, edge_records(), sb.DEFAULT_PRIMERS, True)
    for name, recs in error_cases().items():
        cases[name] = (sb.DEFAULT_REFS, recs, sb.DEFAULT_PRIMERS, False)
    index = {}
    for name, (refs, recs, primers, regions) in cases.items():
        bam = os.path.join(OUT, name + ".bam")
        sb.write_bam(bam, refs, [sb.encode_record(*r) for r in recs], level=9)
        reg = None
        if regions:
            reg = os.path.join(OUT, name + ".regions.bed")
            open(reg, "w").write(REGIONS)
        bed = os.path.join(OUT, name + ".mappings.bed")
        e = rh.run_reference_mapping_info(bam, bed, reg, primers)
        if e is not None and os.path.exists(bed):
            os.remove(bed)
        index[name] = {"regions": bool(regions), "primers": primers, "fslr_version": rh.FSLR_VERSION_FOR_TESTS,
                       "reference_error": None if e is None else type(e).__name__}
        print(name, index[name]["reference_error"], os.path.getsize(bam), os.path.getsize(bed) if e is None else "-")
    json.dump(index, open(os.path.join(OUT, "cases.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY — CPU restatement of `collect_mapping_info.mapping_info`
(/root/reference/fslr/collect_mapping_info.py:7-181), the producer of `<base>.mappings.bed`.

Plain Python over the records the stub BAM reader (oracle/stubs/pysam) yields; no pandas on the ordering path, so that
the two sorts of the reference (collect_mapping_info.py:163,174 — pandas multi-key `sort_values`, a stable lexsort) are
restated as explicit stable sorts.  Pinned against the reference itself on the fixtures of tests/golden/bam_cases
(tests/test_mapping_info_oracle.py).  The product path (fslr_b200/) never imports this module.
"""
import os
import sys

_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")

COLUMNS = ["chrom", "rstart", "rend", "qname", "n_alignments", "aln_size", "qstart", "qend", "strand", "mapq", "qlen",
           "alignment_score", "short_anchor<50bp", "fslr_version", "inferred_by_primer", "seq"]      # :176-177


class MappingInfoError(Exception):
    """The input makes the reference stop (`quit()` at :46-47,101-103) or raise (missing AS tag, no CIGAR, bad name)."""


def _pysam():
    if _STUBS not in sys.path:
        sys.path.insert(0, _STUBS)
    import pysam
    return pysam


def query_pos(a):
    """get_query_pos_from_cigartuples, :7-16"""
    ct = a.cigartuples
    if not ct:
        raise MappingInfoError("record without CIGAR")
    qlen = a.infer_read_length()
    start, end = 0, qlen
    if ct[0][0] in (4, 5):
        start += ct[0][1]
    if ct[-1][0] in (4, 5):
        end -= ct[-1][1]
    return start, end, qlen


def mapping_rows(bam_path, regions=None, primers=None, fslr_version="0.0.test"):
    """Returns the rows of the output table, in output order, as dicts keyed like the reference's columns.
    regions: {chrom: [(start, end)]} or None (the parsed regions file, :28-36); primers: {name: sequence}."""
    af = _pysam().AlignmentFile(bam_path, "r")
    groups = {}                                                   # dict keeps first-appearance order, :23-26
    for a in af.fetch(until_eof=True):
        if not a.flag & 4:
            groups.setdefault(a.qname, []).append(a)
    primers = primers or {}
    res = []
    for qname, v in groups.items():
        cand = [i for i, a in enumerate(v) if not a.flag & 2304]  # :42
        if len(cand) > 1:                                         # :43-44 first record with the highest AS
            best = cand[0]
            for i in cand[1:]:
                if _tag(v[i], "AS") > _tag(v[best], "AS"):
                    best = i
            cand = [best]
        if len(cand) != 1:                                        # :46-48
            raise MappingInfoError("no primary record for %s" % qname)
        pri = cand[0]
        pri_rev = bool(v[pri].flag & 16)
        seq = v[pri].get_forward_sequence()
        if not seq:                                               # :101-103
            raise MappingInfoError("primary record of %s has no sequence" % qname)
        temp = []
        for i, a in enumerate(v):
            qs, qe, qlen = query_pos(a)
            rev = bool(a.flag & 16)
            if rev != pri_rev:                                    # :59-62
                st = qlen - qe
                qe = st + qe - qs
                qs = st
            chrom = af.get_reference_name(a.rname)
            start, end = a.reference_start + 1, a.reference_end   # :70-71
            row = {"qname": qname, "n_alignments": len(v), "chrom": chrom, "rstart": start, "rend": end,
                   "strand": "-" if rev else "+", "qstart": qs, "qend": qe, "qlen": qlen, "aln_size": qe - qs,
                   "mapq": a.mapq, "alignment_score": _tag(a, "AS"), "seq": seq if i == pri else "",
                   "fslr_version": fslr_version, "inferred_by_primer": 0}
            if regions:                                           # :72-76,96-97: (start, end] against (s, e]
                row["overlaps_region"] = int(any(start < e and s < end for s, e in regions.get(chrom, ())))
            temp.append(row)
        if len(temp) == 1:                                        # :109-158
            t0 = temp[0]
            names = qname.split(".")[-1].split("_")
            if len(names) != 2:
                raise MappingInfoError("read name %s does not end in <primer>_<primer>" % qname)
            p1, p2 = (x.rstrip("FR") for x in names)
            if not (t0["qstart"] > 5 and t0["qlen"] - t0["qend"] > 5):
                if p1 != "False":
                    if p1 not in primers:
                        raise MappingInfoError("unknown primer %s" % p1)
                    t0["n_alignments"] = 2
                    temp = [_inferred(t0, p1, names[0], 0, len(primers[p1]), fslr_version), t0]
                elif p2 != "False":
                    if p2 not in primers:
                        raise MappingInfoError("unknown primer %s" % p2)
                    t0["n_alignments"] = 2
                    temp = [t0, _inferred(t0, p2, names[1], t0["qlen"] - len(primers[p2]), t0["qlen"], fslr_version)]
        res += temp
    # :163 stable sort by (qname, qstart); :165-172 anchors; :174 stable sort by (n_alignments desc, qname, qstart)
    res.sort(key=lambda r: (r["qname"], r["qstart"]))
    i = 0
    while i < len(res):
        j = i
        while j < len(res) and res[j]["qname"] == res[i]["qname"]:
            j += 1
        bad = int(res[i]["aln_size"] < 50 or res[j - 1]["aln_size"] < 50)
        for k in range(i, j):
            res[k]["short_anchor<50bp"] = bad
        i = j
    res.sort(key=lambda r: (-r["n_alignments"], r["qname"], r["qstart"]))
    return res


def _tag(a, t):
    try:
        return a.get_tag(t)
    except KeyError:
        raise MappingInfoError("record of %s lacks the %s tag" % (a.qname, t))


def _inferred(t0, primer, token, qstart, qend, fslr_version):
    """:123-139 / :142-157"""
    return {"qname": t0["qname"], "n_alignments": 2, "chrom": primer, "rstart": 0, "rend": 0,
            "strand": "-" if token[-1] == "R" else "+", "qstart": qstart, "qend": qend, "qlen": t0["qlen"], "aln_size": 0,
            "mapq": 0, "alignment_score": 0, "seq": "", "fslr_version": fslr_version, "inferred_by_primer": 1}


def mapping_tsv(rows, with_regions=False):
    """The file `df.to_csv(outf, index=False, sep='\\t')` writes (:178-181).  With a regions file the inferred rows have no
    `overlaps_region` entry (:123-139 never set it), so pandas holds that column as float: "1.0" / "0.0" / ""."""
    cols = COLUMNS + (["overlaps_region"] if with_regions else [])
    as_float = with_regions and any("overlaps_region" not in r for r in rows)
    out = ["\t".join(cols)]
    for r in rows:
        f = []
        for c in cols:
            if c == "overlaps_region":
                v = r.get(c)
                f.append("" if v is None else ("%d.0" % v if as_float else "%d" % v))
            else:
                f.append(str(r[c]))
        out.append("\t".join(f))
    return "\n".join(out) + "\n"


def read_regions(path):
    """:28-36"""
    regions = {}
    with open(path) as f:
        for line in f:
            l = line.strip().split("\t")
            regions.setdefault(l[0], []).append((int(l[1]), int(l[2])))
    return regions

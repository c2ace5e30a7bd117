"""Synthetic BAM files for the table-producer row: the input `collect_mapping_info.mapping_info`
(/root/reference/fslr/collect_mapping_info.py:19-26) reads with pysam.  Pure numpy/zlib, no htslib.

`write_bam(path, refs, records)` writes BGZF-compressed BAM; `make_alignments(...)` draws concatemer reads the way
SURVEY §8d describes them (breads at a primer locus, fillings elsewhere), as one primary record (soft clips, full
sequence) plus supplementary records (hard clips, no sequence), with the corner cases mapping_info branches on:
single-alignment reads with and without recognised primers, reverse-strand primaries, unmapped records, secondary
records, several would-be primaries.
"""
import struct
import zlib

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(chunk, level=1):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = c.compress(chunk) + c.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(data) + 25) + data +
            struct.pack("<II", zlib.crc32(chunk), len(chunk)))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def encode_record(qname, flag, ref_id, pos, mapq, cigar, seq, tags):
    """One BAM alignment record (with its block_size prefix).  cigar: [(op, len)], seq: str ('' = absent),
    tags: [(two-letter tag, type char, value)] with integer types cCsSiI, 'Z' strings, 'A' chars, 'f' floats."""
    name = qname.encode() + b"\x00"
    ref_len = sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))
    l_seq = len(seq)
    body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(name), mapq, reg2bin(max(pos, 0), max(pos, 0) + max(ref_len, 1)),
                       len(cigar), flag, l_seq, -1, -1, 0)
    body += name
    body += b"".join(struct.pack("<I", (l << 4) | op) for op, l in cigar)
    codes = [_SEQ_CODE.get(c, 15) for c in seq] + [0]
    body += bytes((codes[i] << 4) | codes[i + 1] for i in range(0, l_seq, 2))
    body += b"\xff" * l_seq
    for tag, typ, val in tags:
        body += tag.encode() + typ.encode()
        if typ in "cCsSiIf":
            body += struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[typ], val)
        elif typ == "A":
            body += val.encode()
        elif typ == "Z":
            body += val.encode() + b"\x00"
        elif typ == "B":                                          # (subtype char, list)
            sub, arr = val
            body += sub.encode() + struct.pack("<i", len(arr))
            body += b"".join(struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub], v) for v in arr)
        else:
            raise ValueError(typ)
    return struct.pack("<i", len(body)) + body


def bam_bytes(refs, records, header_text=None, block=60000, level=1):
    """refs: [(name, length)]; records: iterable of encode_record(...) byte strings.  Returns the BGZF file bytes."""
    text = header_text if header_text is not None else (
        "@HD\tVN:1.6\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in refs)).encode()
    raw = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for n, l in refs:
        raw += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    for r in records:
        raw += r
    raw = bytes(raw)
    return b"".join(_bgzf_block(raw[i:i + block], level) for i in range(0, len(raw), block)) + BGZF_EOF


def write_bam(path, refs, records, **kw):
    with open(path, "wb") as f:
        f.write(bam_bytes(refs, records, **kw))


DEFAULT_REFS = [("chr1", 248387328), ("chr2", 242696752), ("chr17", 84276897), ("chr21", 45090682), ("chrX", 154259566),
                ("L1_TALEN", 8000)]
DEFAULT_PRIMERS = {"21q1": "ACGTACGTAGCTAGCTAGGATCGATCG", "17p6": "TTGACCATGACCATGGACCA", "16p1": "GGATCCGATCGATTAGC"}


def make_alignments(n_reads, seed=0, refs=DEFAULT_REFS, primers=DEFAULT_PRIMERS, max_aln=6, p_single=0.25, p_unmapped=0.05,
                    p_secondary=0.05, p_two_primary=0.03, p_false=0.2, name_style="uuid", with_seq_on_supp=False):
    """Returns (refs, records as argument tuples of encode_record, primers).  Records of a read are emitted together
    (aligner output order) but in a random order within the read, reads in random order."""
    rng = np.random.default_rng(seed)
    pnames = list(primers)
    out = []
    for i in range(n_reads):
        na = 1 if rng.random() < p_single else int(rng.integers(2, max_aln + 1))
        seg = rng.integers(30, 400, size=na)
        gaps = rng.integers(0, 12, size=na + 1)
        if na == 1:                                               # leave room for / against the "gap at both ends" rule
            gaps = rng.choice([0, 3, 5, 6, 40], size=2)
        qlen = int(seg.sum() + gaps.sum())
        toks = []
        for _ in range(2):
            toks.append("False" if rng.random() < p_false else pnames[int(rng.integers(len(pnames)))] + "FR"[int(rng.integers(2))])
        if name_style == "uuid":
            qname = "%08x-%04x.%d_%d.%s_%s" % (int(rng.integers(1 << 32)), i & 0xffff, int(rng.integers(1000)), int(rng.integers(1000)), toks[0], toks[1])
        else:                                                     # long shared prefixes: the string sort needs every chunk
            qname = "read_with_a_long_common_prefix_%07d.%s_%s" % (int(rng.integers(10 ** 7)) if rng.random() < 0.5 else i, toks[0], toks[1])
        bases = "".join("ACGTN"[int(b)] for b in rng.choice(5, size=qlen, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        pri = int(rng.integers(na))
        pri_rev = bool(rng.random() < 0.4)
        recs = []
        qs = int(gaps[0])
        two_pri = na > 1 and rng.random() < p_two_primary
        for k in range(na):
            qe = qs + int(seg[k])
            rev = bool(rng.random() < 0.5) if k != pri else pri_rev
            ref = int(rng.integers(len(refs)))
            rl = refs[ref][1]
            pos = int(rng.integers(0, max(1, rl - 2000)))
            # cigar over the aligned part: M with an occasional I / D / = / X / N
            ops = []
            left = int(seg[k])
            while left > 0:
                m = int(min(left, rng.integers(5, 200)))
                ops.append((int(rng.choice([0, 0, 0, 7, 8])), m))
                left -= m
                if left > 3 and rng.random() < 0.5:
                    kind = int(rng.choice([1, 2, 3]))
                    if kind == 1:
                        ins = int(min(left - 1, rng.integers(1, 4)))
                        ops.append((1, ins)); left -= ins
                    else:
                        ops.append((kind, int(rng.integers(1, 30))))
            a, b = qs, qlen - qe                                  # clipped bases before / after on the read's own strand
            if rev != pri_rev:
                pass
            # clips are stored in the orientation of the alignment: reverse-strand records see the read reversed
            lead, trail = (b, a) if rev else (a, b)
            is_pri = k == pri
            clip = 4 if (is_pri or with_seq_on_supp) else 5
            cig = ([(clip, lead)] if lead else []) + ops + ([(clip, trail)] if trail else [])
            flag = (16 if rev else 0) | (0 if (is_pri or (two_pri and k == (pri + 1) % na)) else 2048)
            if is_pri or (two_pri and k == (pri + 1) % na):
                # the primary carries the whole read in alignment orientation
                if clip == 5:                                     # a second "primary" written with hard clips keeps a partial seq
                    s = bases[qs:qe]
                else:
                    s = bases
                s_al = _revcomp(s) if rev else s
            else:
                s_al = (_revcomp(bases) if rev else bases) if with_seq_on_supp else ""
            astype = "CcSsIi"[int(rng.integers(6))]
            asv = int(rng.integers(0, 120)) if astype in "Cc" else int(rng.integers(0, 30000))
            tags = [("NM", "C", int(rng.integers(0, 20)))]
            if rng.random() < 0.5:
                tags.append(("SA", "Z", "chr1,100,+,10M,60,0;"))
            if rng.random() < 0.2:
                tags.append(("ZB", "B", ("s", [1, -2, 3])))
            tags.append(("AS", astype, asv))
            if rng.random() < 0.5:
                tags.append(("XS", "i", -5))
            recs.append((qname, flag, ref, pos, int(rng.integers(0, 61)), cig, s_al, tags))
            qs = qe + int(gaps[k + 1])
        if rng.random() < p_secondary:                            # a secondary copy (flag 256) of some alignment, no seq
            q = recs[int(rng.integers(len(recs)))]
            recs.append((q[0], (q[1] & 16) | 256, q[2], q[3], 0, q[5], "", [("AS", "i", 1)]))
        if rng.random() < p_unmapped:                             # an unmapped record of another read
            out.append([("%08x-unmapped.%d.False_False" % (int(rng.integers(1 << 32)), i), 4, -1, -1, 0, [], "ACGT", [])])
        order = rng.permutation(len(recs))
        out.append([recs[j] for j in order])
    order = rng.permutation(len(out))
    flat = [r for j in order for r in out[j]]
    return list(refs), flat, dict(primers)


def _revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgtNnXx", "TGCAtgcaNnXx"))

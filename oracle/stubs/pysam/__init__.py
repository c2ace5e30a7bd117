"""TEST INFRASTRUCTURE ONLY — stand-in for `pysam` so that the unmodified reference modules can be imported and run
in the build container: /root/reference/fslr/cluster.py:4 (only `get_chromosome_lengths`, cluster.py:173-175) and
/root/reference/fslr/collect_mapping_info.py:1,22-26 (`AlignmentFile(f, 'r').fetch(until_eof=True)` and the
AlignedSegment attributes mapping_info touches).  A small pure-Python BAM reader with pysam's documented semantics
for exactly those attributes; nothing on the product path imports it."""
import gzip
import struct

_SEQ = "=ACMGRSVTWYHKDBN"
_COMP = str.maketrans("ACGTacgtNnXx", "TGCAtgcaNnXx")           # pysam.AlignedSegment.get_forward_sequence


class AlignedSegment:
    __slots__ = ("qname", "flag", "rname", "reference_start", "mapq", "cigartuples", "seq", "_tags")

    @property
    def query_name(self):
        return self.qname

    @property
    def reference_id(self):
        return self.rname

    @property
    def mapping_quality(self):
        return self.mapq

    @property
    def reference_end(self):
        if not self.cigartuples or self.flag & 4:
            return None
        rlen = sum(l for op, l in self.cigartuples if op in (0, 2, 3, 7, 8))
        return self.reference_start + (rlen or 1)               # htslib bam_endpos: a zero reference span counts as 1

    def infer_read_length(self):
        if not self.cigartuples:
            return None
        return sum(l for op, l in self.cigartuples if op in (0, 1, 4, 5, 7, 8))

    def infer_query_length(self):
        if not self.cigartuples:
            return None
        return sum(l for op, l in self.cigartuples if op in (0, 1, 4, 7, 8))

    def get_tag(self, tag):
        return self._tags[tag]                                   # KeyError when absent, like pysam

    def get_forward_sequence(self):
        if self.seq is None:
            return None
        return self.seq[::-1].translate(_COMP) if self.flag & 16 else self.seq


def _parse_tags(b):
    tags, p = {}, 0
    size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}
    while p + 3 <= len(b):
        tag, typ = b[p:p + 2].decode(), chr(b[p + 2])
        p += 3
        if typ == "A":
            tags[tag] = chr(b[p]); p += 1
        elif typ in fmt:
            tags[tag] = struct.unpack_from("<" + fmt[typ], b, p)[0]; p += size[typ]
        elif typ in "ZH":
            e = b.index(b"\x00", p)
            tags[tag] = b[p:e].decode(); p = e + 1
        elif typ == "B":
            sub = chr(b[p]); n = struct.unpack_from("<i", b, p + 1)[0]
            tags[tag] = list(struct.unpack_from("<%d%s" % (n, fmt[sub]), b, p + 5)); p += 5 + n * size[sub]
        else:
            raise ValueError("bad aux type %r" % typ)
    return tags


class AlignmentFile:
    def __init__(self, path, mode="r", **kw):
        with gzip.open(path, "rb") as f:                         # BGZF = concatenated gzip members
            raw = f.read()
        if raw[:4] != b"BAM\x01":
            raise ValueError("pysam stub: %s is not a BAM file" % path)
        l_text = struct.unpack_from("<i", raw, 4)[0]
        p = 8 + l_text
        n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l = struct.unpack_from("<i", raw, p)[0]
            self.references.append(raw[p + 4:p + 4 + l - 1].decode())
            self.lengths.append(struct.unpack_from("<i", raw, p + 4 + l)[0])
            p += 8 + l
        self._raw, self._start = raw, p

    def get_reference_name(self, rid):
        if rid < 0:
            return None
        return self.references[rid]

    def get_reference_length(self, name):
        return self.lengths[self.references.index(name)]

    def fetch(self, until_eof=False, **kw):
        raw, p = self._raw, self._start
        while p + 4 <= len(raw):
            bs = struct.unpack_from("<i", raw, p)[0]
            ref, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack_from("<iiBBHHHiiii", raw, p + 4)
            q = p + 36
            a = AlignedSegment()
            a.qname = raw[q:q + l_name - 1].decode(); q += l_name
            cig = struct.unpack_from("<%dI" % n_cig, raw, q); q += 4 * n_cig
            a.cigartuples = [(c & 15, c >> 4) for c in cig] if n_cig else None
            sb = raw[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2 + l_seq
            a.seq = "".join(_SEQ[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)) if l_seq else None
            a.flag, a.rname, a.reference_start, a.mapq = flag, ref, pos, mapq
            a._tags = _parse_tags(raw[q:p + 4 + bs])
            yield a
            p += 4 + bs

    def close(self):
        pass

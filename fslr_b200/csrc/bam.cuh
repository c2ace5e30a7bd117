// The producer of `<base>.mappings.bed` on the GPU: collect_mapping_info.mapping_info
// (/root/reference/fslr/collect_mapping_info.py:7-181) over the uncompressed BAM record stream.
//
//   k_bam_parse     one thread per mapped record: fixed fields, CIGAR walk (query interval :7-16, reference end,
//                   read length with hard clips), the AS tag out of the aux block, FNV-1a of the read name
//   (intern)        the name -> dense read id kernels of tsv.cuh: ids in order of first appearance = the dict order of :23-26
//   k_bam_bounds    records grouped by read id (one stable radix sort), file order inside a read
//   k_bam_reads     one thread per read: the primary record (:42-50), n_alignments, the missing-bread rule for
//                   single-alignment reads (:109-158: primer tokens of the read name, gap test, inferred primer row)
//   k_bam_rows      rows in the reference's `res` order: strand flip of the query interval (:59-62), region overlap (:70-76)
//   k_bam_namekey   4 name bytes per pass of an LSD string sort (qname ascending, :163,174)
//   k_bam_anchor / k_bam_final   short_anchor<50bp (:165-172) and the gather into output order (:174)
//   k_bam_outlen / k_bam_emit    the TSV pandas writes (:176-181), sequence of the primary record decoded from 4-bit codes
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tsv.cuh"

namespace bam {

enum { BE_TRUNC = 1, BE_NOCIGAR = 2, BE_COLLISION = 4 /* = tsv::TE_COLLISION */, BE_NOAS = 8, BE_NOPRIMARY = 16, BE_NOSEQ = 32,
       BE_NAME = 64, BE_PRIMER = 128, BE_AUX = 256, BE_RANGE = 512 };
constexpr int MAX_PRIMERS = 64;

struct Primers {                 // by value in kernel parameters: names packed back to back
    int n;
    int off[MAX_PRIMERS + 1];
    int len[MAX_PRIMERS];        // length of the primer SEQUENCE (:133,151)
    char names[1024];
};

struct Recs {                    // per mapped record, file order
    int *flag, *ref, *pos1, *rend, *mapq, *qs, *qe, *qlen, *as, *lseq, *nlen;
    long long *seq_off, *noff;
    unsigned long long *hash;
};
struct Rows {                    // per table row
    int *rid, *chrom, *rstart, *rend, *naln, *aln, *qstart, *qend, *strand, *mapq, *qlen, *as, *inferred, *overlaps, *seq_rec;
};

__device__ __forceinline__ int rd32(const unsigned char *p) { return (int)((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24)); }
__device__ __forceinline__ int rd16(const unsigned char *p) { return (int)((unsigned)p[0] | ((unsigned)p[1] << 8)); }
__device__ __forceinline__ int aux_size(unsigned char t) {
    switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; default: return 0; }
}

__global__ void k_bam_parse(const unsigned char *__restrict__ text, long long n_bytes, const long long *__restrict__ rec_off, int M,
                            int n_ref, unsigned long long seed, Recs R, int *maxlen, int *err) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    int nl = 0, bad = 0;
    if (m < M) {
        const long long o = rec_off[m];
        const unsigned char *p = text + o;
        const long long end = o + 4 + (long long)rd32(p);
        const int l_name = p[12], n_cig = rd16(p + 16), l_seq = rd32(p + 20);
        const long long c0 = o + 36 + l_name, s0 = c0 + 4LL * n_cig, a0 = s0 + ((long long)l_seq + 1) / 2 + (long long)l_seq;
        if (end > n_bytes || a0 > end || l_name < 1 || l_seq < 0) bad |= BE_TRUNC;
        R.flag[m] = rd16(p + 18); R.ref[m] = rd32(p + 4); R.mapq[m] = p[13];
        // a mapped record (no flag 0x4) must name a reference of the header: its id indexes the chromosome tables downstream
        if (!(rd16(p + 18) & 4) && (unsigned)rd32(p + 4) >= (unsigned)n_ref) bad |= BE_RANGE;
        const int pos = rd32(p + 8);
        R.pos1[m] = pos + 1;
        nl = l_name - 1;
        R.noff[m] = o + 36; R.nlen[m] = nl;
        R.hash[m] = (bad & BE_TRUNC) ? 0ull : tsv::fnv1a(text + o + 36, nl, seed);
        R.seq_off[m] = s0; R.lseq[m] = l_seq;
        long long rlen = 0, qlen = 0; int lead = 0, trail = 0, as = 0;
        if (!(bad & BE_TRUNC)) {
            if (n_cig == 0) bad |= BE_NOCIGAR;                          // cigartuples is None: TypeError at :12
            for (int k = 0; k < n_cig; k++) {
                const unsigned v = (unsigned)rd32(text + c0 + 4LL * k);
                const int op = v & 15u; const int len = (int)(v >> 4);
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += len;
                if (op == 0 || op == 1 || op == 4 || op == 5 || op == 7 || op == 8) qlen += len;   // infer_read_length: hard clips count
                if (k == 0 && (op == 4 || op == 5)) lead = len;
                if (k == n_cig - 1 && (op == 4 || op == 5)) trail = len;
            }
            if (rlen == 0) rlen = 1;                                    // htslib bam_endpos
            if (qlen > 0x7fffffffLL || pos + rlen > 0x7fffffffLL) { bad |= BE_RANGE; qlen = 0; rlen = 1; }
            // aux block: TAG(2) TYPE(1) VALUE
            long long a = a0; bool found = false;
            while (a + 3 <= end) {
                const unsigned char t0 = text[a], t1 = text[a + 1], ty = text[a + 2];
                a += 3;
                const bool is_as = t0 == 'A' && t1 == 'S' && !found;
                const int sz = aux_size(ty);
                if (sz) {
                    if (a + sz > end) { bad |= BE_AUX; break; }
                    if (is_as) {
                        found = true;
                        if (ty == 'c') as = (int)(signed char)text[a];
                        else if (ty == 'C') as = text[a];
                        else if (ty == 's') as = (int)(short)rd16(text + a);
                        else if (ty == 'S') as = rd16(text + a);
                        else if (ty == 'i') as = rd32(text + a);
                        else if (ty == 'I') { as = rd32(text + a); if (as < 0) bad |= BE_RANGE; }
                        else bad |= BE_AUX;                              // AS:A / AS:f: not an integer score
                    }
                    a += sz;
                } else if (ty == 'Z' || ty == 'H') {
                    while (a < end && text[a] != 0) a++;
                    if (a >= end) { bad |= BE_AUX; break; }
                    a++;
                    if (is_as) { bad |= BE_AUX; found = true; }
                } else if (ty == 'B') {
                    if (a + 5 > end) { bad |= BE_AUX; break; }
                    const int es = aux_size(text[a]); const long long cnt = (unsigned)rd32(text + a + 1);
                    if (!es || text[a] == 'A' || a + 5 + cnt * es > end) { bad |= BE_AUX; break; }
                    a += 5 + cnt * es;
                    if (is_as) { bad |= BE_AUX; found = true; }
                } else { bad |= BE_AUX; break; }
            }
            if (!found && !(bad & BE_AUX)) bad |= BE_NOAS;              // get_tag('AS') raises KeyError (:44,88)
        }
        R.rend[m] = (int)(pos + rlen);
        R.qlen[m] = (int)qlen; R.qs[m] = lead; R.qe[m] = (int)qlen - trail; R.as[m] = as;
    }
    for (int o = 16; o; o >>= 1) { nl = max(nl, __shfl_xor_sync(0xffffffffu, nl, o)); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
    if ((threadIdx.x & 31) == 0) { if (nl > 0) atomicMax(maxlen, nl); if (bad) atomicOr(err, bad); }
}

// grouped position p holds record g[p]; ks = the sorted read ids
__global__ void k_bam_bounds(int M, const unsigned *__restrict__ ks, int *rd_start, int *rd_end) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const unsigned r = ks[p];
    if (p == 0 || ks[p - 1] != r) rd_start[r] = p;
    if (p == M - 1 || ks[p + 1] != r) rd_end[r] = p + 1;
}

__device__ __forceinline__ bool tok_is(const unsigned char *s, int n, const char *lit, int ln) {
    if (n != ln) return false;
    for (int i = 0; i < n; i++) if (s[i] != (unsigned char)lit[i]) return false;
    return true;
}
__device__ __forceinline__ int find_primer(const Primers &P, const unsigned char *s, int n) {
    for (int k = 0; k < P.n; k++)
        if (tok_is(s, n, P.names + P.off[k], P.off[k + 1] - P.off[k])) return k;
    return -1;
}

// per read: primary record, n_alignments, rows, inferred primer row of single-alignment reads
__global__ void k_bam_reads(int NR, const unsigned char *__restrict__ text, const int *__restrict__ g, const int *__restrict__ rd_start,
                            const int *__restrict__ rd_end, Recs R, Primers P, int *rd_naln, int *rd_nrows, int *rd_pri, int4 *rd_inf,
                            int *maxnaln, int *err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0, naln = 0;
    if (r < NR) {
        const int s = rd_start[r], e = rd_end[r];
        int pri = -1, best = 0;
        for (int p = s; p < e; p++) {
            const int m = g[p];
            if (!(R.flag[m] & 2304)) {                                   // :42-44: first record with the highest AS
                const int a = R.as[m];
                if (pri < 0 || a > best) { pri = m; best = a; }
            }
        }
        naln = e - s;
        int nrows = naln, kind = 0, prim = 0, istrand = 0, iqs = 0, iqe = 0;
        if (pri < 0) { bad |= BE_NOPRIMARY; pri = g[s]; }               // :46-48
        else if (R.lseq[pri] == 0) bad |= BE_NOSEQ;                      // :101-103
        if (naln == 1 && !bad) {                                        // :109-158
            const int m = pri;
            const unsigned char *nm = text + R.noff[m];
            const int nlen = R.nlen[m];
            int st = 0;
            for (int i = 0; i < nlen; i++) if (nm[i] == '.') st = i + 1;
            int us = -1, nus = 0;
            for (int i = st; i < nlen; i++) if (nm[i] == '_') { nus++; if (us < 0) us = i; }
            if (nus != 1) bad |= BE_NAME;                                // p1, p2 = [...] needs exactly two tokens
            else {
                int l1 = us - st, l2 = nlen - us - 1;
                const unsigned char *t1 = nm + st, *t2 = nm + us + 1;
                const bool r1 = l1 > 0 && t1[l1 - 1] == 'R', r2 = l2 > 0 && t2[l2 - 1] == 'R';
                while (l1 > 0 && (t1[l1 - 1] == 'F' || t1[l1 - 1] == 'R')) l1--;
                while (l2 > 0 && (t2[l2 - 1] == 'F' || t2[l2 - 1] == 'R')) l2--;
                const int qs = R.qs[m], qe = R.qe[m], ql = R.qlen[m];
                if (!(qs > 5 && ql - qe > 5)) {
                    if (!tok_is(t1, l1, "False", 5)) {
                        prim = find_primer(P, t1, l1);
                        if (prim < 0) { bad |= BE_PRIMER; prim = 0; }
                        else { kind = 1; istrand = r1; iqs = 0; iqe = P.len[prim]; }
                    } else if (!tok_is(t2, l2, "False", 5)) {
                        prim = find_primer(P, t2, l2);
                        if (prim < 0) { bad |= BE_PRIMER; prim = 0; }
                        else { kind = 2; istrand = r2; iqs = ql - P.len[prim]; iqe = ql; }
                    }
                    if (kind) { naln = 2; nrows = 2; }
                }
            }
        }
        rd_naln[r] = naln; rd_nrows[r] = nrows; rd_pri[r] = pri;
        rd_inf[r] = make_int4(kind | (istrand << 2) | (prim << 3), iqs, iqe, 0);
    }
    for (int o = 16; o; o >>= 1) { naln = max(naln, __shfl_xor_sync(0xffffffffu, naln, o)); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
    if ((threadIdx.x & 31) == 0) { atomicMax(maxnaln, naln); if (bad) atomicOr(err, bad); }
}

struct Regions { int n; const int *chrom, *start, *end; };

__global__ void k_bam_rows(int NR, const int *__restrict__ g, const int *__restrict__ rd_start, const int *__restrict__ rd_end,
                           const int *__restrict__ rd_row0, const int *__restrict__ rd_naln, const int *__restrict__ rd_pri,
                           const int4 *__restrict__ rd_inf, Recs R, Regions G, int n_ref, Rows W) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= NR) return;
    const int s = rd_start[r], e = rd_end[r], pri = rd_pri[r], naln = rd_naln[r];
    const int4 inf = rd_inf[r];
    const int kind = inf.x & 3;
    const bool pri_rev = (R.flag[pri] & 16) != 0;
    int j = rd_row0[r];
    const int ql1 = R.qlen[g[s]];
    for (int pass = 0; pass < 2; pass++) {
        if ((pass == 0 && kind == 1) || (pass == 1 && kind == 2)) {     // :123-139 / :142-157
            W.rid[j] = r; W.chrom[j] = n_ref + (inf.x >> 3); W.rstart[j] = 0; W.rend[j] = 0; W.naln[j] = 2; W.aln[j] = 0;
            W.qstart[j] = inf.y; W.qend[j] = inf.z; W.strand[j] = (inf.x >> 2) & 1; W.mapq[j] = 0; W.qlen[j] = ql1; W.as[j] = 0;
            W.inferred[j] = 1; W.overlaps[j] = -1; W.seq_rec[j] = -1;
            j++;
        }
        if (pass == 0)
            for (int p = s; p < e; p++) {
                const int m = g[p];
                int qs = R.qs[m], qe = R.qe[m];
                const int ql = R.qlen[m];
                const bool rev = (R.flag[m] & 16) != 0;
                if (rev != pri_rev) { const int t = ql - qe; qe = t + qe - qs; qs = t; }      // :59-62
                const int c = R.ref[m], st = R.pos1[m], en = R.rend[m];
                int ov = 0;
                for (int k = 0; k < G.n; k++)                            // (start, end] against (s, e]  (:70-76)
                    if (G.chrom[k] == c && st < G.end[k] && G.start[k] < en) { ov = 1; break; }
                W.rid[j] = r; W.chrom[j] = c; W.rstart[j] = st; W.rend[j] = en; W.naln[j] = naln; W.aln[j] = qe - qs;
                W.qstart[j] = qs; W.qend[j] = qe; W.strand[j] = rev; W.mapq[j] = R.mapq[m]; W.qlen[j] = ql; W.as[j] = R.as[m];
                W.inferred[j] = 0; W.overlaps[j] = ov; W.seq_rec[j] = m == pri ? m : -1;
                j++;
            }
    }
}

// LSD string sort, one pass: key = bytes [4c, 4c+4) of the name of read perm[i], big-endian, zero padded
__global__ void k_bam_namekey(int NR, const unsigned char *__restrict__ text, const int *__restrict__ perm, const int *__restrict__ first_rec,
                              const long long *__restrict__ noff, const int *__restrict__ nlen, int c, unsigned *key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NR) return;
    const int m = first_rec[perm[i]];
    const unsigned char *s = text + noff[m];
    const int n = nlen[m];
    unsigned k = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) { const int q = 4 * c + b; k = (k << 8) | (q < n ? s[q] : 0u); }
    key[i] = k;
}
__global__ void k_bam_gather_key(int n, const int *__restrict__ idx, const int *__restrict__ val, int bias, int sign, unsigned xr, unsigned *key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = (unsigned)(bias + sign * val[idx[i]]) ^ xr;
}
__global__ void k_bam_invert(int n, const int *__restrict__ perm, int *rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[perm[i]] = i;
}
__global__ void k_bam_rowkey(int n, const int *__restrict__ idx, const int *__restrict__ rid, const int *__restrict__ rank, unsigned *key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = (unsigned)rank[rid[idx[i]]];
}
// order[j] = source row of output row j; rows of a read are contiguous: first / last aln_size per read (:165-172)
__global__ void k_bam_anchor(int N, const int *__restrict__ order, const int *__restrict__ rid, const int *__restrict__ aln, int *fa, int *la) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int src = order[j], r = rid[src];
    if (j == 0 || rid[order[j - 1]] != r) fa[r] = aln[src];
    if (j == N - 1 || rid[order[j + 1]] != r) la[r] = aln[src];
}
__global__ void k_bam_final(int N, const int *__restrict__ order, Rows S, const int *__restrict__ fa, const int *__restrict__ la,
                            const int *__restrict__ rank, const int *__restrict__ first_rec, Rows D, int *short_anchor, int *name_rec) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int s = order[j], r = S.rid[s];
    D.rid[j] = rank[r]; D.chrom[j] = S.chrom[s]; D.rstart[j] = S.rstart[s]; D.rend[j] = S.rend[s]; D.naln[j] = S.naln[s];
    D.aln[j] = S.aln[s]; D.qstart[j] = S.qstart[s]; D.qend[j] = S.qend[s]; D.strand[j] = S.strand[s]; D.mapq[j] = S.mapq[s];
    D.qlen[j] = S.qlen[s]; D.as[j] = S.as[s]; D.inferred[j] = S.inferred[s]; D.overlaps[j] = S.overlaps[s]; D.seq_rec[j] = S.seq_rec[s];
    short_anchor[j] = (fa[r] < 50 || la[r] < 50) ? 1 : 0;
    name_rec[j] = first_rec[r];
}

// ---- egress: one line per row (:176-181)
struct Names { const char *text; const int *off; };       // chrom id -> name bytes [off[c], off[c+1])
struct Emit {
    Rows D; const int *short_anchor, *name_rec; Recs R; Names C;
    int ver_len; char ver[48]; int with_regions, ov_float;
};
__device__ __forceinline__ int ov_len(const Emit &E, int ov) {
    if (!E.with_regions) return 0;
    return 1 + (ov < 0 ? 0 : (E.ov_float ? 3 : 1));
}
__global__ void k_bam_outlen(int N, Emit E, long long *len) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const Rows &D = E.D;
    const int c = D.chrom[j], sr = D.seq_rec[j];
    long long n = E.C.off[c + 1] - E.C.off[c];
    n += tsv::dec_digits(D.rstart[j]) + tsv::dec_digits(D.rend[j]) + E.R.nlen[E.name_rec[j]] + tsv::dec_digits(D.naln[j]) +
         tsv::dec_digits(D.aln[j]) + tsv::dec_digits(D.qstart[j]) + tsv::dec_digits(D.qend[j]) + 1 + tsv::dec_digits(D.mapq[j]) +
         tsv::dec_digits(D.qlen[j]) + tsv::dec_digits(D.as[j]) + 1 + E.ver_len + 1 + (sr >= 0 ? E.R.lseq[sr] : 0);
    n += 15 /* tabs */ + ov_len(E, D.overlaps[j]) + 1 /* newline */;
    len[j] = n;
}
__device__ __forceinline__ unsigned char seq_base(const unsigned char *sq, int i) {
    const unsigned char b = sq[i >> 1];
    return (unsigned char)"=ACMGRSVTWYHKDBN"[(i & 1) ? (b & 15) : (b >> 4)];
}
__device__ __forceinline__ unsigned char comp_base(unsigned char c) {     // pysam get_forward_sequence: ACGT (and N, X) only
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}
__global__ void k_bam_emit(int N, Emit E, const unsigned char *__restrict__ text, const long long *__restrict__ out_off, long long base,
                           unsigned char *out) {
    const long long wg = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= N) return;
    const int j = (int)wg;
    const Rows &D = E.D;
    unsigned char *o = out + base + out_off[j];
    const int sr = D.seq_rec[j];
    const int ls = sr >= 0 ? E.R.lseq[sr] : 0;
    int pre = 0;
    if (lane == 0) {
        unsigned char *q = o;
        const int c = D.chrom[j];
        for (int i = E.C.off[c]; i < E.C.off[c + 1]; i++) *q++ = (unsigned char)E.C.text[i];
        *q++ = '\t'; q = tsv::put_dec(q, D.rstart[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.rend[j]);
        *q++ = '\t';
        { const int m = E.name_rec[j]; const unsigned char *nm = text + E.R.noff[m]; const int nl = E.R.nlen[m]; for (int i = 0; i < nl; i++) *q++ = nm[i]; }
        *q++ = '\t'; q = tsv::put_dec(q, D.naln[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.aln[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.qstart[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.qend[j]);
        *q++ = '\t'; *q++ = D.strand[j] ? '-' : '+';
        *q++ = '\t'; q = tsv::put_dec(q, D.mapq[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.qlen[j]);
        *q++ = '\t'; q = tsv::put_dec(q, D.as[j]);
        *q++ = '\t'; *q++ = E.short_anchor[j] ? '1' : '0';
        *q++ = '\t'; for (int i = 0; i < E.ver_len; i++) *q++ = (unsigned char)E.ver[i];
        *q++ = '\t'; *q++ = D.inferred[j] ? '1' : '0';
        *q++ = '\t';
        pre = (int)(q - o);
    }
    pre = __shfl_sync(0xffffffffu, pre, 0);
    if (ls > 0) {
        const unsigned char *sq = text + E.R.seq_off[sr];
        const bool rev = (E.R.flag[sr] & 16) != 0;
        for (int i = lane; i < ls; i += 32)
            o[pre + i] = rev ? comp_base(seq_base(sq, ls - 1 - i)) : seq_base(sq, i);
    }
    if (lane == 0) {
        unsigned char *q = o + pre + ls;
        if (E.with_regions) {
            *q++ = '\t';
            const int ov = D.overlaps[j];
            if (ov >= 0) { *q++ = ov ? '1' : '0'; if (E.ov_float) { *q++ = '.'; *q++ = '0'; } }
        }
        *q++ = '\n';
    }
}

}  // namespace bam


// GPU ingest / egress of `<base>.mappings.bed`: the TSV main.py:209 reads with pandas
// (columns written by collect_mapping_info.py:176-181) and the `<base>.mappings.cluster.bed` main.py:349 writes.
//
//   k_tsv_count / k_tsv_lines   newline positions -> line starts (one pass over the bytes + an exclusive scan)
//   k_tsv_parse                 one thread per line walks the first fields only (the long `seq` column is never touched):
//                               decimal ints of the wanted columns, (offset, length, FNV-1a hash) of qname and chrom
//   k_tsv_intern_*              open-addressing hash table keyed by the 64-bit hash: first row of every distinct string,
//                               byte-for-byte verification against that row, dense ids in order of first appearance
//                               (= pandas.factorize = the singleton numbering order of main.py:336-341)
//   k_tsv_outlen / k_tsv_emit   every input line + "\t<cluster>.0\t<n_reads>.0" (the float columns of main.py:334-342)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsv {

constexpr int CHUNK = 64;                 // bytes per thread in the newline pass
enum { TE_FIELDS = 1, TE_INT = 2, TE_COLLISION = 4, TE_RANGE = 8 };
enum { W_CHROM = 0, W_RSTART, W_REND, W_QNAME, W_NALN, W_ALN, W_QSTART, W_QEND, W_SCORE, W_N };

struct Want { int col[W_N]; int last; };   // field index of every wanted column (-1: absent), highest wanted index

__global__ void k_tsv_count(const unsigned char *__restrict__ text, long long n, int *cnt) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long b0 = c * CHUNK;
    if (b0 >= n) return;
    int k = 0;
    if (b0 + CHUNK <= n) {
#pragma unroll
        for (int j = 0; j < CHUNK / 16; j++) {
            const uint4 v = __ldg((const uint4 *)(text + b0) + j);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned x = w[q] ^ 0x0a0a0a0au;                 // bytes equal to '\n' become 0
                k += __popc(((x - 0x01010101u) & ~x & 0x80808080u));
            }
        }
    } else {
        for (long long i = b0; i < n; i++) k += text[i] == '\n';
    }
    cnt[c] = k;
}
// line_start[l] = offset of the first byte of line l (line 0 starts at 0; every '\n' at p starts a line at p + 1)
__global__ void k_tsv_lines(const unsigned char *__restrict__ text, long long n, const int *__restrict__ pre, long long *line_start) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long b0 = c * CHUNK;
    if (b0 >= n) return;
    int l = pre[c] + 1;
    const long long b1 = b0 + CHUNK < n ? b0 + CHUNK : n;
    for (long long i = b0; i < b1; i++)
        if (text[i] == '\n') line_start[l++] = i + 1;
    if (c == 0) line_start[0] = 0;
}
__device__ __forceinline__ unsigned long long fnv1a(const unsigned char *p, int len, unsigned long long seed) {
    unsigned long long h = 1469598103934665603ull ^ seed;
    for (int i = 0; i < len; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h == ~0ull ? 0x1234567ull : h;                              // ~0 marks an empty slot
}
// rows = data lines (line 0 is the header): row r is line r + 1
__global__ void k_tsv_parse(const unsigned char *__restrict__ text, const long long *__restrict__ line_start, int n_rows, Want w,
                            unsigned long long seed, int *rstart, int *rend,
                            int *naln, int *aln, int *qstart, int *qend, int *score, long long *q_off, int *q_len,
                            unsigned long long *q_hash, long long *c_off, int *c_len, unsigned long long *c_hash, int *err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long s = line_start[r + 1], e = line_start[r + 2] - 1;   // [s, e): the line without its '\n'
    long long p = s;
    int vals[W_N];
#pragma unroll
    for (int k = 0; k < W_N; k++) vals[k] = 0;
    int bad = 0;
    for (int f = 0; f <= w.last; f++) {
        long long fe = p;
        while (fe < e && text[fe] != '\t') fe++;
        if (p > e) { bad |= TE_FIELDS; break; }
        int which = -1;
#pragma unroll
        for (int k = 0; k < W_N; k++) if (w.col[k] == f) which = k;
        if (which == W_QNAME) { q_off[r] = p; q_len[r] = (int)(fe - p); q_hash[r] = fnv1a(text + p, (int)(fe - p), seed); }
        else if (which == W_CHROM) { c_off[r] = p; c_len[r] = (int)(fe - p); c_hash[r] = fnv1a(text + p, (int)(fe - p), seed); }
        else if (which >= 0) {
            long long i = p;
            bool neg = false;
            if (i < fe && text[i] == '-') { neg = true; i++; }
            long long v = 0;
            if (i == fe) bad |= TE_INT;
            for (; i < fe; i++) {
                const int d = text[i] - '0';
                if ((unsigned)d > 9u) {                                 // "12.0": a pandas float column that holds integers
                    if (text[i] == '.') { for (long long j = i + 1; j < fe; j++) if (text[j] != '0') bad |= TE_INT; }
                    else bad |= TE_INT;
                    break;
                }
                v = v * 10 + d;
                if (v > 0x7fffffffLL) { bad |= TE_RANGE; v = 0; }
            }
            vals[which] = (int)(neg ? -v : v);
        }
        p = fe + 1;
    }
    if (bad) atomicOr(err, bad);
    rstart[r] = vals[W_RSTART]; rend[r] = vals[W_REND]; naln[r] = vals[W_NALN]; aln[r] = vals[W_ALN];
    qstart[r] = vals[W_QSTART]; qend[r] = vals[W_QEND];
    if (score) score[r] = vals[W_SCORE];
}
// ---- interning: slot table {hash, first row}; capacity is a power of two
__global__ void k_tsv_intern_insert(int n_rows, const unsigned long long *__restrict__ hash, unsigned long long *keys, int *first,
                                    unsigned mask, int *slot_of_row) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const unsigned long long h = hash[r];
    unsigned s = (unsigned)(h ^ (h >> 32)) & mask;
    for (;;) {
        const unsigned long long cur = atomicCAS(&keys[s], ~0ull, h);
        if (cur == ~0ull || cur == h) break;
        s = (s + 1) & mask;
    }
    atomicMin(&first[s], r);
    slot_of_row[r] = s;
}
// every row's string must equal, byte for byte, the string of the first row of its slot (a 64-bit hash collision between
// two different names is reported, never silently merged); is_first flags the first appearance of every distinct string
__global__ void k_tsv_intern_verify(int n_rows, const unsigned char *__restrict__ text, const long long *__restrict__ off,
                                    const int *__restrict__ len, const int *__restrict__ slot_of_row, const int *__restrict__ first,
                                    int *is_first, int *err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int f = first[slot_of_row[r]];
    is_first[r] = f == r;
    if (f != r) {
        const int l = len[r];
        bool same = len[f] == l;
        const unsigned char *a = text + off[r], *b = text + off[f];
        for (int i = 0; same && i < l; i++) same = a[i] == b[i];
        if (!same) atomicOr(err, TE_COLLISION);
    }
}
__global__ void k_tsv_intern_ids(int n_rows, const int *__restrict__ slot_of_row, const int *__restrict__ first,
                                 const int *__restrict__ id_at_row, int *id, int *first_row_of_id) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int f = first[slot_of_row[r]];
    const int i = id_at_row[f];
    id[r] = i;
    if (f == r && first_row_of_id) first_row_of_id[i] = r;
}
// ---- egress
__device__ __forceinline__ int dec_digits(int v) {
    int d = 1;
    unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    while (u >= 10u) { u /= 10u; d++; }
    return d + (v < 0);
}
__device__ __forceinline__ unsigned char *put_dec(unsigned char *o, int v) {
    const int nd = dec_digits(v);
    unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    if (v < 0) o[0] = '-';
    for (int k = nd - 1; k >= (v < 0); k--) { o[k] = (unsigned char)('0' + u % 10u); u /= 10u; }
    return o + nd;
}
// main.py:334-342: the left merge leaves NaN for singletons, which makes `cluster` / `n_reads` float columns ("12.0");
// a table in which every read ended in a cluster has no NaN and keeps integer columns ("12").  *any_single != 0 <=> floats.
__global__ void k_tsv_any_single(int n_reads, const int *__restrict__ nreads, int *any_single) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool s = r < n_reads && nreads[r] == 1;
    if (__syncthreads_or(s) && threadIdx.x == 0) *any_single = 1;
}
// output length of every line: header + "\tcluster\tn_reads\n"; row: line + "\t<c>.0\t<n>.0\n" (floats) or "\t<c>\t<n>\n"
__global__ void k_tsv_outlen(int n_lines, const long long *__restrict__ line_start, const int *__restrict__ rid,
                             const int *__restrict__ cluster, const int *__restrict__ nreads, const int *__restrict__ any_single,
                             long long *len) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const long long body = line_start[l + 1] - 1 - line_start[l];
    const int frac = *any_single ? 2 : 0;
    if (l == 0) len[l] = body + 17;                                     // "\tcluster\tn_reads\n"
    else { const int r = rid[l - 1]; len[l] = body + 1 + dec_digits(cluster[r]) + frac + 1 + dec_digits(nreads[r]) + frac + 1; }
}
// one warp per line: coalesced byte copy, lane 0 appends the two columns
__global__ void k_tsv_emit(int n_lines, const unsigned char *__restrict__ text, const long long *__restrict__ line_start,
                           const int *__restrict__ rid, const int *__restrict__ cluster, const int *__restrict__ nreads,
                           const int *__restrict__ any_single, const long long *__restrict__ out_off, unsigned char *out) {
    const long long wg = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= n_lines) return;
    const int l = (int)wg;
    const long long s = line_start[l], body = line_start[l + 1] - 1 - s;
    unsigned char *o = out + out_off[l];
    for (long long i = lane; i < body; i += 32) o[i] = text[s + i];
    if (lane == 0) {
        unsigned char *q = o + body;
        if (l == 0) { const char *h = "\tcluster\tn_reads\n"; for (int i = 0; i < 17; i++) q[i] = (unsigned char)h[i]; }
        else {
            const int r = rid[l - 1];
            const bool fl = *any_single != 0;
            *q++ = '\t'; q = put_dec(q, cluster[r]); if (fl) { *q++ = '.'; *q++ = '0'; }
            *q++ = '\t'; q = put_dec(q, nreads[r]); if (fl) { *q++ = '.'; *q++ = '0'; }
            *q++ = '\n';
        }
    }
}

}  // namespace tsv

// fslr_b200 — B200 (sm_100a) implementation of the read-clustering step of kcleal/fslr.
//
// Replaces, behind the C ABI of include/fslr_b200.h, the reference calls of
// /root/reference/fslr/main.py:233-257,334-342 into /root/reference/fslr/cluster.py:
//   keep_fillings (cluster.py:14-31)           -> k_first_last, k_keep, k_fill_records
//   prepare_data + mask_sequences2 (:89-121)    -> k_mask_flags, k_compact_fillings, radix sort by start, k_build_items
//   query_intervals dict order (:189-191)       -> k_first_dp, k_is_first, scan, sort by query rank
//   build_interval_trees / IntervalMap (:124-130, third-party superintervals)
//                                               -> sort by (chrom, start, end desc, data order), k_ub, prefix-max of ends
//   query_interval_trees (:187-227)             -> k_pair (order-free relation + capped degree) and k_replay
//                                                  (reads that can reach edge_threshold, in query order)
//   different_lengths_or_alignments (:178-183), overall_jaccard_similarity (:140-170),
//   calculate_overlap (:133-136), cutoff lookup (:218-219)
//                                               -> integer thresholds T/Lq/Ln/umax (k_read_info, k_records, host umax)
//   get_subgraphs / networkx (:230-234)         -> k_union_entries, k_union_edges, k_flatten (lock-free union-find,
//                                                  root = smallest query rank = the component's first-inserted node)
//   cluster / n_reads columns (main.py:251-257,334-342) -> k_roots, k_number
//
// Why the sequential reference loop parallelises exactly: a filling's scan of
// search_values() results covers a contiguous run of sorted positions [stop, ub] walked downwards, so the
// loop state that later queries can observe is one integer per filling.  Reads whose number of passing
// candidates is below edge_threshold can never break: their stop is the chromosome start and their edges follow
// from the order-free relation.  The remaining reads are replayed in query order by a persistent ticket kernel in
// which a warp only ever waits for reads holding smaller tickets.
#include <cuda_runtime.h>
#include "prims.cuh"
#include "tsv.cuh"
#include <string>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "fslr_b200.h"

#define FSLRC_VERSION 1
#define LMAX FSLRC_MAX_FILLINGS

enum { EF_RANGE = 1, EF_ZERO = 2, EF_TOOMANY = 4, EF_NALN = 8, EF_OVERFLOW = 16 };

// ---------------------------------------------------------------- context
struct fslrc_ctx {
    int device;
    char err[512];
    cudaStream_t stream;
    std::vector<void *> allocs;
    int64_t *h_pin;          // pinned scratch for read-backs (64 x int64)
    cudaEvent_t ev[FSLRC_N_STAGES + 1];
    // pipeline state (kept between the fslrc_mg_* stages)
    struct Pipe *pipe;
    struct TsvState *tsv;    // parsed mappings.bed kept on the device between fslrc_tsv_open and fslrc_tsv_close
    long long launches;      // kernels of this library launched since fslrc_create
};

static void tsv_free(fslrc_ctx *ctx);

static const char *STAGE_NAMES[FSLRC_N_STAGES] = {
    "h2d", "keep_fillings", "data_order_mask", "query_rank_read_lists", "chrom_sort", "records_bands",
    "pair_kernel", "saturating_set", "replay", "union_find", "numbering", "d2h"};

static int fail(fslrc_ctx *c, int code, const char *fmt, const char *a = "") {
    snprintf(c->err, sizeof(c->err), fmt, a);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return FSLRC_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

template <typename T>
static int dalloc(fslrc_ctx *ctx, T **p, int64_t n) {
    void *q = nullptr;
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(T);
    CK(cudaMallocAsync(&q, bytes, ctx->stream));
    ctx->allocs.push_back(q);
    *p = (T *)q;
    return 0;
}
static void free_all(fslrc_ctx *ctx) {
    for (void *p : ctx->allocs) cudaFreeAsync(p, ctx->stream);
    ctx->allocs.clear();
}
#define DA(ptr, n)                                  \
    do {                                            \
        int r__ = dalloc(ctx, &(ptr), (int64_t)(n)); \
        if (r__) return r__;                        \
    } while (0)

// every launch of one of OUR kernels goes through KL so that the count can be reported (bench.py "gpu_launches")
#define KL(kernel, grid, block, ...)                         \
    do {                                                     \
        ctx->launches++;                                     \
        kernel<<<(grid), (block), 0, st>>>(__VA_ARGS__);     \
    } while (0)

static inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

// ---------------------------------------------------------------- small utility kernels
template <typename T>
__global__ void k_fill(T *p, int64_t n, T v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_widen(int64_t n, const unsigned char *__restrict__ c8, const unsigned short *__restrict__ n16, int *chrom, int *naln,
                        int *aln, const int *__restrict__ qstart, const int *__restrict__ qend) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (c8) chrom[i] = c8[i];
    if (n16) naln[i] = n16[i];
    if (aln) aln[i] = qend[i] - qstart[i];                             // aln_size = qend - qstart (collect_mapping_info.py:88)
}
__global__ void k_iota(int *p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// exact threshold in the reference's double arithmetic: min{o >= 0 : fl(o/a) >= p}  (cluster.py:133-136,179,181)
__device__ __forceinline__ int thr_f64(int a, double p) {
    if (!(p > 0.0)) return 0;
    double da = (double)a;
    double x = ceil(__dmul_rn(p, da));
    if (x >= 2147483000.0) return 2147483647;
    long long o = (long long)x;
    while (o > 0 && __ddiv_rn((double)(o - 1), da) >= p) --o;
    while (__ddiv_rn((double)o, da) < p) ++o;
    return (int)o;
}

// ---------------------------------------------------------------- stage 1: keep_fillings (cluster.py:14-31)
__global__ void k_first_last(int A, int R, const int *__restrict__ rid, int *first, int *last, int *err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    int r = rid[i];
    if ((unsigned)r >= (unsigned)R) { atomicOr(err, EF_RANGE); return; }
    atomicMin(&first[r], i);
    atomicMax(&last[r], i);
}
__global__ void k_keep(int A, int R, const int *__restrict__ rid, const int *__restrict__ first, const int *__restrict__ last,
                       const int *__restrict__ qstart, const int *__restrict__ qend, int *flag, int *qmin, int *qmax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    int r = rid[i];
    int keep = 0;
    if ((unsigned)r < (unsigned)R) {
        keep = (i != first[r] && i != last[r]);
        if (keep) { atomicMax(&qmax[r], qend[i]); atomicMin(&qmin[r], qstart[i]); }
    }
    flag[i] = keep;
}
// fillings in bed order as packed records: FR0[k] = {read_id, chrom, start, end}, FR1[k] = {aln_size, n_alignments}
// (start/end = min/max of rstart, rend: cluster.py:111-112).  One coalesced pass over the kept rows.
__global__ void k_fill_records(int A, const int *__restrict__ flag, const int *__restrict__ pos, const int *__restrict__ rid,
                               const int *__restrict__ chrom, const int *__restrict__ rstart, const int *__restrict__ rend,
                               const int *__restrict__ aln, const int *__restrict__ naln, int n_chrom, int4 *FR0, int2 *FR1, int *err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A || !flag[i]) return;
    const int k = pos[i];
    const int c = chrom[i], rs = rstart[i], re = rend[i];
    if ((unsigned)c >= (unsigned)n_chrom || min(rs, re) < 0) atomicOr(err, EF_RANGE);
    FR0[k] = make_int4(rid[i], c, min(rs, re), max(rs, re));
    FR1[k] = make_int2(aln[i], naln[i]);
}

// ---------------------------------------------------------------- stage 2: prepare_data + mask (cluster.py:109-121, 89-106)
__device__ __forceinline__ bool is_masked(const int4 f, int n_chrom, const long long *__restrict__ clen,
                                          const unsigned char *__restrict__ cmasked, int sub_on, long long subtel) {
    if ((unsigned)f.y >= (unsigned)n_chrom) return true;
    bool masked = cmasked[f.y] != 0;                                          // cluster.py:96
    const long long cl = clen[f.y];
    if (sub_on && cl > 1000000 && ((long long)f.z < subtel || cl - (long long)f.w < subtel)) masked = true;   // :94,98-100
    return masked;
}
// flags over the fillings taken in the order `perm` (NULL = bed order)
__global__ void k_mask_flags(int F, const int *__restrict__ perm, const int4 *__restrict__ FR0, int n_chrom,
                             const long long *__restrict__ clen, const unsigned char *__restrict__ cmasked, int sub_on,
                             long long subtel, int *flag, int *err) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    int fk = k;
    if (perm) { fk = perm[k]; if ((unsigned)fk >= (unsigned)F) { atomicOr(err, EF_RANGE); flag[k] = 0; return; } }
    flag[k] = is_masked(FR0[fk], n_chrom, clen, cmasked, sub_on, subtel) ? 0 : 1;
}
// unmasked fillings, compacted: sort key (start) + filling index, or directly the data-order list when perm is given
__global__ void k_compact_fillings(int F, const int *__restrict__ perm, const int *__restrict__ flag, const int *__restrict__ pos,
                                   const int4 *__restrict__ FR0, int *key, int *val) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F || !flag[k]) return;
    const int fk = perm ? perm[k] : k;
    const int u = pos[k];
    if (key) key[u] = FR0[fk].z;
    val[u] = fk;
}
// data items in data order: IT0[d] = {read_id, chrom, start, end}, IT1[d] = {aln_size, n_alignments}
__global__ void k_build_items(int D, const int *__restrict__ dfill, const int4 *__restrict__ FR0, const int2 *__restrict__ FR1,
                              int4 *IT0, int2 *IT1, int *firstdp, int *err) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int fk = dfill[d];
    const int4 f0 = FR0[fk];
    const int2 f1 = FR1[fk];
    IT0[d] = f0; IT1[d] = f1;
    if (f1.x <= 0 || f1.y <= 0) atomicOr(err, EF_ZERO);
    if (f1.y >= 65535) atomicOr(err, EF_RANGE);
    atomicMin(&firstdp[f0.x], d);
}

// ---------------------------------------------------------------- stage 3: query rank (cluster.py:189-191) + per-read lists
__global__ void k_is_first(int D, const int4 *__restrict__ IT0, const int *__restrict__ firstdp, int *flag) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) flag[d] = (firstdp[IT0[d].x] == d);
}
__global__ void k_rank_reads(int R, const int *__restrict__ firstdp, const int *__restrict__ rank_at, int *q_of_rid, int *rid_of_q) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int f = firstdp[r];
    int q = -1;
    if (f != 0x7fffffff) { q = rank_at[f]; rid_of_q[q] = r; }
    q_of_rid[r] = q;
}
__global__ void k_item_q(int D, const int4 *__restrict__ IT0, const int *__restrict__ q_of_rid, int *it_q) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) it_q[d] = q_of_rid[IT0[d].x];
}
// rm order: items grouped by query rank, data order inside a read
__global__ void k_read_bounds(int D, const int *__restrict__ qs /*sorted q*/, const int *__restrict__ rm_dp, int *rmidx, int *off, int *len_end) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= D) return;
    int q = qs[m];
    rmidx[rm_dp[m]] = m;
    if (m == 0 || qs[m - 1] != q) off[q] = m;
    if (m == D - 1 || qs[m + 1] != q) len_end[q] = m + 1;
}
// per read: RI[q] = {qlen2, Lq, n_alignments | Ln << 16, off << 6 | (L - 1)}: the ratio thresholds of
// cluster.py:26-29,178-183 and where the read's fillings live in read-major order
__global__ void k_read_info(int Q, const int *__restrict__ rid_of_q, const int *__restrict__ off, const int *__restrict__ len_end,
                            const int *__restrict__ rm_dp, const int2 *__restrict__ IT1, const int *__restrict__ qmin,
                            const int *__restrict__ qmax, double qlen_c, double naln_c, int4 *RI, int *err) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int r = rid_of_q[q], o = off[q], L = len_end[q] - o;
    if (L > LMAX) { atomicOr(err, EF_TOOMANY); L = LMAX; }
    long long ql = (long long)qmax[r] - (long long)qmin[r];
    int na = IT1[rm_dp[o]].y;
    if (ql <= 0 || na <= 0) { atomicOr(err, EF_ZERO); ql = ql <= 0 ? 1 : ql; na = na <= 0 ? 1 : na; }
    if (ql > 0x7fffffffLL) { atomicOr(err, EF_RANGE); ql = 1; }
    int Ln = thr_f64(na, naln_c);
    if (Ln > 65535) Ln = 65535;
    RI[q] = make_int4((int)ql, thr_f64((int)ql, qlen_c), (na & 0xffff) | (Ln << 16), (int)(((unsigned)o << 6) | (unsigned)((L - 1) & 63)));
}
// ---------------------------------------------------------------- stage 4/5: IntervalMap order + records + bands
__global__ void k_end_keys(int D, const int4 *__restrict__ IT0, unsigned *key) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) key[d] = ~(unsigned)IT0[d].w;                        // ascending ~end == end descending
}
__global__ void k_gather_key(int D, const int *__restrict__ dp_in, const int4 *__restrict__ IT0, int which, unsigned *key) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < D) { const int4 it = IT0[dp_in[k]]; key[k] = (unsigned)(which ? it.y : it.z); }      // start / chromosome of the item
}
// IntervalMap order without sorting by start again: data order is already sorted by start, so a STABLE partition by
// chromosome yields (chrom, start, data order); what is missing is "end descending" inside runs of equal (chrom, start).
// Those runs are short (PCR duplicates), and in data order their members sit in one block of equal starts: every item
// counts, with coalesced neighbour reads, how many members of its run precede it in data order (idx) and how many must
// precede it in the final order (rank: larger end, or equal end and earlier in data order).  The partition moves the run
// as a block, so the item's final position is its partition position + (rank - idx).  val[d] = d | (rank - idx + 32) << 26.
// Runs that do not fit the window raise `overflow` and the caller falls back to the two full radix sorts.
#define TIE_WIN 48
__global__ void k_tie_delta(int D, const int4 *__restrict__ IT0, unsigned *ckey, unsigned *val, unsigned long long *overflow, int *err) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int4 me = IT0[d];
    if (d > 0 && IT0[d - 1].z > me.z) atomicOr(err, EF_RANGE);       // a caller-supplied `order` that does not sort by start
    int idx = 0, rank = 0;
    bool ovf = false;
    for (int k = 1;; k++) {                                          // earlier in data order
        if (d - k < 0) break;
        const int4 o = IT0[d - k];
        if (o.z != me.z) break;
        if (k > TIE_WIN) { ovf = true; break; }
        if (o.y == me.y) { idx++; rank += o.w >= me.w; }
    }
    for (int k = 1;; k++) {                                          // later in data order
        if (d + k >= D) break;
        const int4 o = IT0[d + k];
        if (o.z != me.z) break;
        if (k > TIE_WIN) { ovf = true; break; }
        if (o.y == me.y) rank += o.w > me.w;
    }
    if (idx > 31 || rank > 31) ovf = true;
    if (ovf) { atomicAdd(overflow, 1ull); rank = idx; }
    ckey[d] = (unsigned)me.y;
    val[d] = (unsigned)d | ((unsigned)(rank - idx + 32) << 26);
}
__global__ void k_apply_delta(int D, const unsigned *__restrict__ val, int *s_dp) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D) return;
    const unsigned v = val[p];
    s_dp[p + (int)(v >> 26) - 32] = (int)(v & 0x3ffffffu);
}
// SR0[p] = {start, end, T, q | fi << 26} (fi = index of the filling in its read's list); SR1[p] = the read's RI record;
// RM[2m] = {chrom, start, end, T}, RM[2m+1] = {pos, ub (closed band, replay), lbT, ubT (tight band, pair kernel)}: one
// 32-byte sector per filling in read-major order, written once by k_bands
#define QMASK 0x3ffffff
__global__ void k_records(int D, const int *__restrict__ s_dp, const int *__restrict__ rmidx, const int *__restrict__ it_q,
                          const int4 *__restrict__ IT0, const int2 *__restrict__ IT1, const int4 *__restrict__ RI,
                          double overlap, int4 *SR0, int4 *SR1, int *s_m,
                          int *s_chrom, int *s_end, int *chrom_lo, int *chrom_hi, int *err) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D) return;
    const int d = s_dp[p], m = rmidx[d], q = it_q[d];
    const int4 it = IT0[d];
    const int c = it.y, s = it.z, e = it.w;
    const int2 i1 = IT1[d];
    const int T = thr_f64(max(i1.x, 1), overlap);
    const int4 ri = RI[q];
    if ((ri.z & 0xffff) != i1.y) atomicOr(err, EF_NALN);            // n_alignments must be constant over the rows of a read
    const int fi = m - (int)((unsigned)ri.w >> 6);                   // index of this filling in its read's list
    SR0[p] = make_int4(s, e, T, (int)((unsigned)q | ((unsigned)fi << 26)));
    SR1[p] = ri;
    s_m[p] = m;
    s_chrom[p] = c; s_end[p] = e;
    const int cprev = p > 0 ? IT0[s_dp[p - 1]].y : -1;
    const int cnext = p < D - 1 ? IT0[s_dp[p + 1]].y : -1;
    if (cprev != c) chrom_lo[c] = p;
    if (cnext != c) chrom_hi[c] = p + 1;
}
// Bands of sorted position p, and the read-major record of its filling.
// ub(p): last sorted position on the chromosome with start <= end_p (IntervalMap upper bound), by galloping
// from p (the band is short: ~2 log2(band) probes instead of log2(D)).  Tight band [lbT, ubT]: the positions whose
// interval can reciprocally overlap p by >= T_p (cluster.py:157): above p, start <= end_p - T_p (inside [p, ub]); below p,
// nothing before the first position whose prefix-max end reaches start_p + T_p.
__global__ void k_bands(int D, const int4 *__restrict__ SR0, const int *__restrict__ s_m, const int *__restrict__ s_chrom,
                        const int *__restrict__ pmaxS, const int *__restrict__ chrom_lo, const int *__restrict__ chrom_hi,
                        int4 *RM, unsigned long long *band_pairs, unsigned long long *tight_pairs) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    long long mine = 0, mineT = 0;
    if (p < D) {
        const int4 me = SR0[p];
        const int c = s_chrom[p];
        const int e = me.y, lim = chrom_hi[c], clo = chrom_lo[c];
        int lo = p, step = 1;                                        // invariant: start[lo] <= e
        while (lo + step < lim && SR0[lo + step].x <= e) { lo += step; step <<= 1; }
        int hi = min(lo + step, lim);                                // start[hi] > e, or hi == lim
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (SR0[mid].x <= e) lo = mid; else hi = mid; }
        const long long et = (long long)e - (long long)me.z;        // T = 0 (overlap <= 0): the closed band
        int tl = p, th = lo + 1;                                     // start[tl] <= et or tl == p; start[th] > et or th == ub + 1
        while (th - tl > 1) { int mid = (tl + th) >> 1; if ((long long)SR0[mid].x <= et) tl = mid; else th = mid; }
        const long long st = (long long)me.x + (long long)me.z;
        int lb = p;
        if (p > clo && (long long)pmaxS[p - 1] >= st) {
            lb = p - 1;
            int stp = 1;                                             // invariant: pmaxS[lb] >= st
            while (lb - stp >= clo && (long long)pmaxS[lb - stp] >= st) { lb -= stp; stp <<= 1; }
            int l2 = max(lb - stp, clo - 1);                         // pmaxS[l2] < st, or l2 == clo - 1
            while (lb - l2 > 1) { int mid = (l2 + lb) >> 1; if ((long long)pmaxS[mid] >= st) lb = mid; else l2 = mid; }
        }
        const int m = s_m[p];
        RM[2 * m] = make_int4(c, me.x, me.y, me.z);
        RM[2 * m + 1] = make_int4(p, lo, lb, tl);
        mine = lo - p;
        mineT = tl - lb;
    }
    __shared__ long long s_sum[2][8];                                 // one pair of global atomics per block
#pragma unroll
    for (int o = 16; o; o >>= 1) { mine += __shfl_down_sync(0xffffffffu, mine, o); mineT += __shfl_down_sync(0xffffffffu, mineT, o); }
    if ((threadIdx.x & 31) == 0) { s_sum[0][threadIdx.x >> 5] = mine; s_sum[1][threadIdx.x >> 5] = mineT; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0, b = 0;
        for (int w = 0; w < 8; w++) { a += s_sum[0][w]; b += s_sum[1][w]; }
        if (a) atomicAdd(band_pairs, (unsigned long long)a);
        if (b) atomicAdd(tight_pairs, (unsigned long long)b);
    }
}

// ---------------------------------------------------------------- pair-level pieces
struct Tab {                 // kernel-side view of the tables
    const int4 *SR0, *SR1, *RM, *RI;
    const int *pmaxS, *chrom_lo, *chrom_hi;
    const int *sib;          // per sorted position: position of the read's next filling (cyclic) | (L - 1) << 26; WALK replay only
    int D, Q, Tedge;
};
// per-N Jaccard cutoff as the largest passing union: a kernel parameter of its own (per call, so that concurrent contexts
// with different options never share it), staged into shared memory by the kernels that index it
struct UmaxTab { int v[LMAX + 1]; };
// read-major filling records (see k_bands)
__device__ __forceinline__ int4 rm0(const Tab &t, int m) { return __ldg(&t.RM[2 * m]); }                            // {chrom, start, end, T}
__device__ __forceinline__ int2 rm1(const Tab &t, int m) { return __ldg((const int2 *)&t.RM[2 * m + 1]); }          // {pos, ub}
__device__ __forceinline__ int2 rm2(const Tab &t, int m) { return __ldg((const int2 *)&t.RM[2 * m + 1] + 1); }      // {lbT, ubT}

// a (query) against b, both as read-major records (RM + 2 * off, stride 2): greedy first-fit count of cluster.py:152-161
// plus the lexicographically first matching filling pair
__device__ __forceinline__ int greedy_ab(const int4 *__restrict__ A, int La, const int4 *__restrict__ B, int Lb, int *first_fa, int *first_fb) {
    unsigned long long used = 0;
    int n = 0, ffa = -1, ffb = -1;
    for (int fa = 0; fa < La; fa++) {
        int4 a = __ldg(&A[2 * fa]);
        for (int fb = 0; fb < Lb; fb++) {
            int4 b = __ldg(&B[2 * fb]);
            int ov = min(a.z, b.z) - max(a.y, b.y);
            bool m = (a.x == b.x) && (max(ov, 0) >= max(a.w, b.w));
            if (m) {
                if (ffa < 0) { ffa = fa; ffb = fb; }
                if (!((used >> fb) & 1ull)) { used |= 1ull << fb; n++; break; }
            }
        }
    }
    *first_fa = ffa; *first_fb = ffb;
    return n;
}
__device__ __forceinline__ bool difflen_ok(int qa, int Lqa, int nla, int qb, int Lqb, int nlb) {
    bool q_ok = min(qa, qb) >= max(Lqa, Lqb);
    bool n_ok = min(nla & 0xffff, nlb & 0xffff) >= max((nla >> 16) & 0xffff, (nlb >> 16) & 0xffff);
    return q_ok || n_ok;                                           // cluster.py:178-183 (skip only if both fail)
}

// ---------------------------------------------------------------- stage 6: read-major pair kernel (order-free relation)
// A GROUP of 8 lanes owns one query read a (4 reads per warp, consecutive query ranks = usually one PCR family, so the
// groups of a warp run in step).  For every filling of a the group walks the filling's TIGHT band [lbT, ubT] of sorted
// interval records (the only positions whose interval can reciprocally overlap it by >= --overlap, cluster.py:157) with
// coalesced int4 loads, 8 positions per step.  A hit names a partner read b; the lane gathers b's filling list and
// evaluates a -> b once: different_lengths_or_alignments (cluster.py:178-183), the greedy N-1 intersection
// (cluster.py:152-161) and the per-N Jaccard cutoff (cluster.py:165-170,218-219).  "Once" = at the lexicographically
// first matching filling pair of (a, b); a small per-group hash of partners already settled filters the later hits
// before any gather (a filter only: a miss costs a re-evaluation that the canonical-pair rule then discards).
// Passing pairs are appended to the relation list through warp-aggregated chunk reservations; a read stops as soon as
// edge_threshold partners passed (it is saturating: replayed in query order later, its entries are ignored).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool match4(const int4 a, const int4 b) {
    return (a.x == b.x) && (max(min(a.z, b.z) - max(a.y, b.y), 0) >= max(a.w, b.w));
}
// cluster.py:157 for fillings {chrom, start, end, T}: with --overlap > 0 every T >= 1, so max(ov, 0) >= T <=> ov >= T;
// ALLMATCH (--overlap <= 0, all T = 0): any two fillings on one chromosome match
template <bool ALLMATCH>
__device__ __forceinline__ bool matchT(const int4 a, const int4 b) {
    if (ALLMATCH) return a.x == b.x;
    return (a.x == b.x) && ((min(a.z, b.z) - max(a.y, b.y)) >= max(a.w, b.w));
}
// a -> b for reads with <= 4 fillings, lists in registers.  Returns bit0: evaluated, bit1: (fia, fbp) is the canonical
// (lexicographically first) band hit of the pair.  *n_out = greedy intersection (cluster.py:152-161).
template <bool ALLMATCH>
__device__ __forceinline__ int eval_small(const int4 *A, int La, const int4 *__restrict__ B, int Lb, int fia, int fbp, int *n_out) {
    int4 a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        a[k] = k < La ? A[k] : make_int4(-1, 0, 0, 0x7fffffff);
        b[k] = k < Lb ? __ldg(&B[2 * k]) : make_int4(-2, 0, 0, 0x7fffffff);   // B: read-major records, stride 2
    }
    unsigned m[4], h[4];
#pragma unroll
    for (int fa = 0; fa < 4; fa++) {
        unsigned r = 0, hr = 0;
#pragma unroll
        for (int fb = 0; fb < 4; fb++) {
            r |= (matchT<ALLMATCH>(a[fa], b[fb]) ? 1u : 0u) << fb;
            if (ALLMATCH) hr |= ((a[fa].x == b[fb].x && min(a[fa].z, b[fb].z) - max(a[fa].y, b[fb].y) >= 0) ? 1u : 0u) << fb;
        }
        m[fa] = r; h[fa] = ALLMATCH ? hr : r;                        // h: matching pairs that are band hits (closed overlap)
    }
    unsigned used = 0;
    int n = 0, ffa = -1, ffb = -1;
#pragma unroll
    for (int fa = 0; fa < 4; fa++) {
        if (h[fa] && ffa < 0) { ffa = fa; ffb = __ffs(h[fa]) - 1; }
        const unsigned avail = m[fa] & ~used;
        if (avail) { used |= avail & (0u - avail); n++; }
    }
    *n_out = n;
    return 1 | ((ffa == fia && ffb == fbp) ? 2 : 0);
}
template <bool ALLMATCH>
__device__ __noinline__ int eval_general(const int4 *__restrict__ A, int La, const int4 *__restrict__ B, int Lb, int fia, int fbp, int *n_out) {
    unsigned long long used = 0;
    int n = 0, ffa = -1, ffb = -1;
    for (int fa = 0; fa < La; fa++) {
        const int4 a = __ldg(&A[2 * fa]);                               // A, B: read-major records, stride 2
        bool taken = false;
        for (int fb = 0; fb < Lb; fb++) {
            const int4 b = __ldg(&B[2 * fb]);
            if (matchT<ALLMATCH>(a, b)) {
                if (ffa < 0 && (!ALLMATCH || min(a.z, b.z) - max(a.y, b.y) >= 0)) { ffa = fa; ffb = fb; }
                if (!taken && !((used >> fb) & 1ull)) { used |= 1ull << fb; n++; taken = true; if (ffa >= 0) break; }
            }
        }
    }
    *n_out = n;
    return 1 | ((ffa == fia && ffb == fbp) ? 2 : 0);
}

#define PK_WARPS 8
#define PK_GROUPS (PK_WARPS * 4)
#define PK_HASH 64              // settled-partner filter slots per group
#define PK_CHUNK 256            // relation-entry slots a warp reserves at a time (>= 32)
#define RP_K 64                 // partners a saturating read may have for the replay's LIST mode
#define PL_CHUNK 512            // partner records a warp reserves at a time (>= 4 * RP_K)
#define RP_KL 4                  // partners per lane of a replay group handled in one batch (32 partners per batch)

// Partner record of a saturating read a (replay LIST mode): everything the replay needs to know about partner b without
// touching b's geometry again.  r0 = {b | edge << 31, off_b << 6 | L_b - 1, cg, 0}, r1 = {key[0..3]}:
//   edge    a -> b passes the Jaccard cutoff (cluster.py:218-219),
//   cg      nibble g: 4 | fa* when filling g of b overlaps (closed intervals) a filling of a, fa* = the overlapped filling of
//           a with the highest sorted position: b's scan of g saw a iff it got down to that position,
//   key[fa] the highest sorted position of an interval of b inside the closed band of a's filling fa (-1: none): where
//           a's scan of fa first meets b.
struct PLInfo { unsigned long long off; int n; int pad; };

template <bool ALLMATCH>
__global__ void __launch_bounds__(PK_WARPS * 32) k_pair(Tab t, const UmaxTab um, int shard, int nshard, int lists_only, int *isP, int2 *entries,
                                                         unsigned long long *n_slots, unsigned long long cap_entries,
                                                         int4 *PL, PLInfo *plinfo, unsigned long long *pl_slots, unsigned long long cap_pl,
                                                         unsigned long long *n_tests, unsigned long long *n_real, int *err) {
    __shared__ int4 sA[PK_GROUPS][4];
    __shared__ int4 sB[PK_GROUPS][4];                                              // {lbT, ubT, pos, ub} of a's fillings
    __shared__ int2 sHash[PK_GROUPS][PK_HASH];
    __shared__ int2 sPart[PK_GROUPS][RP_K];                                        // {b | edge << 31, off_b << 6 | L_b - 1}
    __shared__ int s_umax[LMAX + 1];
    for (int k = threadIdx.x; k <= LMAX; k += blockDim.x) s_umax[k] = um.v[k];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gl = lane & 7, g = lane >> 3, grp = w * 4 + g;
    const unsigned ltmask = (1u << lane) - 1u, gmask = 0xffu << (g * 8);
    unsigned long long tests = 0, real = 0, chunk_base = 0, pl_base = 0, nrec_total = 0;
    int chunk_used = PK_CHUNK, pl_used = PL_CHUNK;                                 // nothing reserved yet
    for (int k = gl; k < PK_HASH; k += 8) sHash[grp][k] = make_int2(-1, -1);
    const int stride = gridDim.x * PK_GROUPS;
    for (int q0 = blockIdx.x * PK_GROUPS; q0 < t.Q; q0 += stride) {                // block-uniform trip count
        const int q = q0 + grp;
        const bool mine = nshard <= 1 || ((q >> 8) % nshard) == shard;             // 256-read groups, round robin over ranks
        // pass 0: the reads of this shard; pass 1 (multi-GPU, after the exchange of isP): partner lists of the saturating
        // reads the other shards own
        const bool live = q < t.Q && (lists_only ? (!mine && __ldg(&isP[q]) != 0) : mine);
        int4 ri = make_int4(0, 0, 0, 0);
        if (live) ri = __ldg(&t.RI[q]);
        const int off = (int)((unsigned)ri.w >> 6), La = live ? (ri.w & 63) + 1 : 0;
        __syncwarp();
        if (gl < min(La, 4)) {
            sA[grp][gl] = rm0(t, off + gl);
            const int2 pu = rm1(t, off + gl), bd = rm2(t, off + gl);
            sB[grp][gl] = make_int4(bd.x, bd.y, pu.x, pu.y);
        }
        __syncwarp();
        int cnt = lists_only ? t.Tedge : 0;                                        // passing partners so far
        int nPart = 0;                                                             // partners buffered for the replay; -1: too many / too long
        const int maxLa = __reduce_max_sync(FULL, La);
        for (int fi = 0; fi < maxLa; fi++) {
            // a read keeps scanning while it may still be non-saturating (its entries must be complete) or while its
            // partner list is still within bounds (the replay wants all of it)
            bool fact = fi < La && (cnt < t.Tedge || nPart >= 0);
            int4 f = make_int4(0, 0, 0, 0);
            int2 band = make_int2(1, 0);
            if (fact) {
                if (La <= 4) { f = sA[grp][fi]; band = make_int2(sB[grp][fi].x, sB[grp][fi].y); }
                else { f = rm0(t, off + fi); band = rm2(t, off + fi); nPart = -1; fact = cnt < t.Tedge; }
            }
            for (int ch = 0;; ch++) {
                const int p = band.x + ch * 8 + gl;
                const bool v = fact && (cnt < t.Tedge || nPart >= 0) && p <= band.y;
                if (!__any_sync(FULL, v)) break;                                    // every group of the warp is through its band
                int4 c0 = make_int4(0, 0, 0x7fffffff, -1);
                if (v) c0 = __ldg(&t.SR0[p]);
                const int b = c0.w & QMASK;
                bool pass = false, part = false, longb = false;
                int wb = 0;
                if (v && b != q && (min(f.z, c0.y) - max(f.y, c0.x)) >= max(f.w, c0.z)) {   // cluster.py:157 for this interval pair
                    int2 *hs = &sHash[grp][b & (PK_HASH - 1)];
                    const int2 hv = *hs;
                    if (hv.x != b || hv.y != q) {                                   // not settled earlier in this read's pass
                        const int4 c1 = __ldg(&t.SR1[p]);
                        bool settled = true;
                        if (difflen_ok(ri.x, ri.y, ri.z, c1.x, c1.y, c1.z)) {
                            wb = c1.w;
                            const int offb = (int)((unsigned)wb >> 6), Lb = (wb & 63) + 1;
                            const int fbp = (int)((unsigned)c0.w >> 26);
                            int n, fl;
                            if (La <= 4 && Lb <= 4) fl = eval_small<ALLMATCH>(sA[grp], La, t.RM + 2 * offb, Lb, fi, fbp, &n);
                            else { fl = eval_general<ALLMATCH>(t.RM + 2 * off, La, t.RM + 2 * offb, Lb, fi, fbp, &n); longb = true; }
                            settled = (fl & 2) != 0;
                            if (settled) {
                                tests++;
                                part = n > 0;                                       // the pair can be an effective candidate (cluster.py:216)
                                pass = n > 0 && (La + Lb - n) <= s_umax[n];         // cluster.py:165-170,218-219
                            }
                        }
                        if (settled) *hs = make_int2(b, q);
                    }
                }
                // ---- partner buffer (only used if the read turns out saturating)
                const unsigned am = __ballot_sync(FULL, part) & gmask, lm = __ballot_sync(FULL, longb) & gmask;
                if (am) {
                    const int na = __popc(am);
                    if (ALLMATCH || nPart < 0 || lm || nPart + na > RP_K) nPart = -1;
                    else {
                        if (part) sPart[grp][nPart + __popc(am & ltmask)] = make_int2((int)((unsigned)b | (pass ? 0x80000000u : 0u)), wb);
                        nPart += na;
                    }
                }
                // ---- relation entries of reads still below the threshold: warp-aggregated append
                pass = pass && !lists_only && cnt < t.Tedge;
                const unsigned pm = __ballot_sync(FULL, pass);
                if (pm) {
                    const int n = __popc(pm);
                    if (chunk_used + n > PK_CHUNK) {
                        for (int k = chunk_used + lane; k < PK_CHUNK; k += 32) entries[chunk_base + k] = make_int2(-1, -1);
                        if (lane == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)PK_CHUNK);
                        chunk_base = __shfl_sync(FULL, chunk_base, 0);
                        chunk_used = 0;
                        if (chunk_base + PK_CHUNK > cap_entries) { if (lane == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                    }
                    if (pass) entries[chunk_base + chunk_used + __popc(pm & ltmask)] = make_int2(q, b);
                    chunk_used += n;
                    real += (lane == 0) ? n : 0;
                    cnt += __popc(pm & gmask);
                }
                __syncwarp();                                                       // filter updates visible to the next step
            }
        }
        // ---- saturating read: publish its partner records for the replay
        const bool sat = live && cnt >= t.Tedge;
        if (live && gl == 0 && !lists_only) isP[q] = sat;
        const int nrec = (sat && nPart > 0) ? nPart : 0;
        int tot = nrec;                                                             // records of the warp's 4 groups
        tot = __shfl_sync(FULL, tot, 0) + __shfl_sync(FULL, tot, 8) + __shfl_sync(FULL, tot, 16) + __shfl_sync(FULL, tot, 24);
        if (tot) {
            if (pl_used + tot > PL_CHUNK) {
                if (lane == 0) pl_base = atomicAdd(pl_slots, (unsigned long long)PL_CHUNK);
                pl_base = __shfl_sync(FULL, pl_base, 0);
                pl_used = 0;
                if (pl_base + PL_CHUNK > cap_pl) { if (lane == 0) atomicOr(err, EF_OVERFLOW); pl_base = 0; }
            }
            int before = 0;                                                         // records of the lower groups
            for (int gg = 0; gg < 3; gg++) { const int x = __shfl_sync(FULL, nrec, gg * 8); if (gg < g) before += x; }
            const unsigned long long my0 = pl_base + pl_used + before;
            pl_used += tot;
            nrec_total += tot;                                                       // (statistics: records written)
            for (int j = gl; j < nrec; j += 8) {
                const int2 pr = sPart[grp][j];
                const int offb = (int)((unsigned)pr.y >> 6), Lb = (pr.y & 63) + 1;
                int key[4] = {-1, -1, -1, -1};
                unsigned cg = 0;
                for (int gb = 0; gb < Lb; gb++) {
                    const int4 i0 = rm0(t, offb + gb);
                    const int pg = rm1(t, offb + gb).x;
                    int best = -1, bestfa = 0;
#pragma unroll
                    for (int fa = 0; fa < 4; fa++) {
                        const int4 af = sA[grp][fa];
                        if (fa < La && af.x == i0.x && af.y <= i0.z && af.z >= i0.y) {   // closed overlap: a scan of one visits the other
                            key[fa] = max(key[fa], pg);
                            if (sB[grp][fa].z > best) { best = sB[grp][fa].z; bestfa = fa; }
                        }
                    }
                    if (best >= 0) cg |= (4u | (unsigned)bestfa) << (4 * gb);
                }
                PL[2 * (my0 + j)] = make_int4(pr.x, pr.y, (int)cg, 0);
                PL[2 * (my0 + j) + 1] = make_int4(key[0], key[1], key[2], key[3]);
            }
            if (sat && gl == 0) { PLInfo pi; pi.off = my0; pi.n = nPart; pi.pad = 0; plinfo[q] = pi; }
        }
        if (sat && nPart <= 0 && gl == 0) {
            PLInfo pi; pi.off = 0; pi.n = nPart < 0 ? -1 : 0; pi.pad = 0; plinfo[q] = pi;
            if (nPart < 0) atomicAdd(pl_slots + 3, 1ull);                           // (reads the replay has to WALK)
        }
    }
    if (chunk_used < PK_CHUNK)
        for (int k = chunk_used + lane; k < PK_CHUNK; k += 32) entries[chunk_base + k] = make_int2(-1, -1);
    for (int o = 16; o; o >>= 1) { tests += __shfl_down_sync(FULL, tests, o); real += __shfl_down_sync(FULL, real, o); }
    if (lane == 0) { if (tests) atomicAdd(n_tests, tests); if (real) atomicAdd(n_real, real); if (nrec_total) atomicAdd(pl_slots + 2, nrec_total); }
}

// ---------------------------------------------------------------- stage 7: saturating set
__global__ void k_compact_flagged(int n, const int *__restrict__ flag, const int *__restrict__ pos, int *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}

// ---------------------------------------------------------------- stage 8: replay of saturating reads in query order
// plist holds the saturating reads in ascending query rank, cut into RUNS of reads that depend on each other (consecutive
// ranks of one PCR family, k_run_flags).  A GROUP of 8 lanes takes the next run (ticket) and walks its reads back to back,
// re-running each read's query exactly as cluster.py:197-224 would: per filling, the closed band is walked downwards from
// ub, 8 sorted positions per step; pairs already seen are skipped, edges counted, and the scan breaks at edge_threshold.
// All a later query can observe of this is one integer per filling — the position where the scan stopped — published in
// stop[] (-1 until known).  Whether an earlier-ranked saturating read b "saw" the pair first is a function of b's stops.
//
// The kernel is a non-blocking state machine: the 4 groups of a warp advance one step per loop iteration in lock step;
// a step whose outcome depends on a stop that is not published yet commits only the candidates before it (scan order)
// and is retried on the next iteration — nobody spins, so groups can never block one another, and a group only ever
// depends on reads of smaller tickets (held by resident groups) or on earlier reads of its own run.
#define RG_WARPS 2
#define RG_GROUPS (RG_WARPS * 4)
#define RUN_CAP 64
#define RUN_LONG 1
#define RP_CHUNK 64             // edge slots a group reserves at a time (>= 8)
enum { RF_TESTED = 1, RF_REACH = 2, RF_EDGE = 4, RF_UNRES = 8 };
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// stop words: >= 0 final stop position of a filling's scan; < 0 while unknown: -2 - x means "every candidate at a
// position >= x has been visited already" (progress of a long walk), STOP_UNSTARTED = nothing known yet.
#define STOP_UNSTARTED ((int)0x80000000)
__device__ __forceinline__ int stop_reached(int v) { return v >= 0 ? v : -2 - v; }   // lowest position known to be visited
// streak starts: ticket k continues the previous saturating read's streak iff their first fillings reciprocally overlap
// (same PCR family: they depend on each other).  Also marks the stops of every saturating read as unknown.
__global__ void k_run_flags(int nP, const int *__restrict__ plist, const int4 *__restrict__ RI, const int4 *__restrict__ RM, int *flag,
                            int *stop, int *stopS) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    const int w = RI[plist[k]].w;
    const int off = (int)((unsigned)w >> 6), L = (w & 63) + 1;
    for (int j = 0; j < L; j++) { stop[off + j] = STOP_UNSTARTED; stopS[RM[2 * (off + j) + 1].x] = STOP_UNSTARTED; }
    int f = 1;
    if (k > 0) {
        const int4 x = RM[2 * ((unsigned)RI[plist[k - 1]].w >> 6)], y = RM[2 * off];
        const int ov = min(x.z, y.z) - max(x.y, y.y);
        if (x.x == y.x && max(ov, 0) >= max(x.w, y.w)) f = 0;
    }
    flag[k] = f;
}
// runs: a streak of up to RUN_CAP reads is one run (one group walks it back to back: its reads wait on each other
// anyway); a longer streak is a giant clique whose reads mostly do NOT depend on each other — cut it into runs of RUN_LONG
__global__ void k_run_cut(int nP, const int *__restrict__ sflag, const int *__restrict__ spos, const int *__restrict__ sstart,
                          const int64_t *__restrict__ n_streaks, int *rflag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    const int sid = spos[k] + sflag[k] - 1;
    const int start = sstart[sid], end = (sid + 1 < (int)*n_streaks) ? sstart[sid + 1] : nP;
    const int cap = (end - start) > RUN_CAP ? RUN_LONG : RUN_CAP;
    rflag[k] = ((k - start) % cap) == 0;
}
// sib[p]: where the next filling (cyclic) of p's read sits in sorted order, and the read's filling count
__global__ void k_sib(int D, const int4 *__restrict__ SR0, const int4 *__restrict__ SR1, const int4 *__restrict__ RM, int *sib) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D) return;
    const int w = SR1[p].w;
    const int off = (int)((unsigned)w >> 6), L = (w & 63) + 1, fi = (int)((unsigned)SR0[p].w >> 26);
    const int nxt = off + (fi + 1 == L ? 0 : fi + 1);
    sib[p] = (int)((unsigned)RM[2 * nxt + 1].x | ((unsigned)(L - 1) << 26));
}
// Did read b (owner of the interval at sorted position p, b < a) provably see a first through one of its OTHER fillings?
// Walks b's fillings through the position-indexed sibling ring: sibling at sp saw a's filling fa iff they overlap (closed)
// and b's scan of the sibling got down to fa's position.  A's fillings (<= 4) are in shared memory.
__device__ __forceinline__ bool seen_via_sibling(const Tab &t, const int *stopS, int p, const int4 *A0, const int2 *A1, const int2 *Achr, int La) {
    int sv = __ldg(&t.sib[p]);
    const int hops = (int)((unsigned)sv >> 26);                                    // L - 1 other fillings
    if (hops > 3) return false;
    for (int h = 0; h < hops; h++) {
        const int sp = sv & QMASK;
        const int4 c = __ldg(&t.SR0[sp]);
        const int reached = stop_reached(ld_relaxed(&stopS[sp]));
#pragma unroll
        for (int fa = 0; fa < 4; fa++)
            if (fa < La && sp >= Achr[fa].x && sp < Achr[fa].y && A0[fa].y <= c.y && A0[fa].z >= c.x && reached <= A1[fa].x) return true;
        sv = __ldg(&t.sib[sp]);
    }
    return false;
}
// one candidate b of read a's filling scan when either read has more than 4 fillings (lists stay in global memory)
__device__ __noinline__ int replay_eval_general(const Tab &t, const int *umax, const int *stop, const int *ownStop, int a, int offa, int La, int fi,
                                                const int4 f, int top, int p, int b, int offb, int Lb) {
    for (int g = 0; g < Lb; g++) {                                                 // pair already seen earlier in this very query?
        const int4 bg = rm0(t, offb + g);
        const int pg = rm1(t, offb + g).x;
        for (int f2 = 0; f2 < fi; f2++) {
            const int4 af = rm0(t, offa + f2);
            if (af.x == bg.x && ownStop[f2] <= pg && pg <= rm1(t, offa + f2).y && bg.z >= af.y) return 0;
        }
        if (bg.x == f.x && pg > p && pg <= top && bg.z >= f.y) return 0;
    }
    int ffa, ffb;
    const int n = greedy_ab(t.RM + 2 * offa, La, t.RM + 2 * offb, Lb, &ffa, &ffb);
    if (n == 0) return RF_TESTED;
    if (b < a) {                                                                   // b queried first: did its scans get here?
        bool vis = false, unres = false;
        for (int g = 0; g < Lb; g++) {
            const int4 bf = rm0(t, offb + g);
            const int ubf = rm1(t, offb + g).y;
            const int sf = ld_relaxed(&stop[offb + g]);
            for (int fa = 0; fa < La; fa++) {
                const int4 ag = rm0(t, offa + fa);
                const int pa = rm1(t, offa + fa).x;
                if (ag.x == bf.x && pa <= ubf && ag.z >= bf.y) { if (stop_reached(sf) <= pa) vis = true; else if (sf < 0) unres = true; }
            }
        }
        if (vis) return RF_TESTED;
        if (unres) return RF_TESTED | RF_UNRES;
    }
    return RF_TESTED | RF_REACH | ((La + Lb - n) <= umax[n] ? RF_EDGE : 0);
}
// Two ways to re-run one read's query:
//   LIST mode (the normal case): the only candidates that can ever matter to a's query are intervals of reads b that share
//     a reciprocally overlapping filling pair with a (n_i > 0, cluster.py:216) — everything else is skipped by the reference
//     before it touches `edges` or the break.  The pair kernel already met and evaluated all of them and left one record
//     per partner (<= RP_K): the group loads the records and replays every filling's scan over the partners only.  A
//     partner is first met at its highest interval inside the filling's closed band (key), scan order = descending key,
//     and the break position follows from a selection over the keys — one step per filling, however long the band is.
//   WALK mode (reads with too many partners, e.g. a 500k-read hotspot, more than 4 fillings, or --overlap <= 0): the
//     closed band is walked downwards 8 sorted positions per step and every candidate is evaluated; candidates whose own
//     scan of that very interval already passed a's filling are skipped on two coalesced loads, 64 positions per step.
static_assert(RP_K <= RP_CHUNK && 4 * RP_K <= PL_CHUNK, "chunk sizes");
// group-wide max over the 8 lanes of a group
__device__ __forceinline__ int gmax8(unsigned gmask, int v) {
    v = max(v, __shfl_xor_sync(gmask, v, 1)); v = max(v, __shfl_xor_sync(gmask, v, 2)); v = max(v, __shfl_xor_sync(gmask, v, 4));
    return v;
}
// WALK = false: an instantiation without the band-walking code (half the registers, twice the resident groups) for the
// usual case that every saturating read has partner records
template <bool ALLMATCH, bool WALK>
__global__ void __launch_bounds__(RG_WARPS * 32, 8) k_replay(Tab t, const UmaxTab um, int nP, const int *__restrict__ plist, int nRuns,
                                                           const int *__restrict__ rstart, const int *__restrict__ isP,
                                                           const int4 *__restrict__ PL, const PLInfo *__restrict__ plinfo, int *stop,
                                                           int *stopS, unsigned *ticket, int2 *pedges, unsigned long long *n_slots,
                                                           unsigned long long cap_pedges, unsigned long long *n_tests, int *err,
                                                           unsigned long long *dbg) {
    __shared__ int4 sA0[RG_GROUPS][4];
    __shared__ int2 sA1[RG_GROUPS][4];
    __shared__ int sStop[RG_GROUPS][LMAX];
    __shared__ int4 sP0[RG_GROUPS][RP_K];    // partner records: {b | edge << 31, off_b << 6 | L_b - 1, cg, flags}; flags: 1 visited by a,
    __shared__ int4 sP1[RG_GROUPS][RP_K];    //   4 b saw a first, 8 b did not;  {key[0..3]}
    __shared__ int sKey[RG_GROUPS][RP_K];    // first-visit position in the current filling's scan
    __shared__ int2 sAchr[RG_GROUPS][4];     // [chrom_lo, chrom_hi) of a's fillings (sibling test of the WALK mode)
    __shared__ int s_umax[LMAX + 1];
    __shared__ int sRecTag[RG_GROUPS][32];   // the reads this group replayed last (direct mapped by rank & 31) and their final
    __shared__ int4 sRecStop[RG_GROUPS][32]; //   stops: partners of one run mostly look each other up here, not in global memory
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gl = lane & 7, gsh = lane & 24, grp = w * 4 + (lane >> 3);
    const unsigned gmask = 0xffu << gsh;
    // group state (identical in the 8 lanes of a group)
    int phase = 0;                      // 0 next read, 1 walk: next filling, 2 walk: scanning, 3 finished, 4 list: build, 5 list: filling
    unsigned tk = 0, tk1 = 0;
    int a = 0, offa = 0, La = 0, fi = 0, edges = 0, top = 0, lo = 0, base = 0, posf = 0;
    int nPart = 0;
    bool wide = false;
    int4 ria = make_int4(0, 0, 0, 0), f = make_int4(0, 0, 0, 0);
    unsigned long long tests = 0, chunk_base = 0;
    unsigned long long d_iter = 0, d_steps = 0, d_stall = 0, d_sleep = 0;
    int d_fsteps = 0, d_fstall = 0;
    int chunk_used = RP_CHUNK;
    for (int k = threadIdx.x; k <= LMAX; k += blockDim.x) s_umax[k] = um.v[k];
    __syncthreads();
    for (int k = gl; k < 32; k += 8) sRecTag[grp][k] = -1;
    for (;;) {
        __syncwarp();
        d_iter += lane == 0;
        if (phase == 0) {
            if (tk == tk1) {                                                       // run finished: take the next ticket
                unsigned run = 0;
                if (gl == 0) run = atomicAdd(ticket, 1u);
                run = __shfl_sync(gmask, run, gsh);
                if (run >= (unsigned)nRuns) phase = 3;
                else {
                    tk = (unsigned)__ldg(&rstart[run]);
                    tk1 = (run + 1 < (unsigned)nRuns) ? (unsigned)__ldg(&rstart[run + 1]) : (unsigned)nP;
                }
            }
            if (phase == 0) {
                a = __ldg(&plist[tk]);
                ria = __ldg(&t.RI[a]);                                             // {qlen2, Lq, naln | Ln << 16, off << 6 | L - 1}
                offa = (int)((unsigned)ria.w >> 6); La = (ria.w & 63) + 1;
                fi = 0; edges = 0;
                if (La <= 4 && gl < La) {
                    const int4 r0 = rm0(t, offa + gl);
                    sA0[grp][gl] = r0; sA1[grp][gl] = rm1(t, offa + gl);
                    if (WALK) sAchr[grp][gl] = make_int2(__ldg(&t.chrom_lo[r0.x]), __ldg(&t.chrom_hi[r0.x]));
                }
                phase = 1;
                const PLInfo pi = plinfo[a];
                if (!WALK && pi.n < 0) { atomicOr(err, EF_OVERFLOW); phase = 3; }  // (cannot happen: the host picks WALK when such reads exist)
                if (gl == 0) { sRecTag[grp][a & 31] = La <= 4 ? a : -1; sRecStop[grp][a & 31] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff); }
                if (!ALLMATCH && pi.n >= 0) {                                      // the pair kernel left a's partner records
                    nPart = pi.n;
                    for (int jb = 0; jb < nPart; jb += 8 * RP_KL) {
                        int4 r0[RP_KL], r1[RP_KL];
#pragma unroll
                        for (int k = 0; k < RP_KL; k++) {
                            const int j = jb + gl + 8 * k;
                            if (j < nPart) { r0[k] = __ldg(&PL[2 * (pi.off + j)]); r1[k] = __ldg(&PL[2 * (pi.off + j) + 1]); }
                        }
#pragma unroll
                        for (int k = 0; k < RP_KL; k++) {
                            const int j = jb + gl + 8 * k;
                            if (j < nPart) {
                                const int b = r0[k].x & QMASK;
                                r0[k].w = (b < a && !__ldg(&isP[b])) ? 1 : 0;      // b < a and never breaking: it saw the pair
                                sP0[grp][j] = r0[k];
                                sP1[grp][j] = r1[k];
                            }
                        }
                    }
                    phase = 5;
                }
            }
        }
        if (__all_sync(FULL, phase == 3)) break;
        __syncwarp();
        bool stalled = false;
        // ------------------------------------------------------------ LIST mode: one filling's scan over the partners
        if (phase == 5) {
            if (La <= 4) { f = sA0[grp][fi]; top = sA1[grp][fi].y; posf = sA1[grp][fi].x; }
            else { f = rm0(t, offa + fi); const int2 pu = rm1(t, offa + fi); posf = pu.x; top = pu.y; }
            lo = __ldg(&t.chrom_lo[f.x]);
            // pass 1: where does the scan first meet each partner (its highest interval inside the closed band); did an
            // earlier-ranked partner's own query see a first?  cls: 0 not met, 1 seen, 2 reach, 3 reach + edge, 4 undecided
            int mxReach = -1, mxUn = -1, nEdge = 0;
            for (int jb = 0; jb < nPart; jb += 8 * RP_KL) {                        // 32 partners per batch (usually one batch)
            int4 q0[RP_KL];
            int keyk[RP_KL], sv[RP_KL][4];
            bool poll[RP_KL];
#pragma unroll
            for (int k = 0; k < RP_KL; k++) {                                      // stage A: keys; who needs b's stops?
                const int j = jb + gl + 8 * k;
                keyk[k] = -1; poll[k] = false;
                q0[k] = make_int4(0, 0, 0, 1);
                if (j < nPart) {
                    q0[k] = sP0[grp][j];
                    if (!(q0[k].w & 1)) {
                        const int4 r1 = sP1[grp][j];
                        keyk[k] = fi == 0 ? r1.x : fi == 1 ? r1.y : fi == 2 ? r1.z : r1.w;
                        poll[k] = keyk[k] >= 0 && (q0[k].x & QMASK) < a && !(q0[k].w & 12);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RP_KL; k++) {                                      // stage B: all the loads, back to back
                if (poll[k]) {
                    const int b = q0[k].x & QMASK;
                    if (sRecTag[grp][b & 31] == b) {                               // replayed by this very group a moment ago
                        const int4 c = sRecStop[grp][b & 31];
                        sv[k][0] = c.x; sv[k][1] = c.y; sv[k][2] = c.z; sv[k][3] = c.w;
                    } else {
                        const int offb = (int)((unsigned)q0[k].y >> 6), Lb = (q0[k].y & 63) + 1;
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            sv[k][g] = (g < Lb && (((unsigned)q0[k].z >> (4 * g)) & 4u)) ? ld_relaxed(&stop[offb + g]) : 0x7fffffff;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RP_KL; k++) {                                      // stage C: did b's own query see a first?
                const int j = jb + gl + 8 * k;
                if (poll[k]) {
                    bool vis = false, unres = false;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const unsigned cgg = ((unsigned)q0[k].z >> (4 * g)) & 15u;
                        if (cgg & 4u) {
                            if (stop_reached(sv[k][g]) <= sA1[grp][cgg & 3u].x) vis = true; else if (sv[k][g] < 0) unres = true;
                        }
                    }
                    if (vis) q0[k].w |= 4; else if (!unres) q0[k].w |= 8;
                    sP0[grp][j].w = q0[k].w;
                }
                if (keyk[k] >= 0) {
                    if ((q0[k].x & QMASK) > a || (q0[k].w & 8)) { mxReach = max(mxReach, keyk[k]); nEdge += (unsigned)q0[k].x >> 31; }
                    else if (!(q0[k].w & 4)) mxUn = max(mxUn, keyk[k]);
                }
                if (j < nPart) sKey[grp][j] = keyk[k];
            }
            }
            __syncwarp(gmask);
            // the break (cluster.py:223-224): the first reached partner, in scan order, at which `edges` is >= edge_threshold
            const int need = t.Tedge - edges;
            int brkkey = -1;
            if (need <= 0) brkkey = gmax8(gmask, mxReach);
            else {
                nEdge += __shfl_xor_sync(gmask, nEdge, 1); nEdge += __shfl_xor_sync(gmask, nEdge, 2); nEdge += __shfl_xor_sync(gmask, nEdge, 4);
                if (nEdge >= need) {                                               // the need-th highest edge partner
                    int thr = 0x7fffffff;
                    for (int r = 0; r < need; r++) {
                        int m = -1;
                        for (int j = gl; j < nPart; j += 8) {
                            const int key = sKey[grp][j];
                            if (key >= 0 && key < thr) {
                                const int4 r0 = sP0[grp][j];
                                if (r0.x < 0 && ((r0.x & QMASK) > a || (r0.w & 8))) m = max(m, key);
                            }
                        }
                        thr = gmax8(gmask, m);
                    }
                    brkkey = thr;
                }
            }
            const int unkey = gmax8(gmask, mxUn);
            if (gl == 0) d_steps++;
            if (unkey > brkkey) { stalled = true; if (gl == 0) d_stall++; }        // an undecided partner comes first: retry later
            else {
                int ne = 0;
                for (int j0 = 0; j0 < nPart; j0 += 8) {                            // commit: everything met at or above the break
                    const int j = j0 + gl;
                    bool emit = false;
                    if (j < nPart) {
                        const int key = sKey[grp][j];
                        if (key >= 0 && key >= brkkey) {
                            const int4 r0 = sP0[grp][j];
                            sP0[grp][j].w = r0.w | 1;                              // a's query has now seen this pair
                            emit = r0.x < 0 && ((r0.x & QMASK) > a || (r0.w & 8));
                        }
                    }
                    const unsigned em = (__ballot_sync(gmask, emit) >> gsh) & 0xffu;
                    if (em) {
                        const int n = __popc(em);
                        if (chunk_used + n > RP_CHUNK) {                           // reserve a fresh chunk, pad the old one
                            for (int k = chunk_used + gl; k < RP_CHUNK; k += 8) pedges[chunk_base + k] = make_int2(-1, -1);
                            if (gl == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)RP_CHUNK);
                            chunk_base = __shfl_sync(gmask, chunk_base, gsh);
                            chunk_used = 0;
                            if (chunk_base + RP_CHUNK > cap_pedges) { if (gl == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                        }
                        if (emit) pedges[chunk_base + chunk_used + __popc(em & ((1u << gl) - 1u))] = make_int2(a, sP0[grp][j].x & QMASK);
                        chunk_used += n;
                        ne += n;
                    }
                }
                edges += ne;
                const int stopf = brkkey >= 0 ? brkkey : lo;
                if (gl == 0) { st_relaxed(&stop[offa + fi], stopf); st_relaxed(&stopS[posf], stopf); ((int *)&sRecStop[grp][a & 31])[fi] = stopf; }
                fi++;
                if (fi == La) { tk++; phase = 0; }
            }
        }
        // ------------------------------------------------------------ WALK mode
        else if (WALK) {
        if (phase == 1) {
            if (La <= 4) { f = sA0[grp][fi]; top = sA1[grp][fi].y; posf = sA1[grp][fi].x; }
            else { f = rm0(t, offa + fi); const int2 pu = rm1(t, offa + fi); posf = pu.x; top = pu.y; }
            lo = __ldg(&t.chrom_lo[f.x]);
            base = top;
            wide = false;
            phase = 2;
        }
        if (phase == 2) {
            int stopf = -1;                                                        // >= 0: this filling's scan ended there
            unsigned Ecommit = 0;
            int b = -1;
            if (base < lo || __ldg(&t.pmaxS[base]) < f.y) stopf = lo;              // nothing at or below base overlaps the filling
            else if (wide) {
                // ---- nothing to do in the last step: skip ahead over candidates that are no candidates at all or whose read
                // provably saw a first (its scan of this very interval already passed a's filling), 64 positions per step
                int adv = 64;
                int wq[8], we[8], ws[8];                                           // all 16 loads of the step are issued before any use
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int p = base - 8 * k - gl;
                    wq[k] = -1; we[k] = 0; ws[k] = 0;
                    if (p >= lo) { const int4 c0 = __ldg(&t.SR0[p]); wq[k] = c0.w & QMASK; we[k] = c0.y; ws[k] = ld_relaxed(&stopS[p]); }
                }
                bool needs[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    needs[k] = wq[k] >= 0 && wq[k] != a && we[k] >= f.y && !(wq[k] < a && stop_reached(ws[k]) <= posf);
                if (t.sib && La <= 4) {                                            // ... or through its other filling (reads of 2 fillings:
                    int sp[8];                                                     //     the loads of all 8 positions are batched)
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        sp[k] = -1;
                        if (needs[k] && wq[k] < a) { const int sv = __ldg(&t.sib[base - 8 * k - gl]); if (((unsigned)sv >> 26) == 1u) sp[k] = sv & QMASK; }
                    }
                    int cs[8], ce[8], cv[8];
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (sp[k] >= 0) { const int4 c = __ldg(&t.SR0[sp[k]]); cs[k] = c.x; ce[k] = c.y; cv[k] = stop_reached(ld_relaxed(&stopS[sp[k]])); }
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if (sp[k] >= 0) {
#pragma unroll
                            for (int fa = 0; fa < 4; fa++)
                                if (fa < La && sp[k] >= sAchr[grp][fa].x && sp[k] < sAchr[grp][fa].y && sA0[grp][fa].y <= ce[k] &&
                                    sA0[grp][fa].z >= cs[k] && cv[k] <= sA1[grp][fa].x) needs[k] = false;
                        }
                    }
                }
#pragma unroll
                for (int k = 7; k >= 0; k--) {
                    const unsigned nm = (__ballot_sync(gmask, needs[k]) >> gsh) & 0xffu;
                    if (nm) adv = 8 * k + __ffs(nm) - 1;
                }
                base -= adv;
                if (adv < 64) wide = false;
                if (gl == 0) { d_steps++; st_relaxed(&stop[offa + fi], -2 - (base + 1)); st_relaxed(&stopS[posf], -2 - (base + 1)); }
                d_fsteps++;
            }
            else {
                const int p = base - gl;
                int fl = 0;
                bool cheap = true;                                                 // nothing in this step needed an evaluation
                if (p >= lo) {
                    const int4 c0 = __ldg(&t.SR0[p]);
                    b = c0.w & QMASK;
                    if (b != a && c0.y >= f.y                                      // closed overlap (start_p <= end_f by p <= ub)
                        && !(b < a && stop_reached(ld_relaxed(&stopS[p])) <= posf)     // b's scan of this interval passed a: seen
                        && !(b < a && t.sib && La <= 4 && seen_via_sibling(t, stopS, p, sA0[grp], sA1[grp], sAchr[grp], La))) {
                    cheap = false;
                    const int4 c1 = __ldg(&t.SR1[p]);
                    if (difflen_ok(ria.x, ria.y, ria.z, c1.x, c1.y, c1.z)) {
                        const int offb = (int)((unsigned)c1.w >> 6), Lb = (c1.w & 63) + 1;
                        if (La <= 4 && Lb <= 4) {
                            // ---- lists in registers, everything unrolled; all loads of this candidate are issued together
                            const int bP = (b < a) ? __ldg(&isP[b]) : 1;           // b < a and never breaking: it saw the pair
                            int4 bg[4]; int2 bq[4]; int sb[4];
#pragma unroll
                            for (int g = 0; g < 4; g++) {
                                bg[g] = g < Lb ? rm0(t, offb + g) : make_int4(-2, 0, 0, 0x7fffffff);
                                bq[g] = g < Lb ? rm1(t, offb + g) : make_int2(-1, -1);
                                sb[g] = (b < a && g < Lb) ? ld_relaxed(&stop[offb + g]) : 0x7fffffff;
                            }
                            if (bP) {
                                bool met = false;                                  // pair already seen earlier in this very query?
                                unsigned m[4];
#pragma unroll
                                for (int fa = 0; fa < 4; fa++) {
                                    const int4 af = fa < La ? sA0[grp][fa] : make_int4(-1, 0, 0, 0x7fffffff);
                                    const int aub = sA1[grp][fa].y, ast = sStop[grp][fa];
                                    unsigned r = 0;
#pragma unroll
                                    for (int g = 0; g < 4; g++) {
                                        r |= (matchT<ALLMATCH>(af, bg[g]) ? 1u : 0u) << g;
                                        met |= (fa < fi) && af.x == bg[g].x && ast <= bq[g].x && bq[g].x <= aub && bg[g].z >= af.y;
                                    }
                                    m[fa] = r;
                                }
#pragma unroll
                                for (int g = 0; g < 4; g++) met |= bg[g].x == f.x && bq[g].x > p && bq[g].x <= top && bg[g].z >= f.y;
                                if (!met) {
                                    unsigned used = 0; int n = 0;
#pragma unroll
                                    for (int fa = 0; fa < 4; fa++) { const unsigned av = m[fa] & ~used; if (av) { used |= av & (0u - av); n++; } }
                                    fl = RF_TESTED;
                                    if (n > 0) {
                                        bool vis = false, unres = false;
                                        if (b < a) {                               // b queried first: did its scans get here?
#pragma unroll
                                            for (int g = 0; g < 4; g++) {
#pragma unroll
                                                for (int fa = 0; fa < 4; fa++) {
                                                    const int4 af = fa < La ? sA0[grp][fa] : make_int4(-1, 0, 0, 0);
                                                    const int pa = sA1[grp][fa].x;
                                                    if (af.x == bg[g].x && pa <= bq[g].y && af.z >= bg[g].y) {
                                                        if (stop_reached(sb[g]) <= pa) vis = true; else if (sb[g] < 0) unres = true;
                                                    }
                                                }
                                            }
                                        }
                                        if (vis) { }
                                        else if (unres) fl |= RF_UNRES;
                                        else fl |= RF_REACH | ((La + Lb - n) <= s_umax[n] ? RF_EDGE : 0);
                                    }
                                }
                            }
                        } else if (b > a || __ldg(&isP[b])) {
                            fl = replay_eval_general(t, s_umax, stop, sStop[grp], a, offa, La, fi, f, top, p, b, offb, Lb);
                        }
                    }
                    }
                }
                const unsigned U = (__ballot_sync(gmask, fl & RF_UNRES) >> gsh) & 0xffu;
                const unsigned M = (__ballot_sync(gmask, fl & RF_REACH) >> gsh) & 0xffu;
                const unsigned E = (__ballot_sync(gmask, fl & RF_EDGE) >> gsh) & 0xffu;
                const unsigned Tm = (__ballot_sync(gmask, fl & RF_TESTED) >> gsh) & 0xffu;
                wide = __all_sync(gmask, cheap);
                const int nres = U ? __ffs(U) - 1 : 8;                             // candidates before the first undecided one
                const unsigned rmask = (1u << nres) - 1u;
                int brk = -1;
                for (unsigned mm = M & rmask; mm; mm &= mm - 1) {                  // cluster.py:219-224 in scan order
                    const int l = __ffs(mm) - 1;
                    if (edges + __popc(E & ((2u << l) - 1u)) >= t.Tedge) { brk = l; break; }
                }
                const unsigned cmask = brk >= 0 ? ((2u << brk) - 1u) : rmask;
                Ecommit = E & cmask;
                if (gl == 0) tests += __popc(Tm & cmask);
                edges += __popc(Ecommit);
                if (brk >= 0) stopf = base - brk;
                else {
                    base -= nres; stalled = nres == 0;
                    if (gl == 0 && nres) { st_relaxed(&stop[offa + fi], -2 - (base + 1)); st_relaxed(&stopS[posf], -2 - (base + 1)); }
                }
                if (gl == 0) { d_steps++; d_stall += stalled; }
                d_fsteps++; d_fstall += stalled;
                if (dbg && stalled && d_fstall == 5000 && (U & 1u) && gl == 0) {   // lane 0 is the undecided candidate
                    if (atomicAdd(dbg + 12, 1ull) == 0) {
                        const int wb2 = __ldg(&t.SR1[base]).w; const int ob = (int)((unsigned)wb2 >> 6);
                        dbg[13] = a; dbg[14] = b; dbg[15] = (unsigned)ld_relaxed(&stop[ob]); dbg[16] = (unsigned)ld_relaxed(&stop[ob + 1]);
                        dbg[17] = base; dbg[18] = top; dbg[19] = posf; dbg[20] = rm1(t, ob).x; dbg[21] = rm1(t, ob + 1).x; dbg[22] = sA1[grp][1].x; dbg[23] = fi;
                    }
                }
            }
            if (Ecommit) {
                const int ne = __popc(Ecommit);
                if (chunk_used + ne > RP_CHUNK) {                                  // reserve a fresh chunk, pad the old one
                    for (int k = chunk_used + gl; k < RP_CHUNK; k += 8) pedges[chunk_base + k] = make_int2(-1, -1);
                    if (gl == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)RP_CHUNK);
                    chunk_base = __shfl_sync(gmask, chunk_base, gsh);
                    chunk_used = 0;
                    if (chunk_base + RP_CHUNK > cap_pedges) { if (gl == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                }
                if ((Ecommit >> gl) & 1u) pedges[chunk_base + chunk_used + __popc(Ecommit & ((1u << gl) - 1u))] = make_int2(a, b);
                chunk_used += ne;
            }
            if (stopf >= 0) {                                                      // publish the stop; next filling / read
                if (gl == 0) {
                    sStop[grp][fi] = stopf; st_relaxed(&stop[offa + fi], stopf); st_relaxed(&stopS[posf], stopf);
                    if (La <= 4) ((int *)&sRecStop[grp][a & 31])[fi] = stopf;
                }
                if (dbg && gl == 0 && d_fsteps > 2000) {
                    if (atomicMax(dbg + 4, (unsigned long long)d_fsteps) < (unsigned long long)d_fsteps) {
                        dbg[5] = a; dbg[6] = fi; dbg[7] = top - lo; dbg[8] = d_fstall; dbg[9] = top - stopf; dbg[10] = edges; dbg[11] = La;
                    }
                }
                d_fsteps = 0; d_fstall = 0;
                fi++;
                if (fi == La) { tk++; phase = 0; } else phase = 1;
            }
        }
        }
        if (__all_sync(FULL, stalled || phase == 3)) { __nanosleep(100); d_sleep += lane == 0; }
    }
    if (chunk_used < RP_CHUNK)
        for (int k = chunk_used + gl; k < RP_CHUNK; k += 8) pedges[chunk_base + k] = make_int2(-1, -1);
    for (int o = 16; o; o >>= 1) tests += __shfl_down_sync(FULL, tests, o);
    if (lane == 0 && tests) atomicAdd(n_tests, tests);
    for (int o = 16; o; o >>= 1) { d_iter += __shfl_down_sync(FULL, d_iter, o); d_steps += __shfl_down_sync(FULL, d_steps, o);
                                   d_stall += __shfl_down_sync(FULL, d_stall, o); d_sleep += __shfl_down_sync(FULL, d_sleep, o); }
    if (lane == 0 && dbg) { atomicAdd(dbg, d_iter); atomicAdd(dbg + 1, d_steps); atomicAdd(dbg + 2, d_stall); atomicAdd(dbg + 3, d_sleep); }
}

// ---------------------------------------------------------------- stage 9: union-find (root = smallest query rank)
__device__ __forceinline__ int uf_find(int *parent, int x) {
    for (;;) {
        int p = *(volatile int *)&parent[x];
        if (p == x) return x;
        int gp = *(volatile int *)&parent[p];
        if (gp != p) atomicMin(&parent[x], gp);                                      // path halving, keeps parent <= index
        x = p;
    }
}
__device__ __forceinline__ void uf_union(int *parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }                                  // hook the larger root under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}
// entries (a, b) recorded by k_pair: a not saturating; b > a -> a tested it (edge); b < a -> edge only if b is
// saturating and its scan stopped before reaching a (then a's query tested the pair, direction a -> b)
__global__ void k_union_entries(unsigned long long n, const int2 *__restrict__ entries, const int *__restrict__ isP, Tab t,
                                const int *stop, int *parent, int *ing, unsigned long long *n_edges) {
    unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    bool e = false;
    if (k < n) {
        const int2 ab = entries[k];
        const int a = ab.x, b = ab.y;
        if (a >= 0 && !isP[a]) {
            if (b > a) e = true;
            else if (isP[b]) {
                const int wa = t.RI[a].w, wb = t.RI[b].w;
                const int offa = (int)((unsigned)wa >> 6), La = (wa & 63) + 1, offb = (int)((unsigned)wb >> 6), Lb = (wb & 63) + 1;
                e = true;
                for (int f = 0; f < Lb && e; f++) {
                    const int4 bf = rm0(t, offb + f);
                    const int ubf = rm1(t, offb + f).y, sf = stop[offb + f];
                    for (int g = 0; g < La; g++) {
                        const int4 ag = rm0(t, offa + g);
                        const int pg = rm1(t, offa + g).x;
                        if (ag.x == bf.x && sf <= pg && pg <= ubf && ag.z >= bf.y) { e = false; break; }
                    }
                }
            }
        }
        if (e) { ing[a] = 1; ing[b] = 1; uf_union(parent, a, b); }
    }
    const int cnt = __syncthreads_count(e);                                        // one counter update per block
    if (threadIdx.x == 0 && cnt) atomicAdd(n_edges, (unsigned long long)cnt);
}
__global__ void k_union_edges(unsigned long long n, const int2 *__restrict__ edges, int *parent, int *ing, unsigned long long *n_edges) {
    unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    bool e = false;
    if (k < n) {
        const int2 ab = edges[k];
        if (ab.x >= 0) {                                                            // skip chunk padding
            e = true;
            ing[ab.x] = 1; ing[ab.y] = 1;
            uf_union(parent, ab.x, ab.y);
        }
    }
    const int cnt = __syncthreads_count(e);
    if (threadIdx.x == 0 && cnt && n_edges) atomicAdd(n_edges, (unsigned long long)cnt);
}
__global__ void k_flatten(int Q, int *parent, const int *__restrict__ ing, int *isroot, int *csize) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int r = uf_find(parent, q);
    parent[q] = r;                                                                  // safe: r is a root and stays one
    isroot[q] = (ing[q] && r == q);
    if (ing[q]) atomicAdd(&csize[r], 1);
}
// spanning forest of the local components (multi-GPU exchange)
__global__ void k_forest(int Q, const int *__restrict__ parent, const int *__restrict__ ing, int2 *forest, unsigned long long *n) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool e = q < Q && ing[q] && parent[q] != q;
    const unsigned m = __ballot_sync(0xffffffffu, e);
    unsigned long long at = 0;
    if ((threadIdx.x & 31) == 0 && m) at = atomicAdd(n, (unsigned long long)__popc(m));
    at = __shfl_sync(0xffffffffu, at, 0);
    if (e) forest[at + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = make_int2(q, parent[q]);
}

// ---------------------------------------------------------------- stage 10: cluster / n_reads (main.py:251-257,334-342)
__global__ void k_single_flags(int R, const int *__restrict__ q_of_rid, const int *__restrict__ ing, int *flag) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int q = q_of_rid[r];
    flag[r] = !(q >= 0 && ing[q]);
}
__global__ void k_number(int R, const int *__restrict__ q_of_rid, const int *__restrict__ ing, const int *__restrict__ root,
                         const int *__restrict__ cidx, const int *__restrict__ csize, const int *__restrict__ spos,
                         const int64_t *__restrict__ ncl, int *out_cluster, int *out_n) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int q = q_of_rid[r];
    if (q >= 0 && ing[q]) { int rt = root[q]; out_cluster[r] = cidx[rt]; out_n[r] = csize[rt]; }
    else { out_cluster[r] = (int)(*ncl) + spos[r]; out_n[r] = 1; }                   // singletons after the clusters, bed order
}

// ---------------------------------------------------------------- choose_alignment (cluster.py:237-254, main.py:351-352)
// per read: sum and count of alignment_score over its rows, first row; per cluster: the read with the highest mean
// (IEEE double division, as pandas' groupby.mean of an integer column), first row in table order on ties (idxmax)
__device__ __forceinline__ unsigned long long order_f64(double x) {   // monotone map double -> uint64
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__global__ void k_ca_rows(int A, int R, const int *__restrict__ rid, const int *__restrict__ score, long long *sum, int *cnt, int *first, int *err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    const int r = rid[i];
    if ((unsigned)r >= (unsigned)R) { atomicOr(err, EF_RANGE); return; }
    atomicAdd((unsigned long long *)&sum[r], (unsigned long long)(long long)score[i]);
    atomicAdd(&cnt[r], 1);
    atomicMin(&first[r], i);
}
__global__ void k_ca_best(int R, int C, const long long *__restrict__ sum, const int *__restrict__ cnt, const int *__restrict__ cluster,
                          unsigned long long *best, int *err) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R || cnt[r] == 0) return;
    const int c = cluster[r];
    if ((unsigned)c >= (unsigned)C) { atomicOr(err, EF_RANGE); return; }
    atomicMax(&best[c], order_f64(__ddiv_rn((double)sum[r], (double)cnt[r])));
}
__global__ void k_ca_first(int R, int C, const long long *__restrict__ sum, const int *__restrict__ cnt, const int *__restrict__ first,
                           const int *__restrict__ cluster, const unsigned long long *__restrict__ best, int *minrow) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R || cnt[r] == 0) return;
    const int c = cluster[r];
    if ((unsigned)c >= (unsigned)C) return;
    if (order_f64(__ddiv_rn((double)sum[r], (double)cnt[r])) == best[c]) atomicMin(&minrow[c], first[r]);
}
__global__ void k_ca_flag(int R, int C, const int *__restrict__ cnt, const int *__restrict__ first, const int *__restrict__ cluster,
                          const int *__restrict__ minrow, unsigned char *is_rep, int *rep_read) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    unsigned char f = 0;
    if (cnt[r] > 0) {
        const int c = cluster[r];
        if ((unsigned)c < (unsigned)C && minrow[c] == first[r]) { f = 1; if (rep_read) rep_read[c] = r; }
    }
    is_rep[r] = f;
}

// ---------------------------------------------------------------- integer-issue microbenchmark (roofline denominator)
__global__ void k_int_peak(int iters, int *out) {
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 5, a5 = a0 + 7, a6 = a0 + 11, a7 = a0 + 13;
    const int k = blockIdx.x | 1;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {                                               // 8 independent chains x 2 ops: min/max + add/xor
            a0 = max(a0 + k, a1) ^ j; a1 = min(a1 - k, a2) ^ j; a2 = max(a2 + k, a3) ^ j; a3 = min(a3 - k, a4) ^ j;
            a4 = max(a4 + k, a5) ^ j; a5 = min(a5 - k, a6) ^ j; a6 = max(a6 + k, a7) ^ j; a7 = min(a7 - k, a0) ^ j;
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678) out[0] = a0;
}

// ================================================================ host pipeline
struct Pipe {
    // inputs (device)
    fslrc_table tb;
    fslrc_params pr;
    int A, R, F, D, Q, nP, Tedge, pair_blocks;
    int *err;
    int64_t *cnt;            // device counters: 0 F,1 D,2 Q,3 band,4 tests,5 entry slots,6 nP,7 pedge slots,8 edges,9 ncl,10 forest,
                             //                  11 singletons, 12 runs, 13 entries, 14 tight band
    int *q_of_rid, *rid_of_q;
    int4 *SR0, *SR1, *RM, *RI;
    int *pmaxS, *s_chrom, *chrom_lo, *chrom_hi;
    int *isP, *plist, *stop, *stopS;
    int2 *entries, *pedges;
    int4 *PL; PLInfo *plinfo;
    unsigned long long cap_entries, cap_pedges, cap_pl;
    int *parent, *ing;
    unsigned *ticket;
    unsigned char *prim; size_t prim_bytes;   // scratch of the device-wide primitives
    int stage;               // next event index
    Tab tab;
    UmaxTab um;
};

static int mark(fslrc_ctx *ctx, int stage_end) {   // record the event closing `stage_end`
    CK(cudaEventRecord(ctx->ev[stage_end + 1], ctx->stream));
    return 0;
}
static int n_sms(fslrc_ctx *ctx) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, ctx->device);
    return n;
}
// ---- device-wide primitives (prims.cuh): look-back status words + ticket live in one reusable scratch region
static int prim_scratch(fslrc_ctx *ctx, Pipe *P, size_t status_bytes) {
    const size_t need = status_bytes + 256;
    if (need > P->prim_bytes) {
        void *q; const size_t nb = need + (need >> 1);
        CK(cudaMallocAsync(&q, nb, ctx->stream));
        ctx->allocs.push_back(q);
        P->prim = (unsigned char *)q; P->prim_bytes = nb;
    }
    CK(cudaMemsetAsync(P->prim, 0, need, ctx->stream));          // ticket (first 256 bytes) + status words
    return 0;
}
// exclusive sum of int32; *total (device, optional) receives the grand total
static int xscan(fslrc_ctx *ctx, Pipe *P, const int *in, int *out, int n, int64_t *total = nullptr) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) { if (total) CK(cudaMemsetAsync(total, 0, sizeof(int64_t), st)); return 0; }
    const int tiles = nblk(n, prims::SC_TILE);
    int r = prim_scratch(ctx, P, sizeof(unsigned long long) * tiles); if (r) return r;
    KL(prims::k_scan_excl<int>, tiles, prims::SC_THREADS, in, out, n, (unsigned long long *)(P->prim + 256), (unsigned *)P->prim, (long long *)total);
    return 0;
}
// per-segment inclusive prefix max (seg non-decreasing)
static int segmax_scan(fslrc_ctx *ctx, Pipe *P, const int *seg, const int *val, int *out, int n) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) return 0;
    const int tiles = nblk(n, prims::SC_TILE);
    int r = prim_scratch(ctx, P, sizeof(unsigned long long) * tiles); if (r) return r;
    KL(prims::k_scan_segmax, tiles, prims::SC_THREADS, seg, val, out, n, (unsigned long long *)(P->prim + 256), (unsigned *)P->prim);
    return 0;
}
// stable LSD radix sort of (key, value) pairs on key bits [b0, b1); the inputs are left untouched
static int sort_pairs(fslrc_ctx *ctx, Pipe *P, const unsigned *kin, unsigned *kout, const int *vin, int *vout, int n, int b0, int b1) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) return 0;
    const int npass = std::max(1, (b1 - b0 + 7) / 8);
    if (npass > 4) return fail(ctx, FSLRC_ERR_ARG, "radix sort: more than 32 key bits");
    unsigned *hist; DA(hist, 4 * 256);
    CK(cudaMemsetAsync(hist, 0, sizeof(unsigned) * 4 * 256, st));
    unsigned *tk = nullptr; int *tv = nullptr;
    if (npass > 1) { DA(tk, n); DA(tv, n); }
    KL(prims::k_rs_hist, std::min(nblk(n, 256 * 16), n_sms(ctx) * 8), 256, kin, n, b0, b1, npass, hist);
    KL(prims::k_rs_scan, npass, 256, hist);
    const int tiles = nblk(n, prims::RS_TILE);
    const unsigned *ks = kin; const int *vs = vin;
    for (int p = 0; p < npass; p++) {
        unsigned *kd = ((npass - 1 - p) % 2 == 0) ? kout : tk;       // the last pass lands in (kout, vout)
        int *vd = ((npass - 1 - p) % 2 == 0) ? vout : tv;
        int r = prim_scratch(ctx, P, sizeof(unsigned) * 256 * (size_t)tiles); if (r) return r;
        const int nbits = std::min(8, b1 - b0 - 8 * p);
        ctx->launches++;
        prims::k_rs_onesweep<<<tiles, prims::RS_THREADS, sizeof(prims::RsSmem), st>>>(ks, kd, vs, vd, n, b0 + 8 * p, (1 << nbits) - 1, hist + 256 * p,
                                                                                      (unsigned *)(P->prim + 256), (unsigned *)P->prim);
        ks = kd; vs = vd;
    }
    return 0;
}
static int read_counts(fslrc_ctx *ctx, Pipe *P) {   // device counters + error word -> pinned host
    CK(cudaMemcpyAsync(ctx->h_pin, P->cnt, 48 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_pin + 48, P->err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
static int err_code(fslrc_ctx *ctx) {
    int e = (int)(ctx->h_pin[48] & 0xffffffff);
    if (!e) return 0;
    if (e & EF_RANGE) return fail(ctx, FSLRC_ERR_RANGE, "a table value is out of range (read_id/chrom id, negative coordinate, n_alignments >= 65535 or a bad `order`)");
    if (e & EF_ZERO) return fail(ctx, FSLRC_ERR_ZERO_DIVISOR, "aln_size, qlen2 or n_alignments <= 0 on a filling (the reference raises ZeroDivisionError)");
    if (e & EF_TOOMANY) return fail(ctx, FSLRC_ERR_TOO_MANY_FILLINGS, "a read has more than 64 fillings");
    if (e & EF_NALN) return fail(ctx, FSLRC_ERR_NALN_NOT_CONSTANT, "n_alignments is not constant over the rows of a read");
    return fail(ctx, FSLRC_ERR_OVERFLOW, "internal edge buffer overflow");
}
static int bits_for(int64_t n) { int b = 1; while ((1ll << b) < n && b < 32) b++; return b; }

// ---- stages 1-5: ingestion, orders, records (replicated on every rank)
static int pipe_prepare(fslrc_ctx *ctx, Pipe *P) {
    const fslrc_table &tb = P->tb; const fslrc_params &pr = P->pr;
    cudaStream_t st = ctx->stream;
    const int A = P->A, R = P->R, TB = 256;
    DA(P->err, 1); DA(P->cnt, 48);
    CK(cudaMemsetAsync(P->err, 0, sizeof(int), st));
    CK(cudaMemsetAsync(P->cnt, 0, 48 * sizeof(int64_t), st));
    long long *d_clen; unsigned char *d_cmask;
    DA(d_clen, pr.n_chrom); DA(d_cmask, pr.n_chrom);
    if (pr.n_chrom > 0) {
        CK(cudaMemcpyAsync(d_clen, pr.chrom_len, sizeof(int64_t) * pr.n_chrom, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_cmask, pr.chrom_masked, pr.n_chrom, cudaMemcpyHostToDevice, st));
    }
    // ---- stage 1: keep_fillings
    int *first, *last, *qmin, *qmax, *flagA, *posA;
    DA(first, R); DA(last, R); DA(qmin, R); DA(qmax, R); DA(flagA, A); DA(posA, A);
    DA(P->q_of_rid, R);
    if (R > 0) {
        KL(k_fill<int>, nblk(R, TB), TB, first, R, 0x7fffffff);
        KL(k_fill<int>, nblk(R, TB), TB, last, R, -1);
        KL(k_fill<int>, nblk(R, TB), TB, qmin, R, 0x7fffffff);
        KL(k_fill<int>, nblk(R, TB), TB, qmax, R, (int)0x80000000);
        KL(k_fill<int>, nblk(R, TB), TB, P->q_of_rid, R, -1);
    }
    if (A > 0) {
        KL(k_first_last, nblk(A, TB), TB, A, R, tb.read_id, first, last, P->err);
        KL(k_keep, nblk(A, TB), TB, A, R, tb.read_id, first, last, tb.qstart, tb.qend, flagA, qmin, qmax);
        int r = xscan(ctx, P, flagA, posA, A, P->cnt + 0); if (r) return r;
    }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    const int F = P->F = (int)ctx->h_pin[0];
    if (tb.order && tb.n_order != F) return fail(ctx, FSLRC_ERR_ARG, "order has the wrong length (must equal the number of fillings)");
    int4 *FR0; int2 *FR1;
    DA(FR0, F); DA(FR1, F);
    if (A > 0) KL(k_fill_records, nblk(A, TB), TB, A, flagA, posA, tb.read_id, tb.chrom, tb.rstart, tb.rend, tb.aln_size, tb.n_alignments,
                  pr.n_chrom, FR0, FR1, P->err);
    { int r = mark(ctx, 1); if (r) return r; }
    // ---- stage 2: mask (cluster.py:89-106; dropping masked fillings before or after the sort is the same list) + data order
    // (cluster.py:114): the caller's permutation, or a stable radix sort by start (ties keep bed order)
    int *flagF, *posF;
    DA(flagF, F); DA(posF, F);
    if (F > 0) {
        KL(k_mask_flags, nblk(F, TB), TB, F, tb.order, FR0, pr.n_chrom, d_clen, d_cmask, pr.mask_subtelomere, (long long)pr.subtel, flagF, P->err);
        int r = xscan(ctx, P, flagF, posF, F, P->cnt + 1); if (r) return r;
    }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    const int D = P->D = (int)ctx->h_pin[1];
    int *dfill; DA(dfill, D);
    if (tb.order) {
        if (F > 0) KL(k_compact_fillings, nblk(F, TB), TB, F, tb.order, flagF, posF, FR0, (int *)nullptr, dfill);
    } else {
        int *key, *key2, *v; DA(key, D); DA(key2, D); DA(v, D);
        if (F > 0) KL(k_compact_fillings, nblk(F, TB), TB, F, (const int *)nullptr, flagF, posF, FR0, key, v);
        int r = sort_pairs(ctx, P, (const unsigned *)key, (unsigned *)key2, v, dfill, D, 0, 32); if (r) return r;
    }
    int4 *IT0; int2 *IT1; int *firstdp;
    DA(IT0, D); DA(IT1, D); DA(firstdp, R);
    if (R > 0) KL(k_fill<int>, nblk(R, TB), TB, firstdp, R, 0x7fffffff);
    unsigned *tkey = nullptr, *tkey2 = nullptr, *tval = nullptr, *tval2 = nullptr;
    if (D > 0) KL(k_build_items, nblk(D, TB), TB, D, dfill, FR0, FR1, IT0, IT1, firstdp, P->err);
    if (D > 0 && D < (1 << 26)) {
        DA(tkey, D); DA(tkey2, D); DA(tval, D); DA(tval2, D);
        KL(k_tie_delta, nblk(D, TB), TB, D, IT0, tkey, tval, (unsigned long long *)(P->cnt + 41), P->err);
    }
    { int r = mark(ctx, 2); if (r) return r; }
    // ---- stage 3: query rank + per-read lists
    int *flagD, *posD, *it_q;
    DA(flagD, D); DA(posD, D); DA(it_q, D);
    if (D > 0) {
        KL(k_is_first, nblk(D, TB), TB, D, IT0, firstdp, flagD);
        int r = xscan(ctx, P, flagD, posD, D, P->cnt + 2); if (r) return r;
    }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    const int Q = P->Q = (int)ctx->h_pin[2];
    DA(P->rid_of_q, Q);
    int *qs, *rm_dp, *iotaD, *rmidx, *off, *len_end;
    DA(qs, D); DA(rm_dp, D); DA(iotaD, D); DA(rmidx, D); DA(off, Q); DA(len_end, Q);
    DA(P->RI, Q);
    if (R > 0) KL(k_rank_reads, nblk(R, TB), TB, R, firstdp, posD, P->q_of_rid, P->rid_of_q);
    if (D > 0) {
        KL(k_item_q, nblk(D, TB), TB, D, IT0, P->q_of_rid, it_q);
        KL(k_iota, nblk(D, TB), TB, iotaD, D);
        int r = sort_pairs(ctx, P, (const unsigned *)it_q, (unsigned *)qs, iotaD, rm_dp, D, 0, bits_for(Q)); if (r) return r;
        KL(k_read_bounds, nblk(D, TB), TB, D, qs, rm_dp, rmidx, off, len_end);
        KL(k_read_info, nblk(Q, TB), TB, Q, P->rid_of_q, off, len_end, rm_dp, IT1, qmin, qmax, pr.qlen_c, pr.naln_c, P->RI, P->err);
    }
    { int r = mark(ctx, 3); if (r) return r; }
    // ---- stage 4: IntervalMap order: (chrom, start asc, end desc, data order)
    int *s_dp; DA(s_dp, D);
    const bool tie_ok = tkey && ctx->h_pin[41] == 0;                 // (read back with the stage-3 counters)
    if (D > 0 && tie_ok) {                                           // one stable partition by chromosome + local tie fix
        int r = sort_pairs(ctx, P, tkey, tkey2, (const int *)tval, (int *)tval2, D, 0, bits_for(pr.n_chrom)); if (r) return r;
        KL(k_apply_delta, nblk(D, TB), TB, D, tval2, s_dp);
    } else if (D > 0) {                                              // long runs of equal (chrom, start): two full radix sorts
        // (chrom, start, end desc, data order) as three stable sorts from data order: ~end, start, chrom
        unsigned *ek, *ek2; int *v1, *v2;
        DA(ek, D); DA(ek2, D); DA(v1, D); DA(v2, D);
        KL(k_end_keys, nblk(D, TB), TB, D, IT0, ek);
        int r = sort_pairs(ctx, P, ek, ek2, iotaD, v1, D, 0, 32); if (r) return r;
        KL(k_gather_key, nblk(D, TB), TB, D, v1, IT0, 0, ek);
        r = sort_pairs(ctx, P, ek, ek2, v1, v2, D, 0, 32); if (r) return r;
        KL(k_gather_key, nblk(D, TB), TB, D, v2, IT0, 1, ek);
        r = sort_pairs(ctx, P, ek, ek2, v2, s_dp, D, 0, bits_for(pr.n_chrom)); if (r) return r;
    }
    { int r = mark(ctx, 4); if (r) return r; }
    // ---- stage 5: records, thresholds, bands
    int *s_end;
    DA(P->SR0, D); DA(P->SR1, D); DA(P->RM, 2 * (int64_t)D); DA(P->s_chrom, D); DA(s_end, D);
    DA(P->pmaxS, D); DA(P->chrom_lo, pr.n_chrom); DA(P->chrom_hi, pr.n_chrom);
    if (pr.n_chrom > 0) { CK(cudaMemsetAsync(P->chrom_lo, 0, sizeof(int) * pr.n_chrom, st)); CK(cudaMemsetAsync(P->chrom_hi, 0, sizeof(int) * pr.n_chrom, st)); }
    if (D > 0) {
        if (D >= (1 << 26)) return fail(ctx, FSLRC_ERR_RANGE, "more than 2^26 intervals");
        int *s_m; DA(s_m, D);
        KL(k_records, nblk(D, TB), TB, D, s_dp, rmidx, it_q, IT0, IT1, P->RI, pr.overlap, P->SR0, P->SR1,
                                               s_m, P->s_chrom, s_end, P->chrom_lo, P->chrom_hi, P->err);
        int r = segmax_scan(ctx, P, P->s_chrom, s_end, P->pmaxS, D); if (r) return r;
        KL(k_bands, nblk(D, 256), 256, D, P->SR0, s_m, P->s_chrom, P->pmaxS, P->chrom_lo, P->chrom_hi, P->RM,
           (unsigned long long *)(P->cnt + 3), (unsigned long long *)(P->cnt + 14));
    }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    long long T = pr.edge_threshold;
    P->Tedge = T > 0x7fffffffLL ? 0x7fffffff : (T < -0x7fffffffLL ? -0x7fffffff : (int)T);
    Tab &t = P->tab;
    t.SR0 = P->SR0; t.SR1 = P->SR1; t.RM = P->RM; t.RI = P->RI; t.pmaxS = P->pmaxS;
    t.chrom_lo = P->chrom_lo; t.chrom_hi = P->chrom_hi; t.sib = nullptr; t.D = D; t.Q = Q; t.Tedge = P->Tedge;
    memcpy(P->um.v, pr.umax, sizeof(P->um.v));
    // relation entries: a read records at most edge_threshold + 7 passing partners (one step past the threshold), and every
    // entry is a distinct (filling of a, band position) hit
    const unsigned long long tight = (unsigned long long)ctx->h_pin[14];
    unsigned long long capT = P->Tedge > 0 ? (unsigned long long)Q * ((unsigned long long)P->Tedge + 7ull) : 0ull;
    P->pair_blocks = std::max(1, std::min(nblk(Q, PK_GROUPS), n_sms(ctx) * 8));
    P->cap_entries = std::min<unsigned long long>(capT, tight) + (unsigned long long)PK_CHUNK * PK_WARPS * P->pair_blocks + 64;
    DA(P->entries, P->cap_entries);
    // partner records of saturating reads (replay LIST mode): at most RP_K per read and never more than tight-band hits
    P->cap_pl = std::min<unsigned long long>((unsigned long long)Q * RP_K, tight) + (unsigned long long)PL_CHUNK * PK_WARPS * P->pair_blocks * 2 + 64;
    DA(P->PL, 2 * P->cap_pl); DA(P->plinfo, Q);
    DA(P->isP, Q); DA(P->stop, D); DA(P->stopS, D); DA(P->parent, Q); DA(P->ing, Q); DA(P->ticket, 1);
    if (Q > 0) CK(cudaMemsetAsync(P->isP, 0, sizeof(int) * Q, st));
    return mark(ctx, 5);
}


// ---- stage 6: pair kernel on one shard
static int launch_pair(fslrc_ctx *ctx, Pipe *P, int shard, int nshard, int lists_only) {
    cudaStream_t st = ctx->stream;
    if (P->pr.overlap > 0.0)
        KL(k_pair<false>, P->pair_blocks, PK_WARPS * 32, P->tab, P->um, shard, nshard, lists_only, P->isP, P->entries, (unsigned long long *)(P->cnt + 5),
           P->cap_entries, P->PL, P->plinfo, (unsigned long long *)(P->cnt + 40), P->cap_pl,
           (unsigned long long *)(P->cnt + 4), (unsigned long long *)(P->cnt + 13), P->err);
    else
        KL(k_pair<true>, P->pair_blocks, PK_WARPS * 32, P->tab, P->um, shard, nshard, lists_only, P->isP, P->entries, (unsigned long long *)(P->cnt + 5),
           P->cap_entries, P->PL, P->plinfo, (unsigned long long *)(P->cnt + 40), P->cap_pl,
           (unsigned long long *)(P->cnt + 4), (unsigned long long *)(P->cnt + 13), P->err);
    return 0;
}
static int pipe_pair(fslrc_ctx *ctx, Pipe *P, int shard, int nshard) {
    if (P->Q > 0) { int r = launch_pair(ctx, P, shard, nshard, 0); if (r) return r; }
    return mark(ctx, 6);
}

// ---- stages 7-9 (after isP is complete): saturating set, replay, union-find
static int pipe_replay_union(fslrc_ctx *ctx, Pipe *P, int shard, int nshard) {
    cudaStream_t st = ctx->stream;
    const int Q = P->Q, D = P->D, TB = 256;
    int *posQ;
    DA(posQ, Q);
    if (Q > 0 && nshard > 1) { int r = launch_pair(ctx, P, shard, nshard, 1); if (r) return r; }
    if (Q > 0) {
        int r = xscan(ctx, P, P->isP, posQ, Q, P->cnt + 6); if (r) return r;
    }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    const int nP = P->nP = (int)ctx->h_pin[6];
    DA(P->plist, nP);
    if (Q > 0) KL(k_compact_flagged, nblk(Q, TB), TB, Q, P->isP, posQ, P->plist);
    // stops: a read that never breaks walks every filling's scan to the chromosome start (0 <= any position)
    if (D > 0) { CK(cudaMemsetAsync(P->stop, 0, sizeof(int) * D, st)); CK(cudaMemsetAsync(P->stopS, 0, sizeof(int) * D, st)); }
    CK(cudaMemsetAsync(P->ticket, 0, sizeof(unsigned), st));
    { int r = mark(ctx, 7); if (r) return r; }
    // a saturating read adds at most edge_threshold edges in the scan that reaches the threshold and one per later filling
    const int replay_blocks_max = n_sms(ctx) * 16;
    unsigned long long capp = (unsigned long long)nP * ((unsigned long long)std::max(P->Tedge, 0) + LMAX) + 64;
    unsigned long long alt = 2ull * (unsigned long long)ctx->h_pin[3] + 64;
    P->cap_pedges = std::min(capp, alt) + (unsigned long long)RP_CHUNK * RG_GROUPS * (replay_blocks_max + 1);
    DA(P->pedges, P->cap_pedges);
    if (nP > 0) {
        int *sflag, *spos, *sstart, *rflag, *rpos, *rstart;
        DA(sflag, nP); DA(spos, nP); DA(sstart, nP); DA(rflag, nP); DA(rpos, nP); DA(rstart, nP);
        KL(k_run_flags, nblk(nP, TB), TB, nP, P->plist, P->RI, P->RM, sflag, P->stop, P->stopS);
        int r = xscan(ctx, P, sflag, spos, nP, P->cnt + 15); if (r) return r;
        KL(k_compact_flagged, nblk(nP, TB), TB, nP, sflag, spos, sstart);
        KL(k_run_cut, nblk(nP, TB), TB, nP, sflag, spos, sstart, P->cnt + 15, rflag);
        r = xscan(ctx, P, rflag, rpos, nP, P->cnt + 12); if (r) return r;
        KL(k_compact_flagged, nblk(nP, TB), TB, nP, rflag, rpos, rstart);
        r = read_counts(ctx, P); if (r) return r;
        const int nRuns = (int)ctx->h_pin[12];
        int blocks = std::min(nblk(nRuns, RG_GROUPS), replay_blocks_max);
        const bool walk = !(P->pr.overlap > 0.0) || ctx->h_pin[43] != 0;   // some read without partner records?
        if (walk && D > 0) {
            int *sib; DA(sib, D);
            KL(k_sib, nblk(D, TB), TB, D, P->SR0, P->SR1, P->RM, sib);
            P->tab.sib = sib;
        }
#define REPLAY_ARGS P->tab, P->um, nP, P->plist, nRuns, rstart, P->isP, P->PL, P->plinfo, P->stop, P->stopS, P->ticket, P->pedges, \
                    (unsigned long long *)(P->cnt + 7), P->cap_pedges, (unsigned long long *)(P->cnt + 4), P->err, (unsigned long long *)(P->cnt + 16)
        if (!(P->pr.overlap > 0.0)) KL((k_replay<true, true>), blocks, RG_WARPS * 32, REPLAY_ARGS);
        else if (walk) KL((k_replay<false, true>), blocks, RG_WARPS * 32, REPLAY_ARGS);
        else KL((k_replay<false, false>), std::min(nblk(nRuns, RG_GROUPS), n_sms(ctx) * 16), RG_WARPS * 32, REPLAY_ARGS);
#undef REPLAY_ARGS
    }
    { int r = mark(ctx, 8); if (r) return r; }
    { int r = read_counts(ctx, P); if (r) return r; r = err_code(ctx); if (r) return r; }
    const unsigned long long nent = std::min<unsigned long long>((unsigned long long)ctx->h_pin[5], P->cap_entries);
    const unsigned long long nped = std::min<unsigned long long>((unsigned long long)ctx->h_pin[7], P->cap_pedges);
    if (Q > 0) {
        KL(k_iota, nblk(Q, TB), TB, P->parent, Q);
        CK(cudaMemsetAsync(P->ing, 0, sizeof(int) * Q, st));
    }
    if (nent > 0) KL(k_union_entries, nblk((int64_t)nent, TB), TB, nent, P->entries, P->isP, P->tab, P->stop, P->parent, P->ing,
                                                                           (unsigned long long *)(P->cnt + 8));
    // replayed edges are identical on every rank; rank 0 contributes them once
    if (nped > 0 && shard == 0) KL(k_union_edges, nblk((int64_t)nped, TB), TB, nped, P->pedges, P->parent, P->ing, (unsigned long long *)(P->cnt + 8));
    (void)nshard;
    return mark(ctx, 9);
}

// ---- stage 10: numbering
static int pipe_number(fslrc_ctx *ctx, Pipe *P, int *out_cluster, int *out_n) {
    cudaStream_t st = ctx->stream;
    const int Q = P->Q, R = P->R, TB = 256;
    int *isroot, *cidx, *csize, *sflag, *spos;
    DA(isroot, Q); DA(cidx, Q); DA(csize, Q); DA(sflag, R); DA(spos, R);
    if (Q > 0) {
        CK(cudaMemsetAsync(csize, 0, sizeof(int) * Q, st));
        KL(k_flatten, nblk(Q, TB), TB, Q, P->parent, P->ing, isroot, csize);
        int r = xscan(ctx, P, isroot, cidx, Q, P->cnt + 9); if (r) return r;
    }
    if (R > 0) {
        KL(k_single_flags, nblk(R, TB), TB, R, P->q_of_rid, P->ing, sflag);
        int r = xscan(ctx, P, sflag, spos, R, P->cnt + 11); if (r) return r;
        KL(k_number, nblk(R, TB), TB, R, P->q_of_rid, P->ing, P->parent, cidx, csize, spos, P->cnt + 9, out_cluster, out_n);
    }
    return mark(ctx, 10);
}

static void fill_stats(fslrc_ctx *ctx, Pipe *P, fslrc_stats *s) {
    if (!s) return;
    const int64_t *h = ctx->h_pin;
    memset(s, 0, sizeof(*s));
    s->n_fillings = P->F; s->n_intervals = P->D; s->n_query_reads = P->Q;
    s->band_pairs = h[3]; s->pair_tests = h[4]; s->relation_entries = h[13]; s->saturating_reads = P->nP;
    s->edges = h[8]; s->components = h[9]; s->clustered_reads = (int64_t)P->R - h[11];
    s->partner_records = h[42];
    s->no_clusters = h[9] == 0;
    if (getenv("FSLRC_DEBUG")) fprintf(stderr, "[fslrc] replay: warp-iterations %lld, group-steps %lld, stalled %lld, sleeps %lld, runs %lld, walk-mode reads %lld\n",
                                       (long long)h[16], (long long)h[17], (long long)h[18], (long long)h[19], (long long)h[12], (long long)h[43]);
    if (getenv("FSLRC_DEBUG") && h[28]) fprintf(stderr, "[fslrc] long stall: a %lld waits b %lld stops %d %d base %lld top %lld posf %lld bpos %lld %lld apos2 %lld fi %lld (n=%lld)\n",
                                       (long long)h[29], (long long)h[30], (int)h[31], (int)h[32], (long long)h[33], (long long)h[34], (long long)h[35], (long long)h[36], (long long)h[37], (long long)h[38], (long long)h[39], (long long)h[28]);
    if (getenv("FSLRC_DEBUG") && h[20]) fprintf(stderr, "[fslrc] longest walk: %lld steps (stalled %lld) read %lld filling %lld/%lld band %lld walked %lld edges %lld\n",
                                       (long long)h[20], (long long)h[24], (long long)h[21], (long long)h[22], (long long)h[27], (long long)h[23], (long long)h[25], (long long)h[26]);
    for (int i = 0; i < FSLRC_N_STAGES; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) { ms = 0.f; cudaGetLastError(); }
        s->stage_ms[i] = ms;
    }
}

static inline bool oc_host_path_ok(const fslrc_table *tb) { return tb->aln_size_is_qspan != 0; }   // aln_size may be omitted then
static int check_args(fslrc_ctx *ctx, const fslrc_table *tb, const fslrc_params *pr, const void *oc, const void *on) {
    if (!ctx) return FSLRC_ERR_ARG;
    if (!tb || !pr || !oc || !on) return fail(ctx, FSLRC_ERR_ARG, "null argument");
    if (tb->n_rows < 0 || tb->n_reads < 0 || tb->n_rows > 0x7ffffff0LL || tb->n_reads > 0x7ffffff0LL) return fail(ctx, FSLRC_ERR_ARG, "table size out of range");
    if (tb->n_rows > 0 && (!tb->read_id || (!tb->chrom && !tb->chrom_u8) || !tb->rstart || !tb->rend || (!tb->aln_size && !oc_host_path_ok(tb)) || !tb->qstart || !tb->qend ||
                           (!tb->n_alignments && !tb->n_alignments_u16)))
        return fail(ctx, FSLRC_ERR_ARG, "null column");
    if (pr->n_chrom < 0 || pr->n_chrom > (1 << 20) || (pr->n_chrom > 0 && (!pr->chrom_len || !pr->chrom_masked))) return fail(ctx, FSLRC_ERR_ARG, "bad chromosome tables");
    if (pr->overlap != pr->overlap || pr->qlen_c != pr->qlen_c || pr->naln_c != pr->naln_c) return fail(ctx, FSLRC_ERR_ARG, "NaN option");
    return 0;
}

static int run_device(fslrc_ctx *ctx, Pipe *P, int32_t *oc, int32_t *on, fslrc_stats *stats) {
    int r = pipe_prepare(ctx, P); if (r) return r;
    r = pipe_pair(ctx, P, 0, 1); if (r) return r;
    r = pipe_replay_union(ctx, P, 0, 1); if (r) return r;
    r = pipe_number(ctx, P, oc, on); if (r) return r;
    return 0;
}

extern "C" {

int fslrc_version(void) { return FSLRC_VERSION; }
const char *fslrc_stage_name(int s) { return (s >= 0 && s < FSLRC_N_STAGES) ? STAGE_NAMES[s] : ""; }
const char *fslrc_last_error(const fslrc_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int fslrc_create(int device, fslrc_ctx **out) {
    if (!out) return FSLRC_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return FSLRC_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return FSLRC_ERR_CUDA;
    fslrc_ctx *ctx = new fslrc_ctx();
    ctx->device = device; ctx->launches = 0; ctx->err[0] = 0; ctx->stream = nullptr; ctx->pipe = nullptr; ctx->tsv = nullptr; ctx->h_pin = nullptr;
    if (cudaMallocHost((void **)&ctx->h_pin, 64 * sizeof(int64_t)) != cudaSuccess) { delete ctx; return FSLRC_ERR_CUDA; }
    for (int i = 0; i <= FSLRC_N_STAGES; i++) cudaEventCreate(&ctx->ev[i]);
    cudaFuncSetAttribute(prims::k_rs_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(prims::RsSmem));
    cudaMemPool_t pool;                                   // keep freed scratch cached between calls
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return 0;
}

void fslrc_destroy(fslrc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_all(ctx);
    tsv_free(ctx);
    cudaDeviceSynchronize();
    for (int i = 0; i <= FSLRC_N_STAGES; i++) cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    delete ctx->pipe;
    delete ctx;
}

int fslrc_cluster_device(fslrc_ctx *ctx, const fslrc_table *table, const fslrc_params *params, int32_t *out_cluster,
                         int32_t *out_n_reads, fslrc_stats *stats, void *stream) {
    int r = check_args(ctx, table, params, out_cluster, out_n_reads); if (r) return r;
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    Pipe P; memset(&P, 0, sizeof(P));
    P.tb = *table; P.pr = *params; P.A = (int)table->n_rows; P.R = (int)table->n_reads;
    for (int i = 0; i <= FSLRC_N_STAGES; i++) cudaEventRecord(ctx->ev[i], ctx->stream);
    r = run_device(ctx, &P, out_cluster, out_n_reads, stats);
    if (!r) { r = mark(ctx, 11); }
    if (!r) { r = read_counts(ctx, &P); if (!r) r = err_code(ctx); }
    if (!r) fill_stats(ctx, &P, stats);
    free_all(ctx);
    cudaStreamSynchronize(ctx->stream);
    return r;
}

int fslrc_cluster_host(fslrc_ctx *ctx, const fslrc_table *table, const fslrc_params *params, int32_t *out_cluster,
                       int32_t *out_n_reads, fslrc_stats *stats, void *stream) {
    int r = check_args(ctx, table, params, out_cluster, out_n_reads); if (r) return r;
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    const int64_t A = table->n_rows, R = table->n_reads;
    for (int i = 0; i <= FSLRC_N_STAGES; i++) cudaEventRecord(ctx->ev[i], st);
    fslrc_table d = *table;
    int32_t *cols[8]; const int32_t *src[8] = {table->read_id, table->chrom, table->rstart, table->rend, table->aln_size,
                                               table->qstart, table->qend, table->n_alignments};
    unsigned char *d_c8 = nullptr; unsigned short *d_n16 = nullptr;
    for (int c = 0; c < 8; c++) {
        DA(cols[c], A);
        if (c == 1 && table->chrom_u8) {
            DA(d_c8, A);
            if (A > 0) CK(cudaMemcpyAsync(d_c8, table->chrom_u8, (size_t)A, cudaMemcpyHostToDevice, st));
        } else if (c == 7 && table->n_alignments_u16) {
            DA(d_n16, A);
            if (A > 0) CK(cudaMemcpyAsync(d_n16, table->n_alignments_u16, sizeof(unsigned short) * A, cudaMemcpyHostToDevice, st));
        } else if (c == 4 && !table->aln_size) {
            // derived on the device below
        } else if (A > 0) CK(cudaMemcpyAsync(cols[c], src[c], sizeof(int32_t) * A, cudaMemcpyHostToDevice, st));
    }
    if (A > 0 && (d_c8 || d_n16 || !table->aln_size))
        KL(k_widen, nblk(A, 256), 256, A, d_c8, d_n16, cols[1], cols[7], table->aln_size ? (int *)nullptr : cols[4], cols[5], cols[6]);
    d.read_id = cols[0]; d.chrom = cols[1]; d.rstart = cols[2]; d.rend = cols[3]; d.aln_size = cols[4]; d.qstart = cols[5];
    d.qend = cols[6]; d.n_alignments = cols[7];
    int32_t *d_order = nullptr;
    if (table->order) {
        DA(d_order, table->n_order);
        if (table->n_order > 0) CK(cudaMemcpyAsync(d_order, table->order, sizeof(int32_t) * table->n_order, cudaMemcpyHostToDevice, st));
        d.order = d_order;
    }
    int32_t *d_oc, *d_on;
    DA(d_oc, R); DA(d_on, R);
    CK(cudaEventRecord(ctx->ev[1], st));                                   // closes stage 0 (h2d)
    Pipe P; memset(&P, 0, sizeof(P));
    P.tb = d; P.pr = *params; P.A = (int)A; P.R = (int)R;
    r = run_device(ctx, &P, d_oc, d_on, stats);
    if (!r && R > 0) {
        CK(cudaMemcpyAsync(out_cluster, d_oc, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_n_reads, d_on, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, st));
    }
    if (!r) r = mark(ctx, 11);
    if (!r) { r = read_counts(ctx, &P); if (!r) r = err_code(ctx); }
    if (!r) fill_stats(ctx, &P, stats);
    free_all(ctx);
    cudaStreamSynchronize(st);
    return r;
}

// ---------------------------------------------------------------- multi-GPU staging
int fslrc_mg_prepare(fslrc_ctx *ctx, const fslrc_table *table, const fslrc_params *params, void *stream) {
    int dummy = 0;
    int r = check_args(ctx, table, params, &dummy, &dummy); if (r) return r;
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    free_all(ctx);
    delete ctx->pipe;
    Pipe *P = ctx->pipe = new Pipe(); memset(P, 0, sizeof(*P));
    P->tb = *table; P->pr = *params; P->A = (int)table->n_rows; P->R = (int)table->n_reads;
    for (int i = 0; i <= FSLRC_N_STAGES; i++) cudaEventRecord(ctx->ev[i], ctx->stream);
    r = pipe_prepare(ctx, P);
    if (r) { free_all(ctx); cudaStreamSynchronize(ctx->stream); }
    return r;
}
int fslrc_mg_pair(fslrc_ctx *ctx, int rank, int world, int32_t **counts, int64_t *n_counts) {
    if (!ctx || !ctx->pipe || !counts || !n_counts || world < 1 || rank < 0 || rank >= world) return FSLRC_ERR_ARG;
    Pipe *P = ctx->pipe;
    CK(cudaSetDevice(ctx->device));
    int r = pipe_pair(ctx, P, rank, world); if (r) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    *counts = P->isP; *n_counts = P->Q;
    return 0;
}
int fslrc_mg_replay(fslrc_ctx *ctx, int rank, int world, int32_t **forest, int64_t *n_forest_edges) {
    if (!ctx || !ctx->pipe || !forest || !n_forest_edges) return FSLRC_ERR_ARG;
    Pipe *P = ctx->pipe;
    CK(cudaSetDevice(ctx->device));
    int r = pipe_replay_union(ctx, P, rank, world); if (r) return r;
    cudaStream_t st = ctx->stream;
    const int Q = P->Q, TB = 256;
    int *isroot, *csize; int2 *fo;
    DA(isroot, Q); DA(csize, Q); DA(fo, Q);
    if (Q > 0) {
        CK(cudaMemsetAsync(csize, 0, sizeof(int) * Q, st));
        KL(k_flatten, nblk(Q, TB), TB, Q, P->parent, P->ing, isroot, csize);
        KL(k_forest, nblk(Q, TB), TB, Q, P->parent, P->ing, fo, (unsigned long long *)(P->cnt + 10));
    }
    r = read_counts(ctx, P); if (r) return r;
    r = err_code(ctx); if (r) return r;
    *forest = (int32_t *)fo; *n_forest_edges = ctx->h_pin[10];
    return 0;
}
int fslrc_mg_finish(fslrc_ctx *ctx, const int32_t *all_forest, int64_t n_edges, int32_t *out_cluster, int32_t *out_n_reads,
                    fslrc_stats *stats) {
    if (!ctx || !ctx->pipe || !out_cluster || !out_n_reads || n_edges < 0 || (n_edges > 0 && !all_forest)) return FSLRC_ERR_ARG;
    Pipe *P = ctx->pipe;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int Q = P->Q, TB = 256;
    if (Q > 0) {
        KL(k_iota, nblk(Q, TB), TB, P->parent, Q);
        CK(cudaMemsetAsync(P->ing, 0, sizeof(int) * Q, st));
    }
    if (n_edges > 0) KL(k_union_edges, nblk(n_edges, TB), TB, (unsigned long long)n_edges, (const int2 *)all_forest, P->parent, P->ing, (unsigned long long *)nullptr);
    int r = pipe_number(ctx, P, out_cluster, out_n_reads);
    if (!r) r = mark(ctx, 11);
    if (!r) { r = read_counts(ctx, P); if (!r) r = err_code(ctx); }
    if (!r) fill_stats(ctx, P, stats);
    free_all(ctx);
    cudaStreamSynchronize(st);
    delete ctx->pipe; ctx->pipe = nullptr;
    return r;
}

int fslrc_choose_alignment_host(fslrc_ctx *ctx, int64_t n_rows, int64_t n_reads, int64_t n_clusters, const int32_t *read_id,
                                const int32_t *alignment_score, const int32_t *cluster, uint8_t *out_is_rep, int32_t *out_rep_read,
                                void *stream) {
    if (!ctx) return FSLRC_ERR_ARG;
    if (n_rows < 0 || n_reads < 0 || n_clusters < 0 || n_rows > 0x7ffffff0LL || n_reads > 0x7ffffff0LL || n_clusters > 0x7ffffff0LL ||
        (n_rows > 0 && (!read_id || !alignment_score)) || (n_reads > 0 && (!cluster || !out_is_rep)))
        return fail(ctx, FSLRC_ERR_ARG, "choose_alignment: bad argument");
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    const int A = (int)n_rows, R = (int)n_reads, C = (int)n_clusters, TB = 256;
    int *d_rid, *d_sc, *d_cl, *cnt, *first, *minrow, *rep, *err; long long *sum; unsigned long long *best; unsigned char *flag;
    DA(d_rid, A); DA(d_sc, A); DA(d_cl, R); DA(cnt, R); DA(first, R); DA(minrow, C); DA(rep, C); DA(err, 1); DA(sum, R); DA(best, C); DA(flag, R);
    if (A > 0) {
        CK(cudaMemcpyAsync(d_rid, read_id, sizeof(int) * A, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_sc, alignment_score, sizeof(int) * A, cudaMemcpyHostToDevice, st));
    }
    if (R > 0) {
        CK(cudaMemcpyAsync(d_cl, cluster, sizeof(int) * R, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(cnt, 0, sizeof(int) * R, st));
        CK(cudaMemsetAsync(sum, 0, sizeof(long long) * R, st));
        KL(k_fill<int>, nblk(R, TB), TB, first, R, 0x7fffffff);
    }
    if (C > 0) {
        CK(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * C, st));
        KL(k_fill<int>, nblk(C, TB), TB, minrow, C, 0x7fffffff);
        KL(k_fill<int>, nblk(C, TB), TB, rep, C, -1);
    }
    CK(cudaMemsetAsync(err, 0, sizeof(int), st));
    if (A > 0) KL(k_ca_rows, nblk(A, TB), TB, A, R, d_rid, d_sc, sum, cnt, first, err);
    if (R > 0) {
        KL(k_ca_best, nblk(R, TB), TB, R, C, sum, cnt, d_cl, best, err);
        KL(k_ca_first, nblk(R, TB), TB, R, C, sum, cnt, first, d_cl, best, minrow);
        KL(k_ca_flag, nblk(R, TB), TB, R, C, cnt, first, d_cl, minrow, flag, rep);
        CK(cudaMemcpyAsync(out_is_rep, flag, R, cudaMemcpyDeviceToHost, st));
    }
    if (C > 0 && out_rep_read) CK(cudaMemcpyAsync(out_rep_read, rep, sizeof(int) * C, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 60, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    free_all(ctx);
    cudaStreamSynchronize(st);
    if ((int)(ctx->h_pin[60] & 0xffffffff)) return fail(ctx, FSLRC_ERR_RANGE, "choose_alignment: read id or cluster id out of range");
    return 0;
}

// ---------------------------------------------------------------- mappings.bed ingest / egress on the GPU (tsv.cuh)
}  // extern "C"
struct TsvState {
    std::vector<void *> allocs;
    unsigned char *text; long long n; long long *line_start;
    int n_lines, n_rows, n_reads, n_chrom;
    int *read_id, *chrom, *rstart, *rend, *aln, *qstart, *qend, *naln, *score, *first_row;
    long long *q_off; int *q_len;
    std::vector<std::string> chrom_names;
};
template <typename T>
static int palloc(fslrc_ctx *ctx, T **p, int64_t n) {                  // persistent (until fslrc_tsv_close)
    void *q = nullptr;
    CK(cudaMallocAsync(&q, (size_t)(n > 0 ? n : 1) * sizeof(T), ctx->stream));
    ctx->tsv->allocs.push_back(q);
    *p = (T *)q;
    return 0;
}
#define PA(ptr, n) do { int r__ = palloc(ctx, &(ptr), (int64_t)(n)); if (r__) return r__; } while (0)
static void tsv_free(fslrc_ctx *ctx) {
    if (!ctx->tsv) return;
    for (void *p : ctx->tsv->allocs) cudaFreeAsync(p, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->tsv; ctx->tsv = nullptr;
}
// strings -> dense ids in order of first appearance (pandas.factorize); *n_ids receives the number of distinct strings
static int tsv_intern(fslrc_ctx *ctx, Pipe *P, TsvState *T, const long long *off, const int *len, const unsigned long long *hash,
                      int *id_out, int *first_row_of_id /*nullable, capacity n_rows*/, int *err, int64_t *n_ids_dev) {
    cudaStream_t st = ctx->stream;
    const int n = T->n_rows, TB = 256;
    unsigned cap = 1024; while (cap < 2u * (unsigned)n && cap < (1u << 30)) cap <<= 1;
    unsigned long long *keys; int *first, *slot, *isf, *idat;
    DA(keys, cap); DA(first, cap); DA(slot, n); DA(isf, n); DA(idat, n);
    CK(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * cap, st));
    KL(k_fill<int>, nblk(cap, TB), TB, first, (int64_t)cap, 0x7fffffff);
    KL(tsv::k_tsv_intern_insert, nblk(n, TB), TB, n, hash, keys, first, cap - 1, slot);
    KL(tsv::k_tsv_intern_verify, nblk(n, TB), TB, n, T->text, off, len, slot, first, isf, err);
    int r = xscan(ctx, P, isf, idat, n, n_ids_dev); if (r) return r;
    KL(tsv::k_tsv_intern_ids, nblk(n, TB), TB, n, slot, first, idat, id_out, first_row_of_id);
    return 0;
}
extern "C" {

int fslrc_tsv_open(fslrc_ctx *ctx, const char *text, int64_t n_bytes, uint64_t hash_seed, fslrc_tsv_info *info, void *stream) {
    if (!ctx) return FSLRC_ERR_ARG;
    if (!text || !info || n_bytes <= 0) return fail(ctx, FSLRC_ERR_ARG, "tsv: null or empty input");
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    tsv_free(ctx);
    // ---- header (host): which field holds which column (names of collect_mapping_info.py:176-181)
    const char *names[tsv::W_N] = {"chrom", "rstart", "rend", "qname", "n_alignments", "aln_size", "qstart", "qend", "alignment_score"};
    tsv::Want w; for (int k = 0; k < tsv::W_N; k++) w.col[k] = -1; w.last = -1;
    {
        int64_t e = 0; while (e < n_bytes && text[e] != '\n') e++;
        int f = 0; int64_t p = 0;
        while (p <= e) {
            int64_t q = p; while (q < e && text[q] != '\t') q++;
            for (int k = 0; k < tsv::W_N; k++)
                if ((int64_t)strlen(names[k]) == q - p && memcmp(names[k], text + p, q - p) == 0) { w.col[k] = f; if (f > w.last) w.last = f; }
            f++; p = q + 1;
        }
        for (int k = 0; k < tsv::W_SCORE; k++) if (w.col[k] < 0) return fail(ctx, FSLRC_ERR_ARG, "tsv: header lacks column %s", names[k]);
    }
    TsvState *T = ctx->tsv = new TsvState();
    Pipe Pp; memset((void *)&Pp, 0, sizeof(Pp)); Pipe *P = &Pp;       // (scratch of the scan primitive)
    const bool pad = text[n_bytes - 1] != '\n';
    T->n = n_bytes + (pad ? 1 : 0);
    PA(T->text, T->n + 64);
    CK(cudaMemcpyAsync(T->text, text, n_bytes, cudaMemcpyHostToDevice, st));
    if (pad) CK(cudaMemsetAsync(T->text + n_bytes, '\n', 1, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    // ---- lines
    const int TB = 256;
    const int64_t nchunks = (T->n + tsv::CHUNK - 1) / tsv::CHUNK;
    if (nchunks > 0x7ffffff0LL) { tsv_free(ctx); return fail(ctx, FSLRC_ERR_ARG, "tsv: input too large"); }
    int *cnt, *pre, *err;
    DA(cnt, nchunks); DA(pre, nchunks); DA(err, 1);
    int64_t *dcount; DA(dcount, 4);
    CK(cudaMemsetAsync(err, 0, sizeof(int), st));
    KL(tsv::k_tsv_count, nblk(nchunks, TB), TB, T->text, (long long)T->n, cnt);
    { int r = xscan(ctx, P, cnt, pre, (int)nchunks, dcount); if (r) return r; }
    CK(cudaMemcpyAsync(ctx->h_pin, dcount, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctx->h_pin[0] > 0x7ffffff0LL) { free_all(ctx); tsv_free(ctx); return fail(ctx, FSLRC_ERR_ARG, "tsv: too many lines"); }
    T->n_lines = (int)ctx->h_pin[0];
    T->n_rows = T->n_lines - 1;
    PA(T->line_start, (int64_t)T->n_lines + 1);
    KL(tsv::k_tsv_lines, nblk(nchunks, TB), TB, T->text, (long long)T->n, pre, T->line_start);
    const int n = T->n_rows;
    PA(T->read_id, n); PA(T->chrom, n); PA(T->rstart, n); PA(T->rend, n); PA(T->aln, n); PA(T->qstart, n); PA(T->qend, n); PA(T->naln, n);
    PA(T->q_off, n); PA(T->q_len, n); PA(T->first_row, n);
    T->score = nullptr;
    if (w.col[tsv::W_SCORE] >= 0) PA(T->score, n);
    T->n_reads = 0; T->n_chrom = 0;
    if (n > 0) {
        unsigned long long *qh, *ch; long long *coff; int *clen, *cfirst;
        DA(qh, n); DA(ch, n); DA(coff, n); DA(clen, n); DA(cfirst, n);
        KL(tsv::k_tsv_parse, nblk(n, TB), TB, T->text, T->line_start, n, w, (unsigned long long)hash_seed, T->rstart, T->rend, T->naln, T->aln,
           T->qstart, T->qend, T->score, T->q_off, T->q_len, qh, coff, clen, ch, err);
        int r = tsv_intern(ctx, P, T, T->q_off, T->q_len, qh, T->read_id, T->first_row, err, dcount + 1); if (r) return r;
        r = tsv_intern(ctx, P, T, coff, clen, ch, T->chrom, cfirst, err, dcount + 2); if (r) return r;
        CK(cudaMemcpyAsync(ctx->h_pin, dcount, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_pin + 8, err, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int e = (int)(ctx->h_pin[8] & 0xffffffff);
        if (e) {
            free_all(ctx); tsv_free(ctx);
            if (e & tsv::TE_COLLISION) return fail(ctx, FSLRC_ERR_HASH_COLLISION, "tsv: two different names share a 64-bit hash; call again with another hash_seed");
            if (e & tsv::TE_FIELDS) return fail(ctx, FSLRC_ERR_ARG, "tsv: a line has fewer fields than the header");
            if (e & tsv::TE_RANGE) return fail(ctx, FSLRC_ERR_RANGE, "tsv: an integer field does not fit int32");
            return fail(ctx, FSLRC_ERR_ARG, "tsv: a numeric field is not an integer");
        }
        T->n_reads = (int)ctx->h_pin[1]; T->n_chrom = (int)ctx->h_pin[2];
        // chromosome names (few): offsets of their first rows -> host strings out of the caller's text
        std::vector<int> cf(T->n_chrom);
        CK(cudaMemcpyAsync(cf.data(), cfirst, sizeof(int) * T->n_chrom, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        std::vector<long long> o1(1); std::vector<int> l1(1);
        for (int c = 0; c < T->n_chrom; c++) {
            CK(cudaMemcpyAsync(o1.data(), coff + cf[c], sizeof(long long), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(l1.data(), clen + cf[c], sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            T->chrom_names.emplace_back(text + o1[0], (size_t)l1[0]);
        }
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    free_all(ctx);
    CK(cudaStreamSynchronize(st));
    memset(info, 0, sizeof(*info));
    info->n_rows = T->n_rows; info->n_reads = T->n_reads; info->n_chrom = T->n_chrom; info->has_score = T->score != nullptr;
    info->read_id = T->read_id; info->chrom = T->chrom; info->rstart = T->rstart; info->rend = T->rend; info->aln_size = T->aln;
    info->qstart = T->qstart; info->qend = T->qend; info->n_alignments = T->naln; info->alignment_score = T->score;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[0 + 1], ctx->ev[2]) == cudaSuccess) info->parse_ms = ms; else cudaGetLastError();
    return 0;
}
int fslrc_tsv_chrom_name(fslrc_ctx *ctx, int32_t chrom_id, char *buf, int32_t cap) {
    if (!ctx || !ctx->tsv || !buf || chrom_id < 0 || chrom_id >= ctx->tsv->n_chrom) return FSLRC_ERR_ARG;
    const std::string &s = ctx->tsv->chrom_names[chrom_id];
    if ((int)s.size() + 1 > cap) return FSLRC_ERR_ARG;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
int fslrc_tsv_read_names(fslrc_ctx *ctx, int64_t *offsets, int32_t *lengths) {
    if (!ctx || !ctx->tsv || !offsets || !lengths) return FSLRC_ERR_ARG;
    TsvState *T = ctx->tsv;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int R = T->n_reads;
    if (R == 0) return 0;
    std::vector<int> fr(R);
    CK(cudaMemcpyAsync(fr.data(), T->first_row, sizeof(int) * R, cudaMemcpyDeviceToHost, st));
    std::vector<long long> qo(T->n_rows); std::vector<int> ql(T->n_rows);
    CK(cudaMemcpyAsync(qo.data(), T->q_off, sizeof(long long) * T->n_rows, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ql.data(), T->q_len, sizeof(int) * T->n_rows, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int r = 0; r < R; r++) { offsets[r] = qo[fr[r]]; lengths[r] = ql[fr[r]]; }
    return 0;
}
int fslrc_tsv_write_cluster_bed(fslrc_ctx *ctx, const int32_t *cluster_dev, const int32_t *n_reads_dev, char *out, int64_t cap,
                                int64_t *n_out, void *stream) {
    if (!ctx || !ctx->tsv || !cluster_dev || !n_reads_dev || !n_out) return FSLRC_ERR_ARG;
    TsvState *T = ctx->tsv;
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    Pipe Pp; memset((void *)&Pp, 0, sizeof(Pp)); Pipe *P = &Pp;
    const int L = T->n_lines, TB = 256;
    long long *len, *off; int64_t *tot;
    DA(len, L); DA(off, L); DA(tot, 1);
    KL(tsv::k_tsv_outlen, nblk(L, TB), TB, L, T->line_start, T->read_id, cluster_dev, n_reads_dev, len);
    {
        const int tiles = nblk(L, prims::SC_TILE);
        int r = prim_scratch(ctx, P, sizeof(unsigned long long) * tiles); if (r) return r;
        KL(prims::k_scan_excl<long long>, tiles, prims::SC_THREADS, len, off, L, (unsigned long long *)(P->prim + 256), (unsigned *)P->prim, (long long *)tot);
    }
    CK(cudaMemcpyAsync(ctx->h_pin, tot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = ctx->h_pin[0];
    int rc = 0;
    if (out) {
        if (cap < *n_out) rc = fail(ctx, FSLRC_ERR_ARG, "tsv: output buffer too small");
        else {
            unsigned char *d_out; DA(d_out, *n_out);
            KL(tsv::k_tsv_emit, nblk((int64_t)L * 32, TB), TB, L, T->text, T->line_start, T->read_id, cluster_dev, n_reads_dev, off, d_out);
            CK(cudaMemcpyAsync(out, d_out, *n_out, cudaMemcpyDeviceToHost, st));
        }
    }
    free_all(ctx);
    CK(cudaStreamSynchronize(st));
    return rc;
}
void fslrc_tsv_close(fslrc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    tsv_free(ctx);
}

long long fslrc_launch_count(const fslrc_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fslrc_int_peak(fslrc_ctx *ctx, double *lane_ops_per_s) {
    if (!ctx || !lane_ops_per_s) return FSLRC_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int *d; CK(cudaMalloc(&d, 4));
    const int iters = 4096, blocks = n_sms(ctx) * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_int_peak<<<blocks, threads>>>(64, d);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_int_peak<<<blocks, threads>>>(iters, d);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 3.0;     // add + min/max + xor per chain step
        best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *lane_ops_per_s = best;
    return 0;
}

}  // extern "C"

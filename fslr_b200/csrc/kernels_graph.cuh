// Part of fslr_b200.cu (one translation unit; included after the error flags, LMAX and fslr_b200.h are defined).
// Stages 9-10 and the rows around the path: union-find, cluster numbering, choose_alignment, integer-issue microbenchmark.
#pragma once

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
    if (*(volatile int *)&parent[a] == *(volatile int *)&parent[b]) return;         // (both loads in flight at once) same tree already:
    for (;;) {                                                                      //  most pairs of a PCR family after its first unions
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }                                  // hook the larger root under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}
// entries (a, b) recorded by k_eval / k_pair (a -> b evaluated; only the passing ones matter here): a not saturating;
// b > a -> a tested it (edge); b < a -> edge only if b is saturating and its scan stopped before reaching a (then a's
// query tested the pair, direction a -> b)
#define UE_PER 4
#define UE_THREADS 256
// Only about a third of the list's slots are edges (the rest: padding, pairs of saturating reads, pairs that do not pass), and
// union-find is pointer chasing: what counts is how many chases are in flight.  Every block filters its 1024 slots (entry and
// isP loads batched), compacts the survivors in shared memory and walks them with all lanes busy.
__global__ void __launch_bounds__(UE_THREADS) k_union_entries(unsigned long long n, const int2 *__restrict__ entries, const int *__restrict__ isP,
                                                              Tab t, const int *stop, int *parent, int *ing, unsigned long long *n_edges) {
    __shared__ int2 sE[UE_THREADS * UE_PER * 2];                                    // (a symmetric slot stands for two directed pairs)
    __shared__ int s_warp[UE_THREADS / 32];
    __shared__ int s_ne;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long k0 = (unsigned long long)blockIdx.x * (UE_THREADS * UE_PER) + tid;
    if (tid == 0) s_ne = 0;
    int2 ab[UE_PER];
    int pa[UE_PER], pb[UE_PER];
#pragma unroll
    for (int u = 0; u < UE_PER; u++) {                                              // all entries, then all flags: loads in flight together
        const unsigned long long k = k0 + (unsigned long long)u * UE_THREADS;
        ab[u] = k < n ? __ldg(&entries[k]) : make_int2(-1, -1);
    }
#pragma unroll
    for (int u = 0; u < UE_PER; u++) {                                              // 1: nothing to do for that direction
        const unsigned y = (unsigned)ab[u].y;
        pa[u] = (ab[u].x >= 0 && !(y & EB_NOPASS)) ? __ldg(&isP[ab[u].x]) : 1;
        pb[u] = (ab[u].x >= 0 && (y & EB_SYM) && !(y & EB_NOPASS2)) ? __ldg(&isP[y & QMASK]) : 1;
    }
    int mine = 0;
#pragma unroll
    for (int u = 0; u < UE_PER; u++) mine += !pa[u] + !pb[u];
    int incl = mine;                                                                // block-wide exclusive scan of the survivor counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < UE_THREADS / 32; w++) { const int c = s_warp[w]; if (w < warp) wbase += c; total += c; }
    int at = wbase + incl - mine;
#pragma unroll
    for (int u = 0; u < UE_PER; u++) {                                              // y: b | EB_NOPASS | EB_HEAVY | EB_SYM | EB_NOPASS2
        if (!pa[u]) sE[at++] = make_int2(ab[u].x, ab[u].y & QMASK);
        if (!pb[u]) sE[at++] = make_int2(ab[u].y & QMASK, ab[u].x);
    }
    __syncthreads();
    int ne = 0;
    for (int idx = tid; idx < total; idx += UE_THREADS) {
        const int a = sE[idx].x, b = sE[idx].y;
        bool e = true;
        if (b < a) {
            e = false;
            if (isP[b]) {
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
        if (e) { ing[a] = 1; ing[b] = 1; uf_union(parent, a, b); ne++; }
    }
    for (int o = 16; o; o >>= 1) ne += __shfl_down_sync(0xffffffffu, ne, o);
    if (lane == 0 && ne) atomicAdd(&s_ne, ne);
    __syncthreads();
    if (tid == 0 && s_ne) atomicAdd(n_edges, (unsigned long long)s_ne);             // one counter update per block
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

// Part of fslr_b200.cu (one translation unit; included after the error flags, LMAX and fslr_b200.h are defined).
// Stage 6: the kernel-side table view, the pair evaluation (different_lengths_or_alignments, greedy N-1 intersection,
// Jaccard cutoff) and the read-major pair kernel.
#pragma once

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
#ifndef RP_KL
#define RP_KL 3                  // partners per lane of a replay group handled in one batch (8 * RP_KL partners per batch)
#endif

// Partner record of a saturating read a (replay LIST mode): everything the replay needs to know about partner b without
// touching b's geometry again.  r0 = {b | edge << 31, off_b << 6 | L_b - 1, cg, 0}, r1 = {key[0..3]}:
//   edge    a -> b passes the Jaccard cutoff (cluster.py:218-219),
//   cg      nibble g: 4 | fa* when filling g of b overlaps (closed intervals) a filling of a, fa* = the overlapped filling of
//           a with the highest sorted position: b's scan of g saw a iff it got down to that position,
//   key[fa] the highest sorted position of an interval of b inside the closed band of a's filling fa (-1: none): where
//           a's scan of fa first meets b.
struct PLInfo { unsigned long long off; int n; int pad; };
#define PLF_KNOWN 16            // partner record .w: bit 0 ("b < a and b never breaks: b's query saw the pair") was resolved by k_plist

// heavy_list != NULL: the kernel runs over the reads k_hits listed as heavy (more than 4 fillings, or a hotspot band), every
// rank over all of them (their partner records and isP are needed everywhere; they are few); NULL: over all query reads
// (--overlap <= 0).  Relation entries are only recorded for the reads this shard owns.
template <bool ALLMATCH>
__global__ void __launch_bounds__(PK_WARPS * 32) k_pair(Tab t, const UmaxTab um, int shard, int nshard, const int *__restrict__ heavy_list,
                                                         const unsigned *__restrict__ n_heavy, int *isP, int2 *entries,
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
    const int nq = heavy_list ? (int)*n_heavy : t.Q;
    for (int q0 = blockIdx.x * PK_GROUPS; q0 < nq; q0 += stride) {                 // block-uniform trip count
        const bool live = q0 + grp < nq;
        const int q = live ? (heavy_list ? __ldg(&heavy_list[q0 + grp]) : q0 + grp) : -1;
        const bool mine = live && (nshard <= 1 || ((q >> 8) % nshard) == shard);   // 256-read groups, round robin over ranks
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
        int cnt = 0;                                                               // passing partners so far
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
                int4 c0 = make_int4(0, 0, 0x7fffffff, -1), c1 = make_int4(0, 0, 0, 0);
                if (v) { c0 = __ldg(&t.SR0[p]); c1 = __ldg(&t.SR1[p]); }           // both records of the position, one round trip
                const int b = c0.w & QMASK;
                bool pass = false, part = false, longb = false;
                int wb = 0;
                if (v && b != q && (min(f.z, c0.y) - max(f.y, c0.x)) >= max(f.w, c0.z)) {   // cluster.py:157 for this interval pair
                    int2 *hs = &sHash[grp][b & (PK_HASH - 1)];
                    const int2 hv = *hs;
                    if (hv.x != b || hv.y != q) {                                   // not settled earlier in this read's pass
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
                pass = pass && cnt < t.Tedge;
                const unsigned pc = __ballot_sync(FULL, pass);
                pass = pass && mine;
                const unsigned pm = __ballot_sync(FULL, pass);
                cnt += __popc(pc & gmask);
                if (pm) {
                    const int n = __popc(pm);
                    if (chunk_used + n > PK_CHUNK) {
                        for (int k = chunk_used + lane; k < PK_CHUNK; k += 32) entries[chunk_base + k] = make_int2(-1, -1);
                        if (lane == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)PK_CHUNK);
                        chunk_base = __shfl_sync(FULL, chunk_base, 0);
                        chunk_used = 0;
                        if (chunk_base + PK_CHUNK > cap_entries) { if (lane == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                    }
                    if (pass) entries[chunk_base + chunk_used + __popc(pm & ltmask)] = make_int2(q, (int)((unsigned)b | EB_HEAVY));
                    chunk_used += n;
                    real += (lane == 0) ? n : 0;
                }
                __syncwarp();                                                       // filter updates visible to the next step
            }
        }
        // ---- saturating read: publish its partner records for the replay
        const bool sat = live && cnt >= t.Tedge;
        if (live && gl == 0) isP[q] = sat;
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

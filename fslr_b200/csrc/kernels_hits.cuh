// Part of fslr_b200.cu (one translation unit; included after kernels_pair.cuh).
// Stage 6, dense form: the pair stage of reads with <= 4 fillings and short tight bands ("light" reads, nearly all of
// them) is split into the three kernels the north star names —
//   k_hits   candidate generation: a group of 8 lanes per query read walks the tight bands of the read's fillings with
//            coalesced int4 loads and does nothing but the interval-level test of cluster.py:157; every hit
//            (a, filling fa, sorted position p) goes to a list through warp-aggregated chunk reservations;
//   k_eval   pair test proper, ONE LANE PER HIT, every lane of a warp busy: different_lengths_or_alignments
//            (cluster.py:178-183), the match matrix, the greedy N-1 intersection (cluster.py:152-161) and the per-N
//            Jaccard cutoff (cluster.py:165-170,218-219).  A read pair (a, b) is evaluated exactly once in direction a -> b:
//            at its CANONICAL hit = the lexicographically first matching filling pair (fa, fb), which the lane verifies
//            from the match matrix.  The result overwrites the hit in place
//            (no compaction pass, no atomics on the list): {a, b | flags}, or {-1, -1};
//   k_plist  after the saturating set is known: one lane per recorded pair of a saturating read writes the partner record
//            the replay's LIST mode consumes (keys, cg; same content as k_pair's records).
// Reads with more than 4 fillings or with a tight band longer than PCAP positions (hotspots) are "heavy": they keep the
// sequential per-read kernel k_pair with its early exit (a 500k-read clique must not enumerate 10^11 hits).
#pragma once

#define HK_WARPS 8
#define HK_GROUPS (HK_WARPS * 4)
#define HK_HASH 32              // per-group filter of partners already hit during this read's scan
#define HK_CHUNK 256            // hit slots a warp reserves at a time (multiple of 32: k_eval's warps stay whole)
__device__ __forceinline__ bool shard_owns(int q, int shard, int nshard) { return nshard <= 1 || ((q >> 8) % nshard) == shard; }
__device__ __forceinline__ bool read_is_heavy(int La, const int *__restrict__ rclass, int q) { return La > 4 || __ldg(&rclass[q]) != 0; }

// the heavy reads, listed for k_pair (every rank lists — and later runs — ALL of them: their partner records are needed everywhere)
__global__ void k_heavy_list(Tab t, const int *__restrict__ rclass, int *heavy_list, unsigned *n_heavy) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool heavy = q < t.Q && read_is_heavy((__ldg(&t.RI[q]).w & 63) + 1, rclass, q);
    const unsigned hv = __ballot_sync(0xffffffffu, heavy);
    if (!hv) return;
    unsigned hb = 0;
    if ((threadIdx.x & 31) == 0) hb = atomicAdd(n_heavy, (unsigned)__popc(hv));
    hb = __shfl_sync(0xffffffffu, hb, 0);
    if (heavy) heavy_list[hb + __popc(hv & ((1u << (threadIdx.x & 31)) - 1u))] = q;
}
// The query reads of this shard only: the j-th owned read is q = ((shard + (j >> 8) * nshard) << 8) | (j & 255) (256-read
// groups, round robin over ranks: consecutive ranks = usually one PCR family stay on one rank).
// The tight bands of a read's (<= 4) fillings are walked as ONE concatenated list of (filling, position) pairs, 8 per step:
// no per-filling loop, and the lanes of a group stay busy whatever the individual band lengths are.
// Symmetric mode (no heavy read in the table, *n_heavy == 0): a pair {a, b} is listed only from its lower-ranked read; k_eval
// derives BOTH directed tests from the one match matrix (b -> a is its transpose), so half the hits, half the evaluations.
__global__ void __launch_bounds__(HK_WARPS * 32, 8) k_hits(Tab t, int shard, int nshard, const int *__restrict__ rclass,
                                                           const unsigned *__restrict__ n_heavy, int2 *hits,
                                                           unsigned long long *n_slots, unsigned long long cap, int *err) {
    const bool sym = *n_heavy == 0;
    __shared__ int2 sHash[HK_GROUPS][HK_HASH];                                     // {b, q}: partner b was hit during read q's scan ...
    __shared__ int sHkey[HK_GROUPS][HK_HASH];                                      // ... at filling pair fa << 6 | fb (the lowest listed so far)
    __shared__ int4 sF[HK_GROUPS][4];                                              // the read's fillings {chrom, start, end, T}
    __shared__ int2 sBd[HK_GROUPS][4];                                             // {first band position - its offset in the list, offset of the next filling}
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gl = lane & 7, gsh = lane & 24, grp = w * 4 + (lane >> 3);
    const unsigned ltmask = (1u << lane) - 1u;
    unsigned long long chunk_base = 0;
    int chunk_used = HK_CHUNK;                                                     // nothing reserved yet
    for (int k = gl; k < HK_HASH; k += 8) sHash[grp][k] = make_int2(-1, -1);
    __syncwarp();
    const int stride = gridDim.x * HK_GROUPS;
    const int nblocks256 = (t.Q + 255) >> 8;
    const int owned = nshard <= 1 ? t.Q : ((nblocks256 - shard + nshard - 1) / nshard) << 8;   // (upper bound: the last group may be short)
    for (int j0 = blockIdx.x * HK_GROUPS; j0 < owned; j0 += stride) {              // block-uniform trip count
        const int j = j0 + grp;
        const int q = nshard <= 1 ? j : (((shard + (j >> 8) * nshard) << 8) | (j & 255));
        const bool inq = j < owned && q < t.Q;
        int4 ri = make_int4(0, 0, 0, 0);
        if (inq) ri = __ldg(&t.RI[q]);
        const int off = (int)((unsigned)ri.w >> 6), La = inq ? (ri.w & 63) + 1 : 0;
        const bool light = inq && !read_is_heavy(La, rclass, q);
        // ---- lane fi < La fetches filling fi and its band; an 8-lane prefix sum of the band lengths lays the list out
        int len = 0;
        int4 f = make_int4(0, 0, 0, 0);
        int2 band = make_int2(1, 0);
        if (light && gl < La) { f = rm0(t, off + gl); band = rm2(t, off + gl); len = band.y - band.x + 1; }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o, 8); if (gl >= o) incl += y; }
        const int total = __shfl_sync(FULL, incl, gsh + 3);                        // (lanes 4-7 of a group hold len = 0)
        __syncwarp();
        if (gl < 4) { sF[grp][gl] = f; sBd[grp][gl] = make_int2(band.x - (incl - len), incl); }
        __syncwarp();
        const int steps = __reduce_max_sync(FULL, (total + 7) >> 3);
        const int e0 = sBd[grp][0].y, e1 = sBd[grp][1].y, e2 = sBd[grp][2].y;
        for (int sidx = 0; sidx < steps; sidx++) {
            const int flat = sidx * 8 + gl;
            const bool v = flat < total;
            int4 c0 = make_int4(0, 0, 0x7fffffff, -1);
            int fi = 0, p = 0;
            if (v) {
                fi = (flat >= e0) + (flat >= e1) + (flat >= e2);
                p = sBd[grp][fi].x + flat;
                c0 = __ldg(&t.SR0[p]);
            }
            const int4 ff = sF[grp][fi];
            const int b = c0.w & QMASK;
            bool hit = v && (sym ? b > q : b != q) && (min(ff.z, c0.y) - max(ff.y, c0.x)) >= max(ff.w, c0.z);   // cluster.py:157, T >= 1
            if (hit) {                                                              // a filter only: k_eval's canonical rule is exact
                const int slot = b & (HK_HASH - 1), key = (fi << 6) | (int)((unsigned)c0.w >> 26);
                const int2 h = sHash[grp][slot];
                if (h.x == b && h.y == q && sHkey[grp][slot] < key) hit = false;   // a hit of (q, b) at a lower filling pair is listed
                else { sHash[grp][slot] = make_int2(b, q); sHkey[grp][slot] = key; }
            }
            const unsigned hm = __ballot_sync(FULL, hit);
            if (hm) {
                const int n = __popc(hm);
                if (chunk_used + n > HK_CHUNK) {
                    for (int k = chunk_used + lane; k < HK_CHUNK; k += 32) hits[chunk_base + k] = make_int2(-1, -1);
                    if (lane == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)HK_CHUNK);
                    chunk_base = __shfl_sync(FULL, chunk_base, 0);
                    chunk_used = 0;
                    if (chunk_base + HK_CHUNK > cap) { if (lane == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                }
                if (hit) hits[chunk_base + chunk_used + __popc(hm & ltmask)] = make_int2((int)((unsigned)q | ((unsigned)fi << 26)), p);
                chunk_used += n;
            }
            __syncwarp();                                                           // filter updates visible to the next step
        }
    }
    if (chunk_used < HK_CHUNK)
        for (int k = chunk_used + lane; k < HK_CHUNK; k += 32) hits[chunk_base + k] = make_int2(-1, -1);
}

// One lane per hit.  La <= 4 always (light reads); partners with more than 4 fillings take the loop over global lists.
#define EV_THREADS 256
#ifndef EV_MINB
#define EV_MINB 4
#endif
__global__ void __launch_bounds__(EV_THREADS, EV_MINB) k_eval(Tab t, const UmaxTab um, int2 *hits, const unsigned *__restrict__ n_heavy,
                                                     const unsigned long long *n_slots, unsigned long long cap, unsigned *cp,
                                                     unsigned long long *n_tests, unsigned long long *n_real) {
    const bool sym = *n_heavy == 0;                                                // symmetric mode: slot i also stands for (b, a): EB_SYM
    __shared__ int s_umax[LMAX + 1];
    for (int k = threadIdx.x; k <= LMAX; k += blockDim.x) s_umax[k] = um.v[k];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    unsigned long long n = *n_slots;
    if (n > cap) n = cap;
    unsigned long long tests = 0, real = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * EV_THREADS;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * EV_THREADS + (threadIdx.x & ~31u); i0 < n; i0 += stride) {   // warp-uniform
        const unsigned long long i = i0 + (threadIdx.x & 31);
        int2 h = make_int2(-1, -1);
        if (i < n) h = hits[i];
        int q = 0, fa = 0, fb = 0, p = 0, b = 0, La = 0, Lb = 0, offa = 0, offb = 0;
        bool small = false, gen = false;
        if (h.x >= 0) {
            q = h.x & QMASK; fa = (int)((unsigned)h.x >> 26); p = h.y;
            const int4 c0 = __ldg(&t.SR0[p]), c1 = __ldg(&t.SR1[p]), ria = __ldg(&t.RI[q]);
            b = c0.w & QMASK; fb = (int)((unsigned)c0.w >> 26);
            if (difflen_ok(ria.x, ria.y, ria.z, c1.x, c1.y, c1.z)) {               // cluster.py:178-183
                offa = (int)((unsigned)ria.w >> 6); La = (ria.w & 63) + 1;
                offb = (int)((unsigned)c1.w >> 6); Lb = (c1.w & 63) + 1;
                gen = Lb > 4;
                small = !gen;
            }
        }
        // ---- lists of up to 4 fillings in registers; rows / columns nobody in the warp has are skipped warp-uniformly
        const int mLa = __reduce_max_sync(FULL, small ? La : 0), mLb = __reduce_max_sync(FULL, small ? Lb : 0);
        int4 A[4], B[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            A[k] = make_int4(-1, 0, 0, 0x7fffffff); B[k] = make_int4(-2, 0, 0, 0x7fffffff);
            if (k < mLa) { if (small && k < La) A[k] = rm0(t, offa + k); }
            if (k < mLb) { if (small && k < Lb) B[k] = rm0(t, offb + k); }
        }
        unsigned m[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned r = 0;
            if (k < mLa) {
#pragma unroll
                for (int g = 0; g < 4; g++)
                    if (g < mLb) r |= (matchT<false>(A[k], B[g]) ? 1u : 0u) << g;
            }
            m[k] = r;
        }
        bool canon = small;
        int nmatch = 0;
        {
            unsigned used = 0, row = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < fa && m[k]) canon = false;                                  // an earlier filling of a matches b
                if (k == fa) row = m[k];
                const unsigned avail = m[k] & ~used;                                // greedy first fit (cluster.py:152-161)
                if (avail) { used |= avail & (0u - avail); nmatch++; }
            }
            if (row & ((1u << fb) - 1u)) canon = false;                             // ... or this filling matches an earlier filling of b
        }
        if (gen) {                                                                  // b has more than 4 fillings (rare): global lists
            unsigned long long used = 0;
            canon = true; nmatch = 0;
            for (int k = 0; k < La; k++) {
                const int4 a = rm0(t, offa + k);
                bool taken = false;
                for (int g = 0; g < Lb; g++) {
                    const int4 bq = rm0(t, offb + g);
                    if (matchT<false>(a, bq)) {
                        if (k < fa || (k == fa && g < fb)) canon = false;
                        if (!taken && !((used >> g) & 1ull)) { used |= 1ull << g; nmatch++; taken = true; }
                    }
                }
            }
            if (canon) atomicOr(&cp[q], CP_LONG);                                   // the replay has to WALK this read
        }
        int2 e = make_int2(-1, -1);
        if ((small || gen) && canon && nmatch > 0) {
            const bool pass = (La + Lb - nmatch) <= s_umax[nmatch];                 // cluster.py:165-170,218-219
            tests++; real += pass;
            e = make_int2(q, (int)((unsigned)b | (pass ? 0u : EB_NOPASS)));
            atomicAdd(&cp[q], 0x10000u + (pass ? 1u : 0u));
            if (sym) {                                                              // b -> a: the same matrix read by columns (b's fillings in
                unsigned usedA = 0;                                                 //  order, each taking the first free matching filling of a)
                int nba = 0;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const unsigned col = ((m[0] >> g) & 1u) | (((m[1] >> g) & 1u) << 1) | (((m[2] >> g) & 1u) << 2) | (((m[3] >> g) & 1u) << 3);
                    const unsigned avail = col & ~usedA;
                    if (avail) { usedA |= avail & (0u - avail); nba++; }
                }
                const bool pass2 = (La + Lb - nba) <= s_umax[nba];
                tests++; real += pass2;
                e.y |= (int)(EB_SYM | (pass2 ? 0u : EB_NOPASS2));
                atomicAdd(&cp[b], 0x10000u + (pass2 ? 1u : 0u));
            }
        }
        if (i < n) hits[i] = e;
    }
    for (int o = 16; o; o >>= 1) { tests += __shfl_down_sync(FULL, tests, o); real += __shfl_down_sync(FULL, real, o); }
    if ((threadIdx.x & 31) == 0) { if (tests) atomicAdd(n_tests, tests); if (real) atomicAdd(n_real, real); }
}

// Saturating set of the light reads from their counters (complete for every read after the multi-GPU sum-all-reduce of cp):
// isP, how many partner records the replay will get (plinfo.n: -1 = WALK), and the counter reset for k_plist's fill.
__global__ void k_light_sat(Tab t, const int *__restrict__ rclass, unsigned *cp, int *isP, PLInfo *plinfo, int *plcount,
                            unsigned long long *pl_slots) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= t.Q) return;
    const int wq = __ldg(&t.RI[q]).w, La = (wq & 63) + 1;
    if (read_is_heavy(La, rclass, q)) { plcount[q] = 0; return; }                  // isP / plinfo of heavy reads: k_pair
    const unsigned c = cp[q];
    const int pass = (int)(c & 0xffffu), np = (int)((c >> 16) & 0x7fffu);
    const bool sat = pass >= t.Tedge;
    isP[q] = sat;
    int n = 0;
    if (sat) n = (np <= RP_K && !(c & CP_LONG)) ? np : -1;
    if (n < 0) atomicAdd(pl_slots + 3, 1ull);                                      // (reads the replay has to WALK)
    PLInfo pi; pi.off = 0; pi.n = n; pi.pad = wq;                                  // (pad: where a's fillings live, for k_plist)
    plinfo[q] = pi;
    plcount[q] = n > 0 ? n : 0;
    cp[q] = 0;
}
// where the records of every light saturating read start: after the heavy reads' chunks (*pl_slots) + the scan of plcount
__global__ void k_plinfo(int Q, const int *__restrict__ plcount, const int *__restrict__ ploff, PLInfo *plinfo,
                         unsigned long long *pl_slots, const int64_t *__restrict__ light_total, unsigned long long cap_pl, int *err) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) {
        if (pl_slots[0] + (unsigned long long)*light_total > cap_pl) atomicOr(err, EF_OVERFLOW);
        atomicAdd(pl_slots + 2, (unsigned long long)*light_total);                  // (statistics: records written)
    }
    if (q >= Q) return;
    if (plcount[q] > 0) plinfo[q].off = pl_slots[0] + (unsigned long long)ploff[q];
}
// One lane per recorded pair: the partner record of saturating read a about partner b (see PLInfo / k_pair for the fields)
#define PLT_THREADS 256
#ifndef PLT_MINB
#define PLT_MINB 4               // 64 registers, no spills (5 blocks: 48 registers + spilled words, measured 1.12 vs 1.05 ms for the stage)
#endif
__global__ void __launch_bounds__(PLT_THREADS, PLT_MINB) k_plist(Tab t, const int2 *__restrict__ ent, const unsigned long long *n_slots,
                                                       unsigned long long n_fixed, unsigned long long cap,
                                                       const PLInfo *__restrict__ plinfo, const int *__restrict__ isP, unsigned *cp, int4 *PL, int *err) {
    unsigned long long n = n_slots ? *n_slots : n_fixed;
    if (n > cap) n = cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * PLT_THREADS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * PLT_THREADS + threadIdx.x; i < n; i += stride) {
        const int2 e = __ldg(&ent[i]);
        if (e.x < 0 || ((unsigned)e.y & EB_HEAVY)) continue;
        const int e1 = e.y & QMASK;
        const PLInfo pi0 = plinfo[e.x];                                             // n > 0: saturating, and the replay takes its list
        const int w1 = __ldg(&t.RI[e1]).w;                                          // (both gathers in flight together)
        const bool sym = ((unsigned)e.y & EB_SYM) != 0;                             // the slot also stands for (e1, e.x)
        PLInfo pi1; pi1.off = 0; pi1.n = 0; pi1.pad = 0;
        if (sym) pi1 = plinfo[e1];
        for (int dir = 0; dir < 2; dir++) {
            const PLInfo pi = dir ? pi1 : pi0;
            if (pi.n <= 0) continue;
            const int a = dir ? e1 : e.x, b = dir ? e.x : e1;
            const int wa = pi.pad, wb = dir ? pi0.pad : w1;                         // (pad = RI[.].w for every light read)
            const bool nopass = ((unsigned)e.y & (dir ? EB_NOPASS2 : EB_NOPASS)) != 0;
            const int offa = (int)((unsigned)wa >> 6), La = (wa & 63) + 1, offb = (int)((unsigned)wb >> 6), Lb = (wb & 63) + 1;
            const unsigned slot = atomicAdd(&cp[a], 1u);
            const int seen = (b < a && !__ldg(&isP[b])) ? 1 : 0;                    // b < a and never breaking: its query saw the pair
            if ((int)slot >= pi.n || La > 4 || Lb > 4) { atomicOr(err, EF_OVERFLOW); continue; }   // (cannot happen: counted by k_eval)
            int4 A[4]; int pa[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                A[k] = make_int4(-1, 0, 0, 0); pa[k] = -1;
                if (k < La) { A[k] = rm0(t, offa + k); pa[k] = rm1(t, offa + k).x; }
            }
            int key[4] = {-1, -1, -1, -1};
            unsigned cg = 0;
            for (int gb = 0; gb < Lb; gb++) {
                const int4 i0 = rm0(t, offb + gb);
                const int pg = rm1(t, offb + gb).x;
                int best = -1, bestfa = 0;
#pragma unroll
                for (int fa = 0; fa < 4; fa++) {
                    if (fa < La && A[fa].x == i0.x && A[fa].y <= i0.z && A[fa].z >= i0.y) {   // closed overlap: a scan of one visits the other
                        key[fa] = max(key[fa], pg);
                        if (pa[fa] > best) { best = pa[fa]; bestfa = fa; }
                    }
                }
                if (best >= 0) cg |= (4u | (unsigned)bestfa) << (4 * gb);
            }
            const unsigned long long at = pi.off + slot;
            PL[2 * at] = make_int4((int)((unsigned)b | (nopass ? 0u : 0x80000000u)), wb, (int)cg, PLF_KNOWN | seen);
            PL[2 * at + 1] = make_int4(key[0], key[1], key[2], key[3]);
        }
    }
}
// multi-GPU: this rank's recorded pairs of light saturating reads with partner lists, compacted for the all-gather
__global__ void k_pent_compact(const int2 *__restrict__ ent, unsigned long long n, const PLInfo *__restrict__ plinfo, int2 *out,
                               unsigned long long *n_out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    bool keep = false, keep2 = false;                                               // (a symmetric slot stands for two directed pairs)
    int2 e = make_int2(-1, -1), e2 = make_int2(-1, -1);
    if (i < n) {
        e = ent[i];
        if (e.x >= 0 && !((unsigned)e.y & EB_HEAVY)) {
            const int b = e.y & QMASK;
            keep = plinfo[e.x].n > 0;
            if ((unsigned)e.y & EB_SYM) {
                keep2 = plinfo[b].n > 0;
                e2 = make_int2(b, (int)((unsigned)e.x | (((unsigned)e.y & EB_NOPASS2) ? EB_NOPASS : 0u)));
            }
            e.y = (int)((unsigned)b | ((unsigned)e.y & EB_NOPASS));
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep), m2 = __ballot_sync(0xffffffffu, keep2);
    unsigned long long at = 0;
    if ((threadIdx.x & 31) == 0 && (m | m2)) at = atomicAdd(n_out, (unsigned long long)(__popc(m) + __popc(m2)));
    at = __shfl_sync(0xffffffffu, at, 0);
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    if (keep) out[at + __popc(m & lt)] = e;
    if (keep2) out[at + __popc(m) + __popc(m2 & lt)] = e2;
}

// Part of fslr_b200.cu (one translation unit; included after the error flags, LMAX and fslr_b200.h are defined).
// Stages 7-8: saturating set, run cutting and the replay of saturating reads in query order (LIST and WALK modes).
#pragma once

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
// progress of a walk whose next position is `base`: "everything >= base + 1 is visited".  A step can carry base below the
// chromosome's first position lo (by up to one step); the word must stay a PROGRESS word (<= -2) there — on the first
// chromosome (lo = 0) -2 - (base + 1) would turn non-negative and read as a final stop.
__device__ __forceinline__ int progress_word(int base, int lo) { return -2 - max(base + 1, lo); }
// streak starts: ticket k continues the previous saturating read's streak iff their first fillings reciprocally overlap
// (same PCR family: they depend on each other).  Also marks the stops of every saturating read as unknown.
// RH (k_replay_list): per ticket one 48-byte header {a, RI[a].w, PL offset (40 bits), n}, {pos of fillings 0..3}, {chromosome
// start of fillings 0..3}: everything the replay of read a needs before its partner records, in ONE coalesced round trip
// (instead of the chain plist -> RI / plinfo -> RM -> chrom_lo on the latency-bound path)
__global__ void k_run_flags(int nP, const int *__restrict__ plist, const int4 *__restrict__ RI, const int4 *__restrict__ RM, int *flag,
                            int *stop, int *stopS, const PLInfo *__restrict__ plinfo, const int *__restrict__ chrom_lo, int4 *RH) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    const int a = plist[k];
    const int w = RI[a].w;
    const int off = (int)((unsigned)w >> 6), L = (w & 63) + 1;
    int pos4[4] = {0, 0, 0, 0}, lo4[4] = {0, 0, 0, 0};
    for (int j = 0; j < L; j++) {
        const int pj = RM[2 * (off + j) + 1].x;
        stop[off + j] = STOP_UNSTARTED; stopS[pj] = STOP_UNSTARTED;
        if (j < 4) { pos4[j] = pj; lo4[j] = chrom_lo[RM[2 * (off + j)].x]; }
    }
    if (RH) {
        const PLInfo pi = plinfo[a];
        RH[3 * k] = make_int4(a, w, (int)(unsigned)(pi.off & 0xffffffffull), (int)((unsigned)((pi.off >> 32) & 0xffull) | ((unsigned)pi.n << 8)));
        RH[3 * k + 1] = make_int4(pos4[0], pos4[1], pos4[2], pos4[3]);
        RH[3 * k + 2] = make_int4(lo4[0], lo4[1], lo4[2], lo4[3]);
    }
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
// ---------------------------------------------------------------- job board of the long skip walks (WALK replay)
// The tail of a giant clique is a few reads whose scans skip through the whole island (every earlier read saw them first):
// sequential in the reference, and one warp covers only 256 positions per ~3.5 us round trip.  The chunks of such a walk are
// independent of each other ("does any position of this chunk need an evaluation?"), so the walker posts the next WB_CHUNKS
// chunks on a board in global memory and every warp of the grid that ran out of tickets works them off (wide_help).
// Slot words: [0] lock, [1] seq (published job id), [2] (seq & 0xffff) << 16 | next chunk, [3] chunks of the job,
// [4] chunks done by others, [5] base, [6..11] a, lo, posf, La, start of the filling, sibling flag, [16..43] a's fillings,
// [64..] result per chunk (first position of the chunk that needs an evaluation, 256 = none).
#define WB_SLOTS 64
#ifndef WB_CHUNKS
#define WB_CHUNKS 256
#endif
#define WB_STRIDE (64 + WB_CHUNKS)
struct WJob { int wa, wlo, wposf, wLa, wfy, sibs; int4 A0[4]; int A1x[4]; int2 Achr[4]; };
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int4 ld_relaxed_v4(const int *p) {
    int4 v;
    asm volatile("ld.relaxed.gpu.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one chunk = the 256 positions cb, cb - 1, ..: where (counted from cb) is the first one whose candidate a's scan has to
// evaluate — anything that is no candidate at all, or whose read provably saw a first (its scan of this very interval, or of
// its other filling, already passed a's filling), is skipped.  All loads are issued on clamped indices before any result is
// touched (a conditional block per position makes the compiler consume each record right behind its load: eight serial
// round trips instead of one); the record stays ONE 16-byte load because all four words are used.
__device__ __forceinline__ int wide_chunk(const Tab &t, const int *stopS, const WJob &J, int cb, int lane) {
    const unsigned FULL = 0xffffffffu;
    int wq[8], we[8], ws[8], wv[8], wx[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int pc = max(cb - 32 * k - lane, J.wlo);
        const int4 c0 = __ldg(&t.SR0[pc]); wq[k] = c0.w; we[k] = c0.y; ws[k] = ld_relaxed(&stopS[pc]);
        wx[k] = c0.x | c0.z;                                                       // (start, T >= 0)
        wv[k] = J.sibs ? __ldg(&t.sib[pc]) : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) wq[k] = (cb - 32 * k - lane >= J.wlo && wx[k] >= 0) ? (wq[k] & QMASK) : -1;
    bool needs[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        needs[k] = wq[k] >= 0 && wq[k] != J.wa && we[k] >= J.wfy && !(wq[k] < J.wa && stop_reached(ws[k]) <= J.wposf);
    if (J.sibs) {                                                                  // ... or through its other filling (reads of 2 fillings)
        int sp[8];
        bool anysp = false;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            sp[k] = -1;
            if (needs[k] && wq[k] < J.wa && ((unsigned)wv[k] >> 26) == 1u) sp[k] = wv[k] & QMASK;
            anysp |= sp[k] >= 0;
        }
        if (__any_sync(FULL, anysp)) {
            int cs[8], ce[8], cv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int sc = max(sp[k], 0);
                const int4 c = __ldg(&t.SR0[sc]); cs[k] = c.x; ce[k] = c.y; cv[k] = ld_relaxed(&stopS[sc]);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int reach = stop_reached(cv[k]);
#pragma unroll
                for (int fa = 0; fa < 4; fa++)
                    if (sp[k] >= 0 && fa < J.wLa && sp[k] >= J.Achr[fa].x && sp[k] < J.Achr[fa].y && J.A0[fa].y <= ce[k] &&
                        J.A0[fa].z >= cs[k] && reach <= J.A1x[fa]) needs[k] = false;
            }
        }
    }
    int adv = 256;
#pragma unroll
    for (int k = 7; k >= 0; k--) {
        const unsigned nm = __ballot_sync(FULL, needs[k]);
        if (nm) adv = 32 * k + __ffs(nm) - 1;
    }
    return adv;
}
// a free slot for this warp's walk (-1: none right now)
__device__ __forceinline__ int wide_job_claim(int *wboard, int id, int lane, unsigned *open_jobs) {
    const unsigned FULL = 0xffffffffu;
    int got = -1;
    for (int r = 0; r < WB_SLOTS / 32 && got < 0; r++) {
        const int sidx = (id + r * 32 + lane) % WB_SLOTS;
        const unsigned fm = __ballot_sync(FULL, ld_relaxed(wboard + (size_t)sidx * WB_STRIDE) == 0);
        if (fm) {
            const int l = __ffs(fm) - 1;
            const int cand = __shfl_sync(FULL, sidx, l);
            int ok = 0;
            if (lane == 0) ok = atomicCAS(wboard + (size_t)cand * WB_STRIDE, 0, id + 1) == 0;
            if (__shfl_sync(FULL, ok, 0)) got = cand;
        }
    }
    if (got >= 0 && lane == 0) atomicAdd(open_jobs, 1u);                            // (helpers poll the board while any slot is taken)
    return got;
}
__device__ __forceinline__ void wide_job_release(int *slot, int lane, unsigned *open_jobs) {
    if (lane == 0) { st_release(slot, 0); atomicSub(open_jobs, 1u); }
}
// publish a job: chunks 1 .. nch - 1 of the walk at wbase are up for grabs (chunk 0 is the owner's); returns the job's tag
__device__ __forceinline__ int wide_job_post(int *slot, const WJob &J, int wbase, int nch, int lane, unsigned *open_jobs) {
    int seq = 0;
    if (lane == 0) {
        seq = ld_relaxed(slot + 1) + 1;
        st_relaxed(slot + 4, 0); st_relaxed(slot + 3, nch); st_relaxed(slot + 5, wbase);
        st_relaxed(slot + 6, J.wa); st_relaxed(slot + 7, J.wlo); st_relaxed(slot + 8, J.wposf); st_relaxed(slot + 9, J.wLa);
        st_relaxed(slot + 10, J.wfy); st_relaxed(slot + 11, J.sibs);
#pragma unroll
        for (int fa = 0; fa < 4; fa++) {
            st_relaxed(slot + 16 + 4 * fa, J.A0[fa].x); st_relaxed(slot + 17 + 4 * fa, J.A0[fa].y);
            st_relaxed(slot + 18 + 4 * fa, J.A0[fa].z); st_relaxed(slot + 19 + 4 * fa, J.A0[fa].w);
            st_relaxed(slot + 32 + fa, J.A1x[fa]);
            st_relaxed(slot + 36 + 2 * fa, J.Achr[fa].x); st_relaxed(slot + 37 + 2 * fa, J.Achr[fa].y);
        }
        st_relaxed(slot + 2, (int)((((unsigned)seq & 0xffffu) << 16) | 1u));        // next chunk to hand out: 1
        __threadfence();
        st_release(slot + 1, seq);
    }
    return __shfl_sync(0xffffffffu, seq, 0);
}
// The owner after its own chunk 0 (adv0): works chunks off like everybody else until all are handed out — or closes the job
// as soon as a chunk found a position that needs an evaluation (or the walk is over: cancel) —, waits for the chunks others
// took and folds the results in scan order.  Returns the positions skipped.
__device__ __forceinline__ int wide_job_finish(const Tab &t, const int *stopS, int *slot, const WJob &J, int wbase, int nch, int seq, int adv0,
                                               bool cancel, int lane, unsigned *open_jobs) {
    const unsigned FULL = 0xffffffffu;
    const unsigned tag = ((unsigned)seq & 0xffffu) << 16;
    int mine = 0, limit = nch;
    bool closing = cancel || adv0 < 256;
    for (;;) {
        int v = 0;
        if (closing) {
            if (lane == 0) v = atomicExch(slot + 2, (int)(tag | 0x8000u));          // nothing is handed out any more (grabs see c >= nch)
            limit = min(__shfl_sync(FULL, v, 0) & 0xffff, nch);                     // chunks [1, limit) were
            break;
        }
        if (lane == 0) v = atomicAdd(slot + 2, 1);
        const int c = __shfl_sync(FULL, v, 0) & 0xffff;
        if (c >= nch) break;                                                        // all handed out
        const int r = wide_chunk(t, stopS, J, wbase - 256 * c, lane);
        if (lane == 0) st_relaxed(slot + 64 + c, r);
        mine++;
        if (r < 256) closing = true;
    }
    if (lane == 0) { while (ld_acquire(slot + 4) < limit - 1 - mine) __nanosleep(32); }
    __syncwarp();
    int total = adv0;
    if (!cancel && adv0 == 256) {                                                  // fold in scan order: 32 results per round trip
        for (int h0 = 1; h0 < limit; h0 += 32) {
            const int h = h0 + lane;
            const int r = h < limit ? ld_relaxed(slot + 64 + h) : 0;
            const unsigned stopm = __ballot_sync(FULL, r < 256);                    // (lanes past the limit stop the fold too)
            if (stopm) {
                const int l = __ffs(stopm) - 1;
                total += 256 * l + (h0 + l < limit ? __shfl_sync(FULL, r, l) : 0);
                break;
            }
            total += 256 * 32;
        }
    }
    return total;
}
// a warp without tickets: work chunks off the board until every run of the replay is finished
__device__ __noinline__ void wide_help(const Tab &t, const int *stopS, int *wboard, unsigned *open_jobs, const unsigned *runs_done, unsigned nRuns,
                                       int lane, unsigned long long *n_helped) {
    const unsigned FULL = 0xffffffffu;
    unsigned rot = (blockIdx.x * 7 + (threadIdx.x >> 5)) % WB_SLOTS;
    unsigned long long helped = 0;
    for (;;) {
        int fin = 0, op = 0;
        if (lane == 0) { fin = ld_acquire((const int *)runs_done) >= (int)nRuns; op = ld_relaxed((const int *)open_jobs); }
        fin = __shfl_sync(FULL, fin, 0); op = __shfl_sync(FULL, op, 0);
        if (fin) break;
        if (op <= 0) { __nanosleep(4000); continue; }
        // a slot with chunks left (two slots per lane, rotated so that the helpers spread over the walks)
        int pick = -1;
        for (int r = 0; r < WB_SLOTS / 32 && pick < 0; r++) {
            const int sidx = (rot + r * 32 + lane) % WB_SLOTS;
            const int4 hd = ld_relaxed_v4(wboard + (size_t)sidx * WB_STRIDE);   // {lock, seq, counter, chunks}
            const bool avail = hd.x != 0 && (hd.z & 0xffff) < hd.w;
            const unsigned am = __ballot_sync(FULL, avail);
            if (am) pick = __shfl_sync(FULL, sidx, __ffs(am) - 1);
        }
        if (pick < 0) { __nanosleep(400); continue; }
        rot = (unsigned)pick;
        int *slot = wboard + (size_t)pick * WB_STRIDE;
        for (;;) {                                                                  // chunks of this walk while there are any
            int v = 0;
            if (lane == 0) v = atomicAdd(slot + 2, 1);
            v = __shfl_sync(FULL, v, 0);
            const int c = v & 0xffff;
            const unsigned tag = (unsigned)v >> 16;
            int seq = 0, nch = 0;
            if (lane == 0) {                                                        // the job this chunk belongs to: published yet? gone already?
                for (;;) {
                    seq = ld_acquire(slot + 1);
                    if (((unsigned)seq & 0xffffu) == tag) { nch = ld_relaxed(slot + 3); break; }
                    if ((((unsigned)seq + 1u) & 0xffffu) == tag) { __nanosleep(20); continue; }   // (counter written, seq not yet)
                    nch = 0; break;                                                 // a job of the past
                }
            }
            nch = __shfl_sync(FULL, nch, 0);
            if (c >= nch) break;
            WJob J;
            const int wbase = ld_relaxed(slot + 5);
            J.wa = ld_relaxed(slot + 6); J.wlo = ld_relaxed(slot + 7); J.wposf = ld_relaxed(slot + 8); J.wLa = ld_relaxed(slot + 9);
            J.wfy = ld_relaxed(slot + 10); J.sibs = ld_relaxed(slot + 11);
#pragma unroll
            for (int fa = 0; fa < 4; fa++) {
                J.A0[fa] = make_int4(ld_relaxed(slot + 16 + 4 * fa), ld_relaxed(slot + 17 + 4 * fa), ld_relaxed(slot + 18 + 4 * fa), ld_relaxed(slot + 19 + 4 * fa));
                J.A1x[fa] = ld_relaxed(slot + 32 + fa);
                J.Achr[fa] = make_int2(ld_relaxed(slot + 36 + 2 * fa), ld_relaxed(slot + 37 + 2 * fa));
            }
            const int r = wide_chunk(t, stopS, J, wbase - 256 * c, lane);
            if (lane == 0) { st_relaxed(slot + 64 + c, r); __threadfence(); atomicAdd(slot + 4, 1); }
            helped++;
        }
    }
    if (lane == 0 && helped && n_helped) atomicAdd(n_helped, helped);
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
__device__ __forceinline__ int gmax8(unsigned gmask, int v) { return __reduce_max_sync(gmask, v); }   // one REDUX instead of 3 shuffles
// WALK = false: an instantiation without the band-walking code (half the registers, twice the resident groups) for the
// usual case that every saturating read has partner records
#ifndef RG_MINB
#define RG_MINB 10              // resident blocks per SM the LIST-only instantiation is compiled for (registers <= 65536 / (64 * RG_MINB))
#endif
#ifndef RG_MINB_WALK
#define RG_MINB_WALK 6          // ... and the instantiation with the band-walking code (wide skip steps: 32 loads per lane in flight)
#endif
#ifdef FSLRC_WALKPROF
#define FSLRC_WALKPROF_ON 1
#else
#define FSLRC_WALKPROF_ON 0
#endif
#ifndef RG_PARK
#define RG_PARK 16              // wide skip steps after which a walk counts as long (its warp stops taking new runs)
#endif
#ifndef RG_REC
#define RG_REC 16               // reads a group remembers the final stops of (direct mapped by rank)
#endif
template <bool ALLMATCH, bool WALK>
__global__ void __launch_bounds__(RG_WARPS * 32, WALK ? RG_MINB_WALK : RG_MINB) k_replay(Tab t, const UmaxTab um, int nP, const int *__restrict__ plist, int nRuns,
                                                           const int *__restrict__ rstart, const int *__restrict__ isP,
                                                           const int4 *__restrict__ PL, const PLInfo *__restrict__ plinfo, int *stop,
                                                           int *stopS, unsigned *ticket, int2 *pedges, unsigned long long *n_slots,
                                                           unsigned long long cap_pedges, unsigned long long *n_tests, int *err,
                                                           unsigned long long *dbg, int *wboard, unsigned *open_jobs, unsigned *runs_done,
                                                           unsigned long long *n_helped) {
    __shared__ int4 sA0[RG_GROUPS][4];
    __shared__ int2 sA1[RG_GROUPS][4];
    __shared__ int sStop[RG_GROUPS][WALK ? LMAX : 1];                              // (WALK mode only)
    __shared__ int4 sP0[RG_GROUPS][RP_K];    // partner records: {b | edge << 31, off_b << 6 | L_b - 1, cg, flags}; flags: 1 visited by a,
    __shared__ int4 sP1[RG_GROUPS][RP_K];    //   4 b saw a first, 8 b did not;  {key[0..3]}
    __shared__ int sKey[RG_GROUPS][RP_K];    // first-visit position in the current filling's scan
    __shared__ int2 sAchr[RG_GROUPS][WALK ? 4 : 1];     // [chrom_lo, chrom_hi) of a's fillings (sibling test of the WALK mode)
    __shared__ int s_umax[LMAX + 1];
    __shared__ int sRecTag[RG_GROUPS][RG_REC];   // the reads this group replayed last (direct mapped by rank) and their final
    __shared__ int4 sRecStop[RG_GROUPS][RG_REC]; //   stops: partners of one run mostly look each other up here, not in global memory
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gl = lane & 7, gsh = lane & 24, grp = w * 4 + (lane >> 3);
    const unsigned gmask = 0xffu << gsh;
    // group state (identical in the 8 lanes of a group)
    int phase = 0;                      // 0 next read, 1 walk: next filling, 2 walk: scanning, 3 finished, 4 list: build, 5 list: filling
    unsigned tk = 0, tk1 = 0;
    int a = 0, offa = 0, La = 0, fi = 0, edges = 0, top = 0, lo = 0, base = 0, posf = 0;
    int nPart = 0;
    bool wide = false, parked = false, hasrun = false;
    int wnch = 1;                       // chunks the group's skip walk asks for per step (doubles while it keeps skipping)
    int wslot = -1;                     // the warp's slot on the job board (-1: none)
    int4 ria = make_int4(0, 0, 0, 0), f = make_int4(0, 0, 0, 0);
    unsigned long long tests = 0, chunk_base = 0;
    unsigned long long d_iter = 0, d_steps = 0, d_stall = 0, d_sleep = 0;
    int d_fsteps = 0, d_fstall = 0;
#ifdef FSLRC_WALKPROF
    long long wp_t0 = 0, wp_cw = 0, wp_cn = 0, wp_it0 = 0, wp_c1 = 0, wp_c2 = 0; int wp_nw = 0, wp_nn = 0, wp_n64 = 0;
#endif
    int chunk_used = RP_CHUNK;
    for (int k = threadIdx.x; k <= LMAX; k += blockDim.x) s_umax[k] = um.v[k];
    __syncthreads();
    for (int k = gl; k < RG_REC; k += 8) sRecTag[grp][k] = -1;
    int wit = 0;
    for (;;) {
        __syncwarp();
        d_iter += lane == 0;
        wit++;
#ifdef FSLRC_WALKPROF
        { const long long now = clock64(); if (wp_it0 && phase == 2) { wp_cn += now - wp_it0; wp_nn++; } wp_it0 = now; }
#endif
        if (WALK) {
            // ---- every group of this warp that still has work is skipping through a long band (the tail of a giant clique):
            // serve one of them with all 32 lanes, 256 positions per chunk (a lone 8-lane walker is instruction-latency bound).
            // A group between two runs takes no new ticket while another group of its warp is deep in such a walk (other warps
            // take the tickets): the walker owns the warp.  A walk that keeps skipping posts its next chunks on the job board
            // (wide_job_*): warps that ran out of tickets anywhere on the GPU work them off, so the sequential tail of a 500k-read
            // clique advances up to WB_CHUNKS * 256 positions per step instead of 256.
            parked = __ballot_sync(FULL, phase == 2 && wide && d_fsteps > RG_PARK) != 0 && phase == 0 && tk == tk1;
            const unsigned actm = __ballot_sync(FULL, phase != 3 && !parked), widem = __ballot_sync(FULL, phase == 2 && wide);
            if (actm != 0 && widem == actm) {
                const unsigned gm = ((actm >> 0) & 1u) | (((actm >> 8) & 1u) << 1) | (((actm >> 16) & 1u) << 2) | (((actm >> 24) & 1u) << 3);
                const int sel = __fns(gm, 0, (wit % __popc(gm)) + 1);              // round robin over the walking groups
                const int src = sel * 8, wgrp = w * 4 + sel;
                WJob J;
                J.wa = __shfl_sync(FULL, a, src); J.wlo = __shfl_sync(FULL, lo, src);
                J.wposf = __shfl_sync(FULL, posf, src); J.wLa = __shfl_sync(FULL, La, src); J.wfy = __shfl_sync(FULL, f.y, src);
                const int wbase = __shfl_sync(FULL, base, src);
                int nch = __shfl_sync(FULL, wnch, src);
                if (wbase >= J.wlo) {
                    const int pmx = __ldg(&t.pmaxS[wbase]);
                    J.sibs = (t.sib && J.wLa <= 4) ? 1 : 0;
#pragma unroll
                    for (int fa = 0; fa < 4; fa++) { J.A0[fa] = sA0[wgrp][fa]; J.A1x[fa] = sA1[wgrp][fa].x; J.Achr[fa] = sAchr[wgrp][fa]; }
                    nch = min(nch, (wbase - J.wlo) / 256 + 1);
                    if (nch > 1 && wslot < 0) wslot = wide_job_claim(wboard, blockIdx.x * RG_WARPS + w, lane, open_jobs);
                    if (wslot < 0) nch = 1;
                    int jseq = 0;
                    if (nch > 1) jseq = wide_job_post(wboard + (size_t)wslot * WB_STRIDE, J, wbase, nch, lane, open_jobs);
                    const int adv = wide_chunk(t, stopS, J, wbase, lane);           // chunk 0 is always the owner's
                    const bool over = pmx < J.wfy;                                 // the walk is over: the normal path closes it
                    int total = adv;
                    if (nch > 1) total = wide_job_finish(t, stopS, wboard + (size_t)wslot * WB_STRIDE, J, wbase, nch, jseq, adv, over, lane, open_jobs);
                    if (over) {
                        if ((lane >> 3) == sel) { wide = false; wnch = 1; }
                        continue;
                    }
                    if ((lane >> 3) == sel) {
                        base -= total;
                        if (total < 256 * nch) { wide = false; wnch = 1; }
                        else wnch = min(4 * nch, WB_CHUNKS);
                        if (gl == 0) { d_steps++; st_relaxed(&stop[offa + fi], progress_word(base, lo)); st_relaxed(&stopS[posf], progress_word(base, lo)); }
                        d_fsteps++;
                    }
#ifdef FSLRC_WALKPROF
                    { const long long now = clock64(); if ((lane >> 3) == sel) { wp_cw += now - wp_it0; wp_nw++; } wp_it0 = 0; }
#endif
                    continue;
                }
            }
            else if (wslot >= 0 && widem == 0) { wide_job_release(wboard + (size_t)wslot * WB_STRIDE, lane, open_jobs); wslot = -1; }
        }
        if (phase == 0 && !(WALK && parked)) {
            if (tk == tk1) {                                                       // run finished: take the next ticket
                unsigned run = 0;
                if (WALK && hasrun && gl == 0) {                                   // (helpers leave when every run is done)
                    __threadfence();
                    const unsigned dn = atomicAdd(runs_done, 1u) + 1u;
#ifdef FSLRC_WALKPROF
                    if (dbg) {                                                     // when were 50 / 90 / 99 / 99.9 / 99.99 / 100 % of the runs finished
                        unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                        const unsigned n = (unsigned)nRuns;
                        if (dn == 1u) dbg[36] = g;                                 // (dbg = cnt + 16: cnt[52..55] and dbg[21..23] are free here)
                        if (dn == n / 2) dbg[37] = g;
                        if (dn == n - n / 10) dbg[23] = g;
                        if (dn == n - n / 100) dbg[38] = g;
                        if (dn == n - n / 1000) dbg[21] = g;
                        if (dn == n - n / 10000) dbg[22] = g;
                        if (dn == n) dbg[39] = g;
                    }
#else
                    (void)dn;
#endif
                }
                hasrun = false;
                if (gl == 0) run = atomicAdd(ticket, 1u);
                run = __shfl_sync(gmask, run, gsh);
                if (run >= (unsigned)nRuns) phase = 3;
                else {
                    tk = (unsigned)__ldg(&rstart[run]);
                    tk1 = (run + 1 < (unsigned)nRuns) ? (unsigned)__ldg(&rstart[run + 1]) : (unsigned)nP;
                    hasrun = true;
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
                if (gl == 0) { sRecTag[grp][a & (RG_REC - 1)] = La <= 4 ? a : -1; sRecStop[grp][a & (RG_REC - 1)] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff); }
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
        if (__all_sync(FULL, phase == 3)) {
            if (WALK) {
                if (wslot >= 0) { wide_job_release(wboard + (size_t)wslot * WB_STRIDE, lane, open_jobs); wslot = -1; }
                wide_help(t, stopS, wboard, open_jobs, runs_done, (unsigned)nRuns, lane, n_helped);
            }
            break;
        }
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
            const bool one = nPart <= 8 * RP_KL;                                   // one batch: its partners stay in registers below
            int4 q0[RP_KL];
            int keyk[RP_KL], ekey[RP_KL];                                          // ekey: key of a reached edge partner, else -1
            for (int jb = 0; jb < nPart; jb += 8 * RP_KL) {                        // 8 * RP_KL partners per batch (usually one batch)
            int sv[RP_KL][4];
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
                    if (sRecTag[grp][b & (RG_REC - 1)] == b) {                               // replayed by this very group a moment ago
                        const int4 c = sRecStop[grp][b & (RG_REC - 1)];
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
                ekey[k] = -1;
                if (keyk[k] >= 0) {
                    if ((q0[k].x & QMASK) > a || (q0[k].w & 8)) {
                        mxReach = max(mxReach, keyk[k]); nEdge += (unsigned)q0[k].x >> 31;
                        if (q0[k].x < 0) ekey[k] = keyk[k];
                    }
                    else if (!(q0[k].w & 4)) mxUn = max(mxUn, keyk[k]);
                }
                if (!one && j < nPart) sKey[grp][j] = keyk[k];
            }
            }
            __syncwarp(gmask);
            // the break (cluster.py:223-224): the first reached partner, in scan order, at which `edges` is >= edge_threshold
            const int need = t.Tedge - edges;
            int brkkey = -1;
            if (need <= 0) brkkey = gmax8(gmask, mxReach);
            else {
                nEdge = __reduce_add_sync(gmask, nEdge);
                if (nEdge >= need) {                                               // the need-th highest edge partner
                    int thr = 0x7fffffff;
                    if (one) {                                                     // selection over registers: one REDUX per round
                        for (int r = 0; r < need; r++) {
                            int m = -1;
#pragma unroll
                            for (int k = 0; k < RP_KL; k++) if (ekey[k] < thr) m = max(m, ekey[k]);
                            thr = gmax8(gmask, m);
                        }
                    } else {
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
                    int bq = 0;
                    if (one) {                                                     // (j0 = 8 k: this lane's k-th partner, still in registers)
                        const int k = j0 >> 3;
                        int key = -1; int4 r0 = make_int4(0, 0, 0, 0);
#pragma unroll
                        for (int kk = 0; kk < RP_KL; kk++) if (kk == k) { key = keyk[kk]; r0 = q0[kk]; }
                        if (j < nPart && key >= 0 && key >= brkkey) {
                            sP0[grp][j].w = r0.w | 1;                              // a's query has now seen this pair
                            emit = r0.x < 0 && ((r0.x & QMASK) > a || (r0.w & 8));
                            bq = r0.x & QMASK;
                        }
                    } else if (j < nPart) {
                        const int key = sKey[grp][j];
                        if (key >= 0 && key >= brkkey) {
                            const int4 r0 = sP0[grp][j];
                            sP0[grp][j].w = r0.w | 1;                              // a's query has now seen this pair
                            emit = r0.x < 0 && ((r0.x & QMASK) > a || (r0.w & 8));
                            bq = r0.x & QMASK;
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
                        if (emit) pedges[chunk_base + chunk_used + __popc(em & ((1u << gl) - 1u))] = make_int2(a, bq);
                        chunk_used += n;
                        ne += n;
                    }
                }
                edges += ne;
                const int stopf = brkkey >= 0 ? brkkey : lo;
                if (gl == 0) { st_relaxed(&stop[offa + fi], stopf); st_relaxed(&stopS[posf], stopf); ((int *)&sRecStop[grp][a & (RG_REC - 1)])[fi] = stopf; }
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
            wide = false; wnch = 1;
            phase = 2;
#ifdef FSLRC_WALKPROF
            { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); wp_t0 = (long long)g; wp_cw = wp_cn = wp_c1 = wp_c2 = 0; wp_nw = wp_nn = wp_n64 = 0; }
#endif
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
                const bool sibs = t.sib && La <= 4;
                int wq[8], we[8], ws[8], wv[8], wx[8];                             // all 24 loads of the step are issued (on clamped indices)
#pragma unroll
                for (int k = 0; k < 8; k++) {                                      //  before any result is touched
                    const int pc = max(base - 8 * k - gl, lo);
                    const int4 c0 = __ldg(&t.SR0[pc]); wq[k] = c0.w; we[k] = c0.y; ws[k] = ld_relaxed(&stopS[pc]);
                    wx[k] = c0.x | c0.z;                                           // (start, T >= 0: keeps the record ONE 16-byte load)
                    wv[k] = sibs ? __ldg(&t.sib[pc]) : 0;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) wq[k] = (base - 8 * k - gl >= lo && wx[k] >= 0) ? (wq[k] & QMASK) : -1;
                bool needs[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    needs[k] = wq[k] >= 0 && wq[k] != a && we[k] >= f.y && !(wq[k] < a && stop_reached(ws[k]) <= posf);
                if (sibs) {                                                        // ... or through its other filling (reads of 2 fillings:
                    int sp[8];                                                     //     the loads of all 8 positions are batched)
                    bool anysp = false;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        sp[k] = -1;
                        if (needs[k] && wq[k] < a && ((unsigned)wv[k] >> 26) == 1u) sp[k] = wv[k] & QMASK;
                        anysp |= sp[k] >= 0;
                    }
                    if ((__ballot_sync(gmask, anysp) >> gsh) & 0xffu) {
                        int cs[8], ce[8], cv[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int sc = max(sp[k], 0);
                            const int4 c = __ldg(&t.SR0[sc]); cs[k] = c.x; ce[k] = c.y; cv[k] = ld_relaxed(&stopS[sc]);
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int reach = stop_reached(cv[k]);
#pragma unroll
                            for (int fa = 0; fa < 4; fa++)
                                if (sp[k] >= 0 && fa < La && sp[k] >= sAchr[grp][fa].x && sp[k] < sAchr[grp][fa].y && sA0[grp][fa].y <= ce[k] &&
                                    sA0[grp][fa].z >= cs[k] && reach <= sA1[grp][fa].x) needs[k] = false;
                        }
                    }
                }
#pragma unroll
                for (int k = 7; k >= 0; k--) {
                    const unsigned nm = (__ballot_sync(gmask, needs[k]) >> gsh) & 0xffu;
                    if (nm) adv = 8 * k + __ffs(nm) - 1;
                }
                base -= adv;
#ifdef FSLRC_WALKPROF
                wp_n64++;
#endif
                if (adv < 64) wide = false;
                if (gl == 0) { d_steps++; st_relaxed(&stop[offa + fi], progress_word(base, lo)); st_relaxed(&stopS[posf], progress_word(base, lo)); }
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
                    if (gl == 0 && nres) { st_relaxed(&stop[offa + fi], progress_word(base, lo)); st_relaxed(&stopS[posf], progress_word(base, lo)); }
                }
                if (gl == 0) { d_steps++; d_stall += stalled; }
                d_fsteps++; d_fstall += stalled;
                if (dbg && !FSLRC_WALKPROF_ON && stalled && d_fstall == 5000 && (U & 1u) && gl == 0) {   // lane 0 is the undecided candidate
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
                    if (La <= 4) ((int *)&sRecStop[grp][a & (RG_REC - 1)])[fi] = stopf;
                }
#ifdef FSLRC_WALKPROF
                if (dbg && gl == 0 && top - stopf > 100000) {                      // (profiling build: the longest walk by distance)
                    if (atomicMax(dbg + 4, (unsigned long long)(top - stopf)) < (unsigned long long)(top - stopf)) {
#else
                if (dbg && gl == 0 && d_fsteps > 2000) {
                    if (atomicMax(dbg + 4, (unsigned long long)d_fsteps) < (unsigned long long)d_fsteps) {
#endif
                        dbg[5] = a; dbg[6] = fi; dbg[7] = top - lo; dbg[8] = d_fstall; dbg[9] = top - stopf; dbg[10] = edges; dbg[11] = La;
#ifdef FSLRC_WALKPROF
                        { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                          dbg[12] = 1; dbg[13] = g - (unsigned long long)wp_t0; dbg[14] = wp_nw; dbg[15] = wp_cw; dbg[16] = wp_nn; dbg[17] = wp_cn; dbg[18] = wp_n64; dbg[19] = wp_c1; dbg[20] = wp_c2; }
#endif
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

// ---------------------------------------------------------------- LIST-only replay, one WARP per read
// Used when every saturating read has partner records (no WALK read in the table, --overlap > 0): the same state machine as
// k_replay's LIST mode — runs by ticket, one filling's scan per step, non-blocking retry on stops that are not published yet
// — but a whole warp replays one read with ONE partner per lane (two for reads with 33..64 partners), all in registers.
// Every branch is warp-uniform: nothing serialises inside a warp (k_replay's four 8-lane groups run their phases one after the
// other, 15 of 32 lanes active), the break is a handful of warp-wide REDUX rounds, edges leave through one ballot.
#define RL_WARPS 2
#define RL_SLOTS ((RP_K + 31) / 32)
#ifndef RL_MINB
#define RL_MINB 24              // 48 warps per SM at 40 registers (a few spilled words) beat 32 warps at 62: measured 1.21 vs 1.34 ms
#endif
__global__ void __launch_bounds__(RL_WARPS * 32, RL_MINB) k_replay_list(Tab t, int nP, const int4 *__restrict__ RH, int nRuns,
                                                                  const int *__restrict__ rstart, const int *__restrict__ isP,
                                                                  const int4 *__restrict__ PL, int *stop,
                                                                  int *stopS, unsigned *ticket, int2 *pedges, unsigned long long *n_slots,
                                                                  unsigned long long cap_pedges, int *err, unsigned long long *dbg) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    unsigned tk = 0, tk1 = 0;
    unsigned long long chunk_base = 0, d_iter = 0, d_steps = 0, d_stall = 0, d_sleep = 0;
    int chunk_used = RP_CHUNK;                                                     // nothing reserved yet
    for (;;) {
        // ---- next read of the run (or the next run)
        if (tk == tk1) {
            unsigned run = 0;
            if (lane == 0) run = atomicAdd(ticket, 1u);
            run = __shfl_sync(FULL, run, 0);
            if (run >= (unsigned)nRuns) break;
            tk = (unsigned)__ldg(&rstart[run]);
            tk1 = (run + 1 < (unsigned)nRuns) ? (unsigned)__ldg(&rstart[run + 1]) : (unsigned)nP;
        }
        int4 hd = make_int4(0, 0, 0, 0);                                            // the read's header: lanes 0..2 load 16 bytes each
        if (lane < 3) hd = __ldg(&RH[3 * (size_t)tk + lane]);
        const int a = __shfl_sync(FULL, hd.x, 0), wa = __shfl_sync(FULL, hd.y, 0);
        const unsigned hlo = (unsigned)__shfl_sync(FULL, hd.z, 0), hhi = (unsigned)__shfl_sync(FULL, hd.w, 0);
        struct { unsigned long long off; int n; } pi;
        pi.off = ((unsigned long long)(hhi & 0xffu) << 32) | hlo; pi.n = (int)hhi >> 8;
        const int offa = (int)((unsigned)wa >> 6), La = (wa & 63) + 1;
        if (pi.n < 0 || pi.n > RP_K || La > 4) { if (lane == 0) atomicOr(err, EF_OVERFLOW); break; }   // (cannot happen: the host picks k_replay then)
        int posA[4], loA[4];
        posA[0] = __shfl_sync(FULL, hd.x, 1); posA[1] = __shfl_sync(FULL, hd.y, 1); posA[2] = __shfl_sync(FULL, hd.z, 1); posA[3] = __shfl_sync(FULL, hd.w, 1);
        loA[0] = __shfl_sync(FULL, hd.x, 2); loA[1] = __shfl_sync(FULL, hd.y, 2); loA[2] = __shfl_sync(FULL, hd.z, 2); loA[3] = __shfl_sync(FULL, hd.w, 2);
        // one partner per lane and slot: {b | edge << 31, off_b << 6 | L_b - 1, cg, flags}, {key[0..3]}
        int4 p0[RL_SLOTS], p1[RL_SLOTS];
#pragma unroll
        for (int s = 0; s < RL_SLOTS; s++) {
            const int j = s * 32 + lane;
            p0[s] = make_int4(0, 0, 0, 1); p1[s] = make_int4(-1, -1, -1, -1);      // (no partner: "visited", never met)
            if (j < pi.n) { p0[s] = __ldg(&PL[2 * (pi.off + j)]); p1[s] = __ldg(&PL[2 * (pi.off + j) + 1]); }
        }
#pragma unroll
        for (int s = 0; s < RL_SLOTS; s++) {                                        // b < a and never breaking: it saw the pair (k_plist
            const int j = s * 32 + lane;                                            //  resolved it: PLF_KNOWN; k_pair's records are looked up)
            if (j < pi.n) {
                const int b = p0[s].x & QMASK;
                p0[s].w = (p0[s].w & PLF_KNOWN) ? (p0[s].w & 1) : ((b < a && !__ldg(&isP[b])) ? 1 : 0);
            }
        }
        int edges = 0;
        for (int fi = 0; fi < La;) {                                               // one filling's scan per step (retried while it depends on
            d_iter += lane == 0;                                                   //  a stop that is not published yet)
            // where does the scan first meet each partner; did an earlier-ranked partner's own query see a first?
            int key[RL_SLOTS], ekey[RL_SLOTS];
            int mxReach = -1, mxUn = -1, nEdge = 0;
#pragma unroll
            for (int s = 0; s < RL_SLOTS; s++) {
                key[s] = -1; ekey[s] = -1;
                if (s * 32 >= pi.n) continue;                                       // (warp-uniform: most reads have <= 32 partners)
                if (!(p0[s].w & 1)) key[s] = fi == 0 ? p1[s].x : fi == 1 ? p1[s].y : fi == 2 ? p1[s].z : p1[s].w;
                const int b = p0[s].x & QMASK;
                if (key[s] >= 0 && b < a && !(p0[s].w & 12)) {                      // ask b's stops
                    const int offb = (int)((unsigned)p0[s].y >> 6), Lb = (p0[s].y & 63) + 1;
                    int sv[4];
#pragma unroll
                    for (int g = 0; g < 4; g++)
                        sv[g] = (g < Lb && (((unsigned)p0[s].z >> (4 * g)) & 4u)) ? ld_relaxed(&stop[offb + g]) : 0x7fffffff;
                    bool vis = false, unres = false;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const unsigned cgg = ((unsigned)p0[s].z >> (4 * g)) & 15u;
                        if (cgg & 4u) {
                            const int pa = (cgg & 3u) == 0 ? posA[0] : (cgg & 3u) == 1 ? posA[1] : (cgg & 3u) == 2 ? posA[2] : posA[3];
                            if (stop_reached(sv[g]) <= pa) vis = true; else if (sv[g] < 0) unres = true;
                        }
                    }
                    if (vis) p0[s].w |= 4; else if (!unres) p0[s].w |= 8;
                }
                if (key[s] >= 0) {
                    if (b > a || (p0[s].w & 8)) {
                        mxReach = max(mxReach, key[s]);
                        if (p0[s].x < 0) { nEdge++; ekey[s] = key[s]; }
                    } else if (!(p0[s].w & 4)) mxUn = max(mxUn, key[s]);
                }
            }
            // the break (cluster.py:223-224): the first reached partner, in scan order, at which `edges` is >= edge_threshold
            const int need = t.Tedge - edges;
            int brkkey = -1;
            if (need <= 0) brkkey = __reduce_max_sync(FULL, mxReach);
            else if (__reduce_add_sync(FULL, nEdge) >= need) {                     // the need-th highest edge partner
                int thr = 0x7fffffff;
                for (int r = 0; r < need; r++) {
                    int m = -1;
#pragma unroll
                    for (int s = 0; s < RL_SLOTS; s++) if (ekey[s] < thr) m = max(m, ekey[s]);
                    thr = __reduce_max_sync(FULL, m);
                }
                brkkey = thr;
            }
            const int unkey = __reduce_max_sync(FULL, mxUn);
            d_steps += lane == 0;
            if (unkey > brkkey) {                                                  // an undecided partner comes first: retry
                d_stall += lane == 0; d_sleep += lane == 0;
                __nanosleep(100);
                continue;
            }
            // commit: everything met at or above the break has now been seen by a's query
#pragma unroll
            for (int s = 0; s < RL_SLOTS; s++) {
                if (s * 32 >= pi.n) continue;
                bool emit = false;
                if (key[s] >= 0 && key[s] >= brkkey) {
                    emit = p0[s].x < 0 && ((p0[s].x & QMASK) > a || (p0[s].w & 8));
                    p0[s].w |= 1;
                }
                const unsigned em = __ballot_sync(FULL, emit);
                if (em) {
                    const int n = __popc(em);
                    if (chunk_used + n > RP_CHUNK) {                               // reserve a fresh chunk, pad the old one
                        for (int k = chunk_used + lane; k < RP_CHUNK; k += 32) pedges[chunk_base + k] = make_int2(-1, -1);
                        if (lane == 0) chunk_base = atomicAdd(n_slots, (unsigned long long)RP_CHUNK);
                        chunk_base = __shfl_sync(FULL, chunk_base, 0);
                        chunk_used = 0;
                        if (chunk_base + RP_CHUNK > cap_pedges) { if (lane == 0) atomicOr(err, EF_OVERFLOW); chunk_base = 0; }
                    }
                    if (emit) pedges[chunk_base + chunk_used + __popc(em & ltmask)] = make_int2(a, p0[s].x & QMASK);
                    chunk_used += n;
                    edges += n;
                }
            }
            const int lof = fi == 0 ? loA[0] : fi == 1 ? loA[1] : fi == 2 ? loA[2] : loA[3];
            const int posf = fi == 0 ? posA[0] : fi == 1 ? posA[1] : fi == 2 ? posA[2] : posA[3];
            const int stopf = brkkey >= 0 ? brkkey : lof;
            if (lane == 0) { st_relaxed(&stop[offa + fi], stopf); st_relaxed(&stopS[posf], stopf); }
            fi++;
        }
        tk++;
    }
    if (chunk_used < RP_CHUNK)
        for (int k = chunk_used + lane; k < RP_CHUNK; k += 32) pedges[chunk_base + k] = make_int2(-1, -1);
    if (lane == 0 && dbg) { atomicAdd(dbg, d_iter); atomicAdd(dbg + 1, d_steps); atomicAdd(dbg + 2, d_stall); atomicAdd(dbg + 3, d_sleep); }
}

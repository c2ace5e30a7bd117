// Record boundaries of an uncompressed BAM stream, found on the device (used by fslrc_bam_open_bgzf).
#pragma once
#include "inflate.cuh"

// ---------------------------------------------------------------- record boundaries on the device
// BAM records form a chain (every record starts with its own size), which is sequential from the first record.  To walk
// it in parallel the uncompressed stream is cut into 64 KiB tiles; a warp finds the first plausible record start at or
// after its tile's start (fixed fields in range, name printable and NUL-terminated, sizes consistent, and the next two
// records plausible too), then hops through its tile.  Plausibility is only a guess — exactness comes from the check that
// the record a tile's walk leaves at is the very record the next tile found: with tile 0 starting at the known first
// record, every tile's records are then on the true chain.  Any mismatch makes the caller fall back to a host walk.
namespace bam {
constexpr long long TILE = 65536;

FSLR_HD int rd32h(const unsigned char *p) { return (int)((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24)); }
FSLR_HD bool rec_plausible(const unsigned char *text, long long n, long long p, int n_ref) {
    if (p + 36 > n) return false;
    const unsigned char *r = text + p;
    const long long bs = rd32h(r);
    if (bs < 32 || bs > (1 << 28) || p + 4 + bs > n) return false;
    const int ref = rd32h(r + 4), pos = rd32h(r + 8), l_name = r[12], n_cig = r[16] | (r[17] << 8), l_seq = rd32h(r + 20), nref = rd32h(r + 24), npos = rd32h(r + 28);
    if (ref < -1 || ref >= n_ref || pos < -1 || nref < -1 || nref >= n_ref || npos < -1 || l_name < 1 || l_seq < 0) return false;
    if (32LL + l_name + 4LL * n_cig + ((long long)l_seq + 1) / 2 + l_seq > bs) return false;
    if (r[36 + l_name - 1] != 0) return false;
    for (int i = 0; i < l_name - 1; i++) if (r[36 + i] < 33 || r[36 + i] > 126) return false;
    for (int k = 0; k < n_cig && k < 4; k++) if ((r[36 + l_name + 4 * k] & 15) > 8) return false;
    return true;
}
FSLR_HD bool chain_plausible(const unsigned char *text, long long n, long long p, int n_ref, int depth) {
    for (int d = 0; d < depth; d++) {
        if (p == n) return true;
        if (!rec_plausible(text, n, p, n_ref)) return false;
        p += 4 + (long long)rd32h(text + p);
    }
    return p == n || rec_plausible(text, n, p, n_ref);
}
}  // namespace bam
#ifdef __CUDACC__
namespace bam {
// first[t] = first plausible chain start >= max(t * TILE, first_record) (n when there is none)
__global__ void k_bam_tile_first(const unsigned char *__restrict__ text, long long n, long long first_record, int n_ref, int n_tiles,
                                 long long *first) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    long long s = t * TILE; if (s < first_record) s = first_record;
    long long found = n;
    for (long long base = s; base < n; base += 32) {
        const long long p = base + lane;
        const bool ok = p + 36 <= n && chain_plausible(text, n, p, n_ref, 2);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) { found = base + (__ffs(m) - 1); break; }
    }
    if (lane == 0) first[t] = found;
}
// pass 0 (out == nullptr): count[t] = mapped records starting in the tile, *nrec += all of them, exit_[t] = where the walk
// leaves the tile.  pass 1: the offsets of the mapped records, at base[t].
__global__ void k_bam_tile_walk(const unsigned char *__restrict__ text, long long n, int n_tiles, const long long *__restrict__ first,
                                int *count, unsigned long long *nrec, long long *exit_, const int *__restrict__ base, long long *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const long long end = (t + 1) * TILE < n ? (t + 1) * TILE : n;
    long long p = first[t];
    int c = 0, a = 0;
    int w = out ? base[t] : 0;
    while (p < end && p + 36 <= n) {
        const long long bs = rd32h(text + p);
        if (bs < 32) break;                                              // (cannot happen after the plausibility scan; never loop forever)
        const int flag = text[p + 18] | (text[p + 19] << 8);
        if (!(flag & 4)) { if (out) out[w++] = p; c++; }
        a++;
        p += 4 + bs;
    }
    if (!out) { count[t] = c; exit_[t] = p; if (a) atomicAdd(nrec, (unsigned long long)a); }
}
// the chain check: tile t's walk must leave exactly at the record tile t+1 found; the last tile must leave at the end
__global__ void k_bam_tile_check(int n_tiles, long long n, long long first_record, const long long *__restrict__ first,
                                 const long long *__restrict__ exit_, int *bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const long long want = t + 1 < n_tiles ? first[t + 1] : n;
    if (exit_[t] != want) atomicExch(bad, 1);
    if (t == 0 && first[0] != first_record && first_record < n) atomicExch(bad, 1);
}
}  // namespace bam
#endif

// Host build of fslr_b200/csrc/inflate.cuh for tests/test_inflate_host.py (g++; the same source runs on the device).
#include "../fslr_b200/csrc/inflate.cuh"
extern "C" int host_inflate(const unsigned char *in, long long n_in, unsigned char *out, long long n_out) {
    inflate::Work w;
    return inflate::inflate_stream(in, n_in, out, n_out, w, 0);
}

#include "../fslr_b200/csrc/bam_chain.cuh"
// sequential emulation of k_bam_tile_first / k_bam_tile_walk / k_bam_tile_check over a whole stream: returns the number of
// tiles whose chain check fails (0 = the device path would accept), and the number of records found through the tiles
extern "C" long long host_tile_chain(const unsigned char *text, long long n, long long first_record, int n_ref, long long *n_found) {
    const long long T = bam::TILE;
    const int n_tiles = (int)((n + T - 1) / T);
    long long bad = 0, found = 0, prev_exit = -1;
    for (int t = 0; t < n_tiles; t++) {
        long long s = t * T; if (s < first_record) s = first_record;
        long long f = n;
        for (long long p = s; p < n; p++) if (p + 36 <= n && bam::chain_plausible(text, n, p, n_ref, 2)) { f = p; break; }
        if (t == 0 && f != first_record && first_record < n) bad++;
        if (t > 0 && prev_exit != f) bad++;
        const long long end = (t + 1) * T < n ? (t + 1) * T : n;
        long long p = f;
        while (p < end && p + 36 <= n) { const long long bs = bam::rd32h(text + p); if (bs < 32) break; found++; p += 4 + bs; }
        prev_exit = p;
    }
    if (prev_exit != n) bad++;
    *n_found = found;
    return bad;
}

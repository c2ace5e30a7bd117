// Host build of the WARP-cooperative control flow of fslr_b200/csrc/inflate.cuh with a one-lane "warp" (INF_EMULATE_WARP).
#define INF_EMULATE_WARP 1
#include "../fslr_b200/csrc/inflate.cuh"
extern "C" int host_inflate_warp(const unsigned char *in, long long n_in, unsigned char *out, long long n_out) {
    inflate::Work w;
    return inflate::inflate_stream(in, n_in, out, n_out, w, 0);
}

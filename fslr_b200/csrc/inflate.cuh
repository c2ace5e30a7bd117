// DEFLATE (RFC 1951) decoder for BGZF blocks, written against the RFC (Huffman helpers after puff.c, see below): one decoder instance per BGZF block (<= 64 KiB of
// output), so a BAM file inflates block-parallel on the device and only the COMPRESSED bytes cross PCIe.
// The decoder is a plain sequential function compiled for both host and device (FSLR_HD): the host build is what
// tests/test_inflate_host.py checks against zlib on CPU; the device build runs one decoder per warp (lane 0 decodes,
// the whole warp performs the LZ77 copies and flushes literals), all blocks of the file in flight at once.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FSLR_HD __host__ __device__ __forceinline__
#else
#define FSLR_HD inline
#endif

namespace inflate {

enum { INF_OK = 0, INF_EOF = 1, INF_BADBLOCK = 2, INF_BADCODE = 3, INF_OVERRUN = 4, INF_BADDIST = 5, INF_BADLEN = 6, INF_SIZE = 7 };

struct Bits {                      // LSB-first bit reader over [in, in + n)
    const unsigned char *in;
    long long n, pos;
    unsigned long long buf;
    int cnt;
    int err;
};
FSLR_HD void bits_init(Bits &b, const unsigned char *in, long long n) { b.in = in; b.n = n; b.pos = 0; b.buf = 0; b.cnt = 0; b.err = 0; }
FSLR_HD void bits_fill(Bits &b) {
    if (b.cnt <= 32 && b.pos + 4 <= b.n) {               // four independent byte loads: one memory latency, not four
        const unsigned char *p = b.in + b.pos;
        const unsigned v = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
        b.buf |= (unsigned long long)v << b.cnt; b.cnt += 32; b.pos += 4;
        return;
    }
    while (b.cnt <= 56 && b.pos < b.n) { b.buf |= (unsigned long long)b.in[b.pos++] << b.cnt; b.cnt += 8; }
}
FSLR_HD unsigned bits_get(Bits &b, int need) {          // need <= 16
    if (b.cnt < need) { bits_fill(b); if (b.cnt < need) { b.err = INF_EOF; return 0; } }
    const unsigned v = (unsigned)(b.buf & ((1ull << need) - 1ull));
    b.buf >>= need; b.cnt -= need;
    return v;
}

// canonical Huffman code: count[l] codes of length l, symbols ordered by (length, symbol value).
// huff_build / huff_decode follow the canonical-code construction and the bit-by-bit first-code-per-length decoder of
// zlib's contrib/puff/puff.c (`construct` / `decode`, Mark Adler, zlib licence) — the textbook form of RFC 1951 3.2.2; the
// table lookup for short codes, the warp-cooperative copies and the block driver around them are this file's own.
struct Huff {
    unsigned short *count;         // [16]
    unsigned short *symbol;
};
// returns 0 for a complete code, >0 for an incomplete one, <0 for an over-subscribed one
FSLR_HD int huff_build(Huff &h, const unsigned char *len, int n) {
    unsigned short offs[16];
    for (int l = 0; l < 16; l++) h.count[l] = 0;
    for (int s = 0; s < n; s++) h.count[len[s]]++;
    if (h.count[0] == n) return 0;                       // no codes at all: legal for the distance code
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= h.count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
    for (int s = 0; s < n; s++) if (len[s]) h.symbol[offs[len[s]]++] = (unsigned short)s;
    return left;
}
// one symbol: walk the code lengths, comparing against the first code of each length (RFC 1951 3.2.2)
FSLR_HD int huff_decode(Bits &b, const Huff &h) {
    if (b.cnt < 15) bits_fill(b);
    int code = 0, first = 0, index = 0;
    unsigned long long buf = b.buf;
    const int avail = b.cnt;
    for (int l = 1; l <= 15; l++) {
        if (l > avail) { b.err = INF_EOF; return -1; }
        code |= (int)(buf & 1ull); buf >>= 1;
        const int c = h.count[l];
        if (code - c < first) { b.buf = buf; b.cnt = avail - l; return h.symbol[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    b.err = INF_BADCODE;
    return -1;
}

constexpr int LBITS = 10, DBITS = 8;   // codes up to this many bits decode with one table lookup
struct Work {                      // per-decoder scratch (shared memory on the device)
    unsigned short lsym[288], dsym[32], lcount[16], dcount[16];
    unsigned short ltab[1 << LBITS], dtab[1 << DBITS];   // (symbol << 4) | code length, 0 = longer code: walk the lengths
    unsigned char len[320];
};
// the table entry of the bit pattern `idx` (as it would sit in the low bits of the bit buffer)
FSLR_HD unsigned short tab_entry(const Huff &h, unsigned idx, int K) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= K; l++) {
        code |= (int)(idx & 1u); idx >>= 1;
        const int c = h.count[l];
        if (code - c < first) return (unsigned short)((h.symbol[index + (code - first)] << 4) | l);
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return 0;
}
FSLR_HD int huff_decode(Bits &b, const Huff &h);
FSLR_HD int decode_fast(Bits &b, const Huff &h, const unsigned short *tab, int K) {
    if (b.cnt < 15) bits_fill(b);
    const unsigned e = tab[(unsigned)b.buf & ((1u << K) - 1u)];
    const int l = (int)(e & 15u);
    if (l && l <= b.cnt) { b.buf >>= l; b.cnt -= l; return (int)(e >> 4); }
    return huff_decode(b, h);
}

// length codes 257..285 and distance codes 0..29 (RFC 1951 3.2.5) in closed form; s = symbol - 257 for lengths
FSLR_HD int extra_len(int s) { return s < 8 || s == 28 ? 0 : (s - 4) >> 2; }
FSLR_HD int base_len(int s) { return s < 8 ? 3 + s : s == 28 ? 258 : 3 + ((4 + (s & 3)) << extra_len(s)); }
FSLR_HD int extra_dist(int s) { return s < 4 ? 0 : (s - 2) >> 1; }
FSLR_HD int base_dist(int s) { return s < 4 ? 1 + s : 1 + ((2 + (s & 1)) << extra_dist(s)); }

#if defined(__CUDA_ARCH__)
#define INF_WARP 1
#define INF_LANES 32
#elif defined(INF_EMULATE_WARP)      // host test build of the warp-cooperative control flow with a one-lane "warp"
#define INF_WARP 1
#define INF_LANES 1
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp() {}
#else
#define INF_WARP 0
#endif

// Inflates one raw DEFLATE stream into out[0, n_out).  Returns INF_OK when exactly n_out bytes were produced by a stream
// that ends with a final block.  Device build: called by all 32 lanes of a warp with identical arguments; `lane` is the
// caller's lane, decisions are taken by lane 0 and broadcast.
FSLR_HD int inflate_stream(const unsigned char *in, long long n_in, unsigned char *out, long long n_out, Work &w, int lane) {
    Bits b; bits_init(b, in, n_in);
    long long op = 0;
    int last = 0, rc = INF_OK;
    (void)lane;
    do {
        int type = 0;
#if INF_WARP
        if (lane == 0) {
#endif
        last = (int)bits_get(b, 1);
        type = (int)bits_get(b, 2);
        if (b.err) rc = b.err;
#if INF_WARP
        }
        last = __shfl_sync(0xffffffffu, last, 0); type = __shfl_sync(0xffffffffu, type, 0); rc = __shfl_sync(0xffffffffu, rc, 0);
#endif
        if (rc) return rc;
        if (type == 0) {                                   // stored: skip to a byte boundary, LEN, NLEN, bytes
            long long src = 0; int len = 0;
#if INF_WARP
            if (lane == 0) {
#endif
            const int drop = b.cnt & 7; b.buf >>= drop; b.cnt -= drop;
            const unsigned l = bits_get(b, 16), nl = bits_get(b, 16);
            if (b.err) rc = b.err;
            else if ((l ^ 0xffffu) != nl) rc = INF_BADBLOCK;
            else {
                src = b.pos - (b.cnt >> 3);               // bytes still in the bit buffer belong to the stored data
                len = (int)l;
                if (src + len > b.n) rc = INF_EOF;
                else if (op + len > n_out) rc = INF_OVERRUN;
                else { b.pos = src + len; b.buf = 0; b.cnt = 0; }
            }
#if INF_WARP
            }
            rc = __shfl_sync(0xffffffffu, rc, 0); len = __shfl_sync(0xffffffffu, len, 0); src = __shfl_sync(0xffffffffu, src, 0);
            if (rc) return rc;
            for (int i = lane; i < len; i += INF_LANES) out[op + i] = in[src + i];
            __syncwarp();
#else
            if (rc) return rc;
            for (int i = 0; i < len; i++) out[op + i] = in[src + i];
#endif
            op += len;
            continue;
        }
        if (type == 3) return INF_BADBLOCK;
        Huff hl, hd; hl.symbol = w.lsym; hd.symbol = w.dsym; hl.count = w.lcount; hd.count = w.dcount;
#if INF_WARP
        if (lane == 0) {
#endif
        if (type == 1) {                                   // fixed code (RFC 1951 3.2.6)
            for (int s = 0; s < 144; s++) w.len[s] = 8;
            for (int s = 144; s < 256; s++) w.len[s] = 9;
            for (int s = 256; s < 280; s++) w.len[s] = 7;
            for (int s = 280; s < 288; s++) w.len[s] = 8;
            huff_build(hl, w.len, 288);
            for (int s = 0; s < 30; s++) w.len[s] = 5;
            huff_build(hd, w.len, 30);
        } else {                                           // dynamic code (RFC 1951 3.2.7)
            const int nlen = (int)bits_get(b, 5) + 257, ndist = (int)bits_get(b, 5) + 1, ncode = (int)bits_get(b, 4) + 4;
            const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            if (b.err) rc = b.err;
            else if (nlen > 286 || ndist > 30) rc = INF_BADBLOCK;
            else {
                for (int i = 0; i < 19; i++) w.len[i] = 0;
                for (int i = 0; i < ncode; i++) w.len[order[i]] = (unsigned char)bits_get(b, 3);
                Huff hc; unsigned short csym[19], ccount[16]; hc.symbol = csym; hc.count = ccount;
                if (huff_build(hc, w.len, 19) != 0) rc = INF_BADCODE;    // the code-length code must be complete
                int idx = 0;
                while (!rc && idx < nlen + ndist) {
                    const int sym = huff_decode(b, hc);
                    if (sym < 0) { rc = b.err ? b.err : INF_BADCODE; break; }
                    if (sym < 16) w.len[idx++] = (unsigned char)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (idx == 0) { rc = INF_BADLEN; break; } val = w.len[idx - 1]; rep = 3 + (int)bits_get(b, 2); }
                        else if (sym == 17) rep = 3 + (int)bits_get(b, 3);
                        else rep = 11 + (int)bits_get(b, 7);
                        if (idx + rep > nlen + ndist) { rc = INF_BADLEN; break; }
                        while (rep--) w.len[idx++] = (unsigned char)val;
                    }
                }
                if (!rc && b.err) rc = b.err;
                if (!rc && w.len[256] == 0) rc = INF_BADCODE;             // no end-of-block code
                if (!rc) {
                    // (the code lengths were read into w.len[0 .. nlen + ndist); the literal/length table is built first,
                    // the distance lengths follow it in the same array)
                    int e = huff_build(hl, w.len, nlen);
                    if (e < 0 || (e > 0 && nlen - hl.count[0] != 1)) rc = INF_BADCODE;
                    e = huff_build(hd, w.len + nlen, ndist);
                    if (e < 0 || (e > 0 && ndist - hd.count[0] != 1)) rc = INF_BADCODE;
                }
            }
        }
#if INF_WARP
        }
        rc = __shfl_sync(0xffffffffu, rc, 0);
#endif
        if (rc) return rc;
        // ---- one-lookup tables for the short codes, filled by the whole warp
#if INF_WARP
        __syncwarp();
        for (int i = lane; i < (1 << LBITS); i += INF_LANES) w.ltab[i] = tab_entry(hl, (unsigned)i, LBITS);
        for (int i = lane; i < (1 << DBITS); i += INF_LANES) w.dtab[i] = tab_entry(hd, (unsigned)i, DBITS);
        __syncwarp();
#else
        for (int i = 0; i < (1 << LBITS); i++) w.ltab[i] = tab_entry(hl, (unsigned)i, LBITS);
        for (int i = 0; i < (1 << DBITS); i++) w.dtab[i] = tab_entry(hd, (unsigned)i, DBITS);
#endif
        // ---- symbols of the block
#if INF_WARP
        // lane 0 decodes; literals go straight to memory, every match is broadcast and copied by the warp
        for (;;) {
            int mlen = 0, mdist = 0, done = 0;
            if (lane == 0) {
                for (;;) {
                    const int sym = decode_fast(b, hl, w.ltab, LBITS);
                    if (sym < 0) { rc = b.err ? b.err : INF_BADCODE; break; }
                    if (sym < 256) { if (op >= n_out) { rc = INF_OVERRUN; break; } out[op++] = (unsigned char)sym; continue; }
                    if (sym == 256) { done = 1; break; }
                    const int s = sym - 257;
                    if (s >= 29) { rc = INF_BADCODE; break; }
                    mlen = base_len(s) + (int)bits_get(b, extra_len(s));
                    const int ds = decode_fast(b, hd, w.dtab, DBITS);
                    if (ds < 0 || ds >= 30) { rc = b.err ? b.err : INF_BADCODE; break; }
                    mdist = base_dist(ds) + (int)bits_get(b, extra_dist(ds));
                    if (b.err) { rc = b.err; break; }
                    if (mdist > op) { rc = INF_BADDIST; break; }
                    if (op + mlen > n_out) { rc = INF_OVERRUN; break; }
                    break;
                }
            }
            rc = __shfl_sync(0xffffffffu, rc, 0); done = __shfl_sync(0xffffffffu, done, 0);
            if (rc) return rc;
            op = __shfl_sync(0xffffffffu, op, 0);
            if (done) break;
            mlen = __shfl_sync(0xffffffffu, mlen, 0); mdist = __shfl_sync(0xffffffffu, mdist, 0);
            __syncwarp();                                  // lane 0's literal stores are visible to the copying lanes
            // overlapping copy (dist < len repeats the last `dist` bytes): byte i comes from out[op - dist + (i % dist)]
            for (int i = lane; i < mlen; i += INF_LANES) out[op + i] = out[op - mdist + (i % mdist)];
            __syncwarp();
            op += mlen;
        }
#else
        for (;;) {
            const int sym = decode_fast(b, hl, w.ltab, LBITS);
            if (sym < 0) return b.err ? b.err : INF_BADCODE;
            if (sym < 256) { if (op >= n_out) return INF_OVERRUN; out[op++] = (unsigned char)sym; continue; }
            if (sym == 256) break;
            const int s = sym - 257;
            if (s >= 29) return INF_BADCODE;
            const int mlen = base_len(s) + (int)bits_get(b, extra_len(s));
            const int ds = decode_fast(b, hd, w.dtab, DBITS);
            if (ds < 0 || ds >= 30) return b.err ? b.err : INF_BADCODE;
            const int mdist = base_dist(ds) + (int)bits_get(b, extra_dist(ds));
            if (b.err) return b.err;
            if (mdist > op) return INF_BADDIST;
            if (op + mlen > n_out) return INF_OVERRUN;
            for (int i = 0; i < mlen; i++) { out[op] = out[op - mdist]; op++; }
        }
#endif
    } while (!last);
    return op == n_out ? INF_OK : INF_SIZE;
}

}  // namespace inflate

#ifdef __CUDACC__
namespace inflate {
constexpr int INF_WARPS = 4;       // decoders per CTA
// one warp per BGZF block: in_off/in_len = the raw DEFLATE payload inside the file, out_off/out_len = where it lands
__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate(const unsigned char *__restrict__ comp, const long long *__restrict__ in_off,
                                                            const int *__restrict__ in_len, const long long *__restrict__ out_off,
                                                            const int *__restrict__ out_len, int n_blocks, unsigned char *out, int *err) {
    __shared__ Work w[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * INF_WARPS + warp;
    if (blk >= n_blocks) return;
    const int n_out = out_len[blk];
    int rc = 0;
    if (n_out > 0) rc = inflate_stream(comp + in_off[blk], in_len[blk], out + out_off[blk], n_out, w[warp], lane);
    if (rc && lane == 0) atomicCAS(err, 0, (blk << 4) | rc);
}
}  // namespace inflate
#endif

// K3: resize-back + feather alpha + composite, fused.  Replaces /root/reference/diffuerase.py:70-112
//
//   up    = cv2.resize(inpainted, (W0, H0))                                   (:73)   11-bit fixed point
//   d_in  = cv2.distanceTransform(m_bin,  DIST_L2, 5)                          (:95)
//   d_out = cv2.distanceTransform(~m_bin, DIST_L2, 5)                          (:96)
//   alpha = clip(0.5 + (d_in - d_out) / (2*feather_px), 0, 1)                  (:99-100)
//   out   = u8(clip(rint(alpha*up + (1-alpha)*orig)))                          (:112)  fp32, no FMA
//
// The two full-frame distance transforms are never materialised: alpha only leaves {0,1}
// where the chamfer distance is < feather_px, and a 5x5-chamfer distance < F is decided by
// the mask bits within Chebyshev radius ceil(F)-1 (oracle/prepost.py model_feather_alpha,
// bit-exact against cv2 4.13).  One CTA owns a strip of up to 16 full-width output rows:
//   staging  (TMA variant) one thread brings the strip of `orig` into shared memory with bulk async
//            copies (cp.async.bulk + mbarrier); after the fix-ups the strip leaves with a bulk store,
//            so the 6 B/px pass-through never touches registers.  The register variant (widths that
//            are not multiples of 16) streams orig -> out through 128-bit loads / stores instead.
//   phase 1  mask rows [y0-R, y0+TH+R) -> bit rows in shared memory (128-bit loads, 1 bit/px)
//   phase 2  each thread classifies 16-pixel groups: 5 (or 2R+1) bit-row windows -> per-pixel
//            chamfer class by bit-parallel shifts -> alpha level (12 levels at feather 3)
//   phase 3  every 4-pixel quad with alpha > 0 somewhere becomes a work item in a per-warp queue
//            (warp-scan compaction, items carried over between iterations so worker rounds run
//            with 32 busy lanes); a worker fetches the 8 source-pixel pairs of its quad up front,
//            runs OpenCV's fixed-point bilinear (dp2a horizontal pass, multiply-high vertical
//            pass), blends in non-FMA fp32 and patches the quad in the staged strip.
// HBM traffic per frame: orig 3 + mask 1 (x (TH+2R)/TH from L2) + out 3 B/px, plus the small
// inference-resolution frame, i.e. the algorithmic 7*H0*W0 + 3*h*w of SURVEY section 8d
// (ncu: 4.75 GB per 300 frames = 0.98 x algorithmic).  On dense masks the kernel is bound by the
// issue slots of phase 3 (~98 instructions per blended pixel), on sparse masks by HBM.
#include <math.h>

#include <algorithm>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace vv {

struct Tap {
    int ofs;
    int w;
};
int build_linear_taps(void *workspace, int H, int W, int h, int w, const Tap **xt, const Tap **yt, cudaStream_t st, int which);

constexpr int K3_TH = 16;          // maximum output rows per CTA strip
constexpr int K3_MAX_CLASSES = 31;    // distinct chamfer costs < feather_px (30 at feather 8, the limit of this table): 5 bit planes
constexpr int K3_MAX_ENTRIES = 224;   // window offsets with cost < feather_px, radius <= 7

struct FeatherTable {      // passed by value as a kernel parameter (constant bank, uniform reads)
    float div;             // f32(2 * feather_px)
    int radius;            // window radius R = ceil(F) - 1
    int n;                 // entries below, sorted by ascending cost, all < feather_px (generic path)
    float cost[K3_MAX_ENTRIES];
    int8_t dx[K3_MAX_ENTRIES];
    int8_t dy[K3_MAX_ENTRIES];
    uint8_t cls[K3_MAX_ENTRIES];    // cost class of the entry, 1 .. n_cls in ascending cost (entries are sorted by cost)
    float ccost[K3_MAX_CLASSES + 1];   // cost of class c (index 0 unused)
    int n_cls;
};

__device__ __forceinline__ float alpha_from(float d_in, float d_out, float div) {
    const float a = __fadd_rn(0.5f, __fdiv_rn(__fsub_rn(d_in, d_out), div));
    return fminf(fmaxf(a, 0.f), 1.f);
}

__device__ __forceinline__ uint32_t blend_u8(float a, float one_minus_a, uint32_t up, uint32_t orig) {
    const float v = __fadd_rn(__fmul_rn(a, (float)up), __fmul_rn(one_minus_a, (float)orig));
    // a in (0,1), both inputs in [0,255]: the sum is within half an ulp of [0,255], so np.clip is a no-op
    return rint_u8_bits(v) & 0xffu;                // round-half-even, like np.rint
}

__device__ __forceinline__ int vlin3(int b0, int b1, int h0, int h1) {
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

// 32-bit window of a shared-memory bit row starting at frame column `col0` (may be negative
// down to -32; the row has one zero pad word on each side).
__device__ __forceinline__ uint32_t bit_window(const uint32_t *row, int col0) {
    const int c = col0 + 32;                       // shift into the padded coordinate system
    const int k = c >> 5, off = c & 31;
    return __funnelshift_r(row[k], row[k + 1], off);
}

// The 6 bytes (RGB of source pixel sx, RGB of sx+1) of one inference-resolution row as two words
// (lo = bytes 0..3, hi = bytes 4..7 from the pair's first byte).  Fast path: three aligned 32-bit
// loads + two byte permutes; byte loads near the row end (where sx+1 is clamped to w-1; its weight
// is 0 there) or when the row is not 4-byte aligned.
struct PixelPair {
    uint32_t lo, hi;
};
__device__ __forceinline__ PixelPair load_pixel_pair(const uint8_t *__restrict__ row, int sx, int w, bool row_aligned4) {
    const int o = sx * 3, a4 = o & ~3;
    PixelPair r;
    if (row_aligned4 && a4 + 12 <= w * 3) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + a4);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(o & 3);
        r.lo = __byte_perm(w0, w1, sel);
        r.hi = __byte_perm(w1, w2, sel);
    } else {
        const int o1 = min(sx + 1, w - 1) * 3;
        r.lo = __ldg(row + o) | (__ldg(row + o + 1) << 8) | (__ldg(row + o + 2) << 16) | (__ldg(row + o1) << 24);
        r.hi = __ldg(row + o1 + 1) | (__ldg(row + o1 + 2) << 8);
    }
    return r;
}

// Horizontal pass for channel C of one source row: S[sx][C]*w0 + S[sx+1][C]*w1 as ONE dot product
// (wts = w0 | w1 << 16, both u16; the two source bytes gathered by a byte permute).
template <int C>
__device__ __forceinline__ uint32_t hpass(const PixelPair &pair, uint32_t wts) {
    const uint32_t two = __byte_perm(pair.lo, pair.hi, 0x30 + 0x11 * C);   // bytes C, C+3
    return __dp2a_lo(wts, two, 0u);
}

// Vertical pass (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2 with the weights pre-shifted by 16,
// so that each product-and-shift is one multiply-high.
__device__ __forceinline__ uint32_t vpass(uint32_t b0s, uint32_t b1s, uint32_t h0, uint32_t h1) {
    return (__umulhi(b0s, h0 >> 4) + __umulhi(b1s, h1 >> 4) + 2u) >> 2;
}

// ---- exact x2 up-scale (W0 == 2w, H0 == 2h: 1080p <-> 960x540, the headline configuration) -------
// OpenCV's taps are then the constants 512 / 1536 (of 2048) on both axes and its fixed-point arithmetic
// collapses to small integers:  horizontal  m = far + 3*near  (exact: (s0*c0 + s1*c1) >> 4 == 32*m),
// vertical  out = ((m_a >> 2) + ((3*m_b) >> 2) + 2) >> 2  with m_a from the source row of weight 1/4
// and m_b from the row of weight 3/4.  Border columns / rows use clamped indices, which reproduces
// OpenCV's coefficient clamp on x (4*s >> 2 == s) and its index clip on y.  The 12 channel values of
// a quad are computed as 6 words of two 16-bit lanes.
//
// The 12 source bytes P[i-1], P[i], P[i+1], P[i+2] (i = xq/2, indices clamped) of one row, as 3 words.
__device__ __forceinline__ void x2_load_row(const uint8_t *__restrict__ rowp, int xq, int W0, uint32_t &r0, uint32_t &r1,
                                            uint32_t &r2) {
    // byte offset 3*(i-1) is odd, so the 12 bytes always span four words.  At the two border quads the
    // outermost word would lie outside the row: its load is redirected to a neighbour word and the
    // clamped pixel (P[-1] -> P[0], P[w] -> P[w-1]) is patched in with one byte permute.
    const bool left = xq == 0, right = xq == W0 - 4;
    const int bo = 3 * (xq >> 1) - 3;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(rowp + (bo & ~3));
    const uint32_t w0 = __ldg(p + (left ? 1 : 0)), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + (right ? 2 : 3));
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(bo & 3);
    r0 = __byte_perm(w0, w1, sel), r1 = __byte_perm(w1, w2, sel), r2 = __byte_perm(w2, w3, sel);
    if (left) r0 = __byte_perm(r0, r1, 0x3543);      // [x x x S0] [S1 S2 S3 S4] -> [S0 S1 S2 S0]
    if (right) r2 = __byte_perm(r1, r2, 0x4324);     // [T7 T8 T9 T10] [T11 x x x] -> [T11 T9 T10 T11]
}
// The same 12 bytes addressed through ONE pointer per row: `inw` = the frame as 32-bit words, `row_word` = word
// offset of the source row.  The border words are not loaded at all (predicated off; the patched bytes replace
// them), so there is one 64-bit address computation per row instead of three.
__device__ __forceinline__ void x2_load_row_w(const uint32_t *__restrict__ inw, int row_word, int xq, int W0, uint32_t &r0,
                                              uint32_t &r1, uint32_t &r2) {
    const bool left = xq == 0, right = xq == W0 - 4;
    const int bo = 3 * (xq >> 1) - 3;                                  // -3 at the left border
    const uint32_t *p = inw + (row_word + (bo >> 2));                  // arithmetic shift: word -1 at the left border
    const uint32_t w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    const uint32_t w0 = left ? w1 : __ldg(p), w3 = right ? w2 : __ldg(p + 3);
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(bo & 3);
    r0 = __byte_perm(w0, w1, sel), r1 = __byte_perm(w1, w2, sel), r2 = __byte_perm(w2, w3, sel);
    if (left) r0 = __byte_perm(r0, r1, 0x3543);      // [x x x S0] [S1 S2 S3 S4] -> [S0 S1 S2 S0]
    if (right) r2 = __byte_perm(r1, r2, 0x4324);     // [T7 T8 T9 T10] [T11 x x x] -> [T11 T9 T10 T11]
}
// Split form for software pipelining: the raw words of a row are fetched one worker round ahead (no use of the
// loaded values, so nothing waits), and assembled when the round is blended.
__device__ __forceinline__ void x2_fetch_row(const uint32_t *__restrict__ inw, int row_word, int xq, int W0, uint32_t wv[4]) {
    const bool left = xq == 0, right = xq == W0 - 4;
    const int bo = 3 * (xq >> 1) - 3;
    const uint32_t *p = inw + (row_word + (bo >> 2));
    wv[1] = __ldg(p + 1), wv[2] = __ldg(p + 2);
    wv[0] = left ? 0u : __ldg(p), wv[3] = right ? 0u : __ldg(p + 3);  // the patched bytes below never come from these two
}
__device__ __forceinline__ void x2_assemble_row(const uint32_t wv[4], int xq, int W0, uint32_t &r0, uint32_t &r1, uint32_t &r2) {
    const int bo = 3 * (xq >> 1) - 3;
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(bo & 3);
    r0 = __byte_perm(wv[0], wv[1], sel), r1 = __byte_perm(wv[1], wv[2], sel), r2 = __byte_perm(wv[2], wv[3], sel);
    if (xq == 0) r0 = __byte_perm(r0, r1, 0x3543);           // [x x x S0] [S1 S2 S3 S4] -> [S0 S1 S2 S0]
    if (xq == W0 - 4) r2 = __byte_perm(r1, r2, 0x4324);      // [T7 T8 T9 T10] [T11 x x x] -> [T11 T9 T10 T11]
}
__device__ __forceinline__ void x4_fetch_row(const uint32_t *__restrict__ inw, int row_word, int xq, int W0, uint32_t wv[4]) {
    const int bo = xq == 0 ? 0 : 3 * (xq >> 2) - 3;       // left border: start at pixel 0 and duplicate it on assembly
    // one address per row: the second word always lies inside the row (bo <= 3w - 6), the third one only leaves it at
    // the right border quad, where the assembly replaces the bytes that would come from it (C := B)
    const uint32_t *p = inw + (row_word + (bo >> 2));
    wv[0] = __ldg(p), wv[1] = __ldg(p + 1);
    wv[2] = xq == W0 - 4 ? 0u : __ldg(p + 2);
}
__device__ __forceinline__ void x4_assemble_row(const uint32_t wv[4], int xq, int W0, uint32_t &r0, uint32_t &r1, uint32_t &r2) {
    const bool left = xq == 0, right = xq == W0 - 4;
    const int bo = left ? 0 : 3 * (xq >> 2) - 3;
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(bo & 3);
    const uint32_t s0 = __byte_perm(wv[0], wv[1], sel), s1 = __byte_perm(wv[1], wv[2], sel), s2 = __byte_perm(wv[2], 0u, sel);
    r0 = s0, r1 = s1, r2 = s2;
    if (left) {            // s = [B0 B1 B2 C0] [C1 C2 . .]  ->  A := B
        r0 = __byte_perm(s0, 0u, 0x0210);
        r1 = __byte_perm(s0, s1, 0x4321);
        r2 = s1 >> 8;
    }
    if (right) {           // s = [A0 A1 A2 B0] [B1 B2 . .]  ->  C := B
        r1 = __byte_perm(s0, s1, 0x4354);
        r2 = s1 >> 8;
    }
}
// Horizontal pass of a quad: m[j] holds channel values 2j (low lane) and 2j+1 (high lane) of the
// 12 output values (pixel k/3, channel k%3); near = the source pixel of weight 3/4.
__device__ __forceinline__ void x2_hpass(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t m[6]) {
    const uint32_t e0 = __byte_perm(r0, 0, 0x4140), e1 = __byte_perm(r0, 0, 0x4342);   // bytes (0,1), (2,3)
    const uint32_t e2 = __byte_perm(r1, 0, 0x4140), e3 = __byte_perm(r1, 0, 0x4342);   // (4,5), (6,7)
    const uint32_t e4 = __byte_perm(r2, 0, 0x4140), e5 = __byte_perm(r2, 0, 0x4342);   // (8,9), (10,11)
    const uint32_t b34 = __byte_perm(e1, e2, 0x5432), b78 = __byte_perm(e3, e4, 0x5432);
    m[0] = b34 * 3u + e0;                                                    // near (3,4)  far (0,1)
    m[1] = __byte_perm(e2, e1, 0x7632) * 3u + __byte_perm(e1, e3, 0x5410);   // near (5,3)  far (2,6)
    m[2] = e2 * 3u + b78;                                                    // near (4,5)  far (7,8)
    m[3] = e3 * 3u + b34;                                                    // near (6,7)  far (3,4)
    m[4] = __byte_perm(e4, e3, 0x5410) * 3u + __byte_perm(e2, e4, 0x7632);   // near (8,6)  far (5,9)
    m[5] = b78 * 3u + e5;                                                    // near (7,8)  far (10,11)
}
// Vertical pass on both lanes at once.  Only the low byte of each lane of the result is meaningful:
// the bits the word shifts leak between lanes stay above bit 11 of the low lane and never carry.
__device__ __forceinline__ uint32_t x2_vpass(uint32_t m_a, uint32_t m_b) {
    return ((m_a >> 2) + (((m_b * 3u) >> 2) & 0x3fff3fffu) + 0x00020002u) >> 2;
}

// The same value with fewer ALU-pipe instructions, left in BYTES 1 AND 3 of the word:  (x >> 2) + (y >> 2) ==
// ((x & ~3) + (y & ~3)) >> 2, so out == ((m_a & ~3) + (3 m_b & ~3) + 8) >> 4; the lanes stay below 4096, and the
// final shift is a left shift by 4 instead (multiply pipe), which byte-aligns bits 4..11 of both lanes.
__device__ __forceinline__ uint32_t x2_vpass_b13(uint32_t m_a, uint32_t m_b) {
    return ((m_a & 0xfffcfffcu) + ((m_b * 3u) & 0xfffcfffcu) + 0x00080008u) << 4;
}

// ---- exact x4 horizontal up-scale (W0 == 4w: 4K <- 960 wide, BASELINE config 4) ---------------------------
// OpenCV's horizontal taps are then (768,1280), (256,1792), (1792,256), (1280,768) of 2048 for the four pixels of
// a quad = 256 * (3,5), (1,7), (7,1), (5,3): with A, B, C = source pixels i-1, i, i+1 (i = xq / 4, clamped at the
// borders, which reproduces OpenCV's coefficient clamp) the horizontal pass is m = {3A+5B, A+7B, 7B+C, 5B+3C}
// (exact: h >> 4 == 16 * m, m <= 2040).
//
// The 9 source bytes A0 A1 A2 B0 B1 B2 C0 C1 C2 of one row as [A0 A1 A2 B0] [B1 B2 C0 C1] [C2 . . .].
__device__ __forceinline__ void x4_load_row(const uint8_t *__restrict__ rowp, int xq, int W0, uint32_t &r0, uint32_t &r1,
                                            uint32_t &r2) {
    const bool left = xq == 0, right = xq == W0 - 4;
    const int i = xq >> 2;
    const int bo = left ? 0 : 3 * i - 3;                  // left border: start at pixel 0 and duplicate it below
    const int last = ((W0 >> 2) * 3 - 1) >> 2;            // last word of the row (row bytes = 3 * w, a multiple of 4)
    const int wi = bo >> 2;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(rowp);
    const uint32_t w0 = __ldg(p + wi), w1 = __ldg(p + min(wi + 1, last)), w2 = __ldg(p + min(wi + 2, last));
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(bo & 3);
    const uint32_t s0 = __byte_perm(w0, w1, sel), s1 = __byte_perm(w1, w2, sel), s2 = __byte_perm(w2, 0u, sel);
    r0 = s0, r1 = s1, r2 = s2;
    if (left) {            // s = [B0 B1 B2 C0] [C1 C2 . .]  ->  A := B
        r0 = __byte_perm(s0, 0u, 0x0210);
        r1 = __byte_perm(s0, s1, 0x4321);
        r2 = s1 >> 8;
    }
    if (right) {           // s = [A0 A1 A2 B0] [B1 B2 . .]  ->  C := B
        r1 = __byte_perm(s0, s1, 0x4354);
        r2 = s1 >> 8;
    }
}
// m[j] holds channel values 2j (low lane) and 2j+1 (high lane) of the 12 output values (pixel k/3, channel k%3).
__device__ __forceinline__ void x4_hpass(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t m[6]) {
    const uint32_t e0 = __byte_perm(r0, 0, 0x4140), e1 = __byte_perm(r0, 0, 0x4342);   // (A0,A1) (A2,B0)
    const uint32_t e2 = __byte_perm(r1, 0, 0x4140), e3 = __byte_perm(r1, 0, 0x4342);   // (B1,B2) (C0,C1)
    const uint32_t e4 = __byte_perm(r2, 0, 0x4140);                                    // (C2, .)
    const uint32_t b01 = __byte_perm(e1, e2, 0x5432);                                  // (B0,B1)
    const uint32_t a12 = __byte_perm(e0, e1, 0x5432);                                  // (A1,A2)
    const uint32_t c12 = __byte_perm(e3, e4, 0x5432);                                  // (C1,C2)
    const uint32_t X = __byte_perm(e1, e0, 0x5410);                                    // (A2,A0)
    const uint32_t Y = __byte_perm(e2, e1, 0x7632);                                    // (B2,B0)
    const uint32_t Q = __byte_perm(e4, e3, 0x5410);                                    // (C2,C0)
    m[0] = e0 * 3u + b01 * 5u;                                               // (3A0+5B0, 3A1+5B1)
    m[1] = X + ((X & 0xffffu) << 1) + Y * 5u + ((Y & 0xffff0000u) << 1);     // (3A2+5B2,  A0+7B0)
    m[2] = a12 + e2 * 7u;                                                    // ( A1+7B1,  A2+7B2)
    m[3] = b01 * 7u + e3;                                                    // (7B0+C0,  7B1+C1)
    m[4] = Y * 5u + ((Y & 0xffffu) << 1) + Q + ((Q & 0xffff0000u) << 1);     // (7B2+C2,  5B0+3C0)
    m[5] = e2 * 5u + c12 * 3u;                                               // (5B1+3C1, 5B2+3C2)
}

constexpr int K3_THREADS = 256;    // register pass-through kernel
constexpr int K3_THREADS_TMA = 512;   // TMA-staged kernel (maximum; chosen at launch): the strip occupies shared memory, 2 CTAs per SM
constexpr int K3_QUEUE1 = 128;    // work items (4-pixel quads) per warp iteration and group: 32 lanes x 4 quads

// Work item = one 4-pixel quad (x aligned to 4) that contains at least one pixel with alpha > 0:
//   .x = x | row-in-strip << 16 | need(4) << 20 | inside(4) << 24
//   .y = one LUT index nibble per pixel (SMALL_R): class | inside << 3 for pixel i at bits 4i..4i+3;
//        generic radius: one cost-class byte per pixel (0 = no opposite pixel within the window)

// K3_NT = 16-pixel groups per thread and iteration.
// TMA: the strip of original pixels is brought into shared memory by bulk async copies (one
// elected thread, mbarrier completion), the blended quads are patched into it there, and the whole
// strip leaves with a bulk store - the 6 B/px pass-through never touches registers.  Requires VEC.
// X2: exact x2 up-scale worker (requires VEC, SMALL_R, TMA and 4-byte aligned source rows).
template <bool VEC, bool SMALL_R, int K3_NT, bool TMA, bool X2 = false>
__global__ void __launch_bounds__(TMA ? K3_THREADS_TMA : K3_THREADS, TMA ? 2 : 4)
    k3_upscale_feather_composite(const uint8_t *__restrict__ inp, const uint8_t *__restrict__ orig,
                                 const uint8_t *__restrict__ mask, uint8_t *__restrict__ out,
                                 const Tap *__restrict__ xt, const Tap *__restrict__ yt, int h, int w, int H0, int W0,
                                 int strips_per_frame, int th, const __grid_constant__ FeatherTable ft) {
    extern __shared__ __align__(128) uint32_t smem_base[];
    const int nthreads = TMA ? (int)blockDim.x : K3_THREADS;
    // TMA: [strip th*W0*3 bytes][mbarrier 16 B] then the common part
    const int strip_words = TMA ? (th * W0 * 3) / 4 : 0;
    uint8_t *strip = reinterpret_cast<uint8_t *>(smem_base);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_base + strip_words);
    uint32_t *smem = smem_base + strip_words + (TMA ? 4 : 0);
    const int R = SMALL_R ? 2 : ft.radius;
    const int Wp = (W0 + 31) >> 5;
    const int row_words = Wp + 2;
    const int rows_s = th + 2 * R;
    uint32_t *bits = smem;                                               // [rows_s][row_words]
    float *lut = reinterpret_cast<float *>(smem + rows_s * row_words);   // [16] alpha levels (SMALL_R)
    uint2 *queue = reinterpret_cast<uint2 *>(smem + ((rows_s * row_words + 16 + 1) & ~1)) +
                   (threadIdx.x >> 5) * (K3_QUEUE1 * K3_NT + 32);                  // + carried-over items                        // per-warp work queue
    const int lane = threadIdx.x & 31;

    const long long t = blockIdx.x / strips_per_frame;
    const int y0 = (blockIdx.x % strips_per_frame) * th;
    const uint8_t *mask_t = mask + t * H0 * (long long)W0;
    const uint8_t *orig_t = orig + t * H0 * (long long)W0 * 3;
    uint8_t *out_t = out + t * H0 * (long long)W0 * 3;
    const uint32_t strip_bytes = (uint32_t)(min(th, H0 - y0) * W0 * 3);
    if (TMA) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar, strip_bytes);
            const uint8_t *src = orig_t + (long long)y0 * W0 * 3;
            for (uint32_t off = 0; off < strip_bytes; off += 32768u)
                bulk_g2s(strip + off, src + off, min(32768u, strip_bytes - off), bar);
        }
    }

    // ---------------- phase 1: mask strip -> bit rows
    {
        uint16_t *b16 = reinterpret_cast<uint16_t *>(bits);
        const int halves = 2 * row_words;
        for (int id = threadIdx.x; id < rows_s * halves; id += nthreads) {
            const int i = id / halves, hw = id - i * halves;
            const int y = y0 - R + i, x0 = (hw - 2) * 16;
            uint32_t v = 0;
            if (y >= 0 && y < H0 && x0 >= 0 && x0 < W0) {
                const uint8_t *p = mask_t + (long long)y * W0 + x0;
                if (VEC) {
                    v = nonzero_bits16(ldg128(p));
                } else {
                    const int n = min(16, W0 - x0);
                    for (int k = 0; k < n; ++k) v |= (uint32_t)(p[k] != 0) << k;
                }
            }
            b16[id] = (uint16_t)v;
        }
        if (SMALL_R && threadIdx.x < 16) {
            // alpha levels: index = class (0 = no hit within the window, 1..5 = cost classes
            // 1, 1.4, 2, 2.1969, 2.8) | inside << 3
            const float cost[6] = {8192.f, 1.0f, 1.4f, 2.0f, 2.1969f, __fadd_rn(1.4f, 1.4f)};
            const int cls = threadIdx.x & 7, inside = threadIdx.x >> 3;
            float a = inside ? 1.f : 0.f;
            if (cls <= 5) a = inside ? alpha_from(cost[cls], 0.f, ft.div) : alpha_from(0.f, cost[cls], ft.div);
            lut[threadIdx.x] = a;
        }
    }
    __syncthreads();

    // ---------------- phase 2: 16-pixel groups
    const int G = (W0 + 15) >> 4;
    const uint8_t *inp_t = inp + t * h * (long long)w * 3;
    const bool inp_aligned4 = ((w * 3) % 4 == 0) && ((uintptr_t)inp_t % 4 == 0);
    const bool hard = !SMALL_R && ft.n == 0;
    // which LUT levels have alpha > 0 (SMALL_R): all inside levels, outside levels whose cost < F
    uint32_t lut_pos = 0;
    if (SMALL_R) {
#pragma unroll
        for (int k = 0; k < 16; ++k) lut_pos |= (uint32_t)(lut[k] > 0.f) << k;
    }

    const int n_tasks = th * G;
    bool landed = false;       // TMA: the strip has arrived in shared memory
    // K3_NT tasks (16-pixel groups) per thread and iteration: all their loads are issued before the
    // first store, which keeps enough bytes in flight to cover HBM latency; same trip count for the
    // whole warp.
    const int n_iters = (n_tasks + nthreads * K3_NT - 1) / (nthreads * K3_NT);
    // Work items are carried over between iterations so that the worker rounds below always run
    // with all 32 lanes busy; one extra iteration (no new tasks) drains the remainder.
    int qcount = 0;
    const int step_row = nthreads / G, step_g = nthreads - step_row * G;
    int next_row = (int)threadIdx.x / G, next_g = (int)threadIdx.x - next_row * G;
    for (int it = 0; it <= n_iters; ++it) {
        const bool drain = it == n_iters;
        int row[K3_NT], x0[K3_NT];
        bool active[K3_NT];
        uint4 pa[K3_NT], pb[K3_NT], pc[K3_NT];
#pragma unroll
        for (int k = 0; k < K3_NT; ++k) {
            // task (row, group) = (id / G, id % G) for id = (it * K3_NT + k) * nthreads + tid, advanced
            // incrementally: one integer division per thread instead of one per task
            row[k] = next_row;
            x0[k] = next_g * 16;
            active[k] = next_row < th && y0 + next_row < H0;
            next_row += step_row;
            next_g += step_g;
            if (next_g >= G) next_g -= G, ++next_row;
            if (VEC && !TMA && active[k]) {
                const int pix_off = ((y0 + row[k]) * W0 + x0[k]) * 3;
                pa[k] = ldg128(orig_t + pix_off), pb[k] = ldg128(orig_t + pix_off + 16), pc[k] = ldg128(orig_t + pix_off + 32);
            }
        }
        uint32_t need[K3_NT], L0[K3_NT], L1[K3_NT], L2[K3_NT], M2[K3_NT];
        uint32_t L3[SMALL_R ? 1 : K3_NT], L4[SMALL_R ? 1 : K3_NT];     // generic radius: cost-class planes 3 and 4
#pragma unroll
        for (int k = 0; k < K3_NT; ++k) {
            need[k] = L0[k] = L1[k] = L2[k] = M2[k] = 0;
            if (!SMALL_R) L3[k] = L4[k] = 0;
            if (!active[k]) continue;
            const int y = y0 + row[k];
            const int npx = VEC ? 16 : min(16, W0 - x0[k]);
            // pass the original pixels through; quads with alpha > 0 somewhere are rewritten below
            const int pix_off = (y * W0 + x0[k]) * 3;
            if (TMA) {
                // the strip travels by bulk copy
            } else if (VEC) {
                stg128_stream(out_t + pix_off, pa[k]);
                stg128_stream(out_t + pix_off + 16, pb[k]);
                stg128_stream(out_t + pix_off + 32, pc[k]);
            } else {
                for (int q = 0; q < npx * 3; ++q) out_t[pix_off + q] = orig_t[pix_off + q];
            }
            // window of valid columns: bit j <-> column x0 - 8 + j
            const int c0 = x0[k] - 8;
            const int lo = max(0, -c0), hi = min(32, W0 - c0);
            const uint32_t colvalid = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
            const uint32_t *brow = bits + (row[k] + R) * row_words;   // centre row of this pixel row
            M2[k] = bit_window(brow, c0);
            const uint32_t pxmask = npx >= 16 ? 0xffffu : ((1u << npx) - 1u);
            if (SMALL_R) {
                uint32_t Mw[5], Zw[5];
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int yy = y + d - 2;
                    Mw[d] = bit_window(brow + (d - 2) * row_words, c0);
                    Zw[d] = (yy >= 0 && yy < H0) ? (~Mw[d] & colvalid) : 0u;
                }
                const uint32_t anyM = Mw[0] | Mw[1] | Mw[2] | Mw[3] | Mw[4];
                // the 5x5 window of pixel i covers bits i+6 .. i+10, i.e. bits 6..25 for the whole group
                if (anyM & 0x03ffffc0u) {
                    auto classes = [](const uint32_t *S, uint32_t *hc) {
                        const uint32_t A1 = S[1] | S[3], A0 = S[0] | S[4];
                        hc[0] = (S[2] << 1) | (S[2] >> 1) | A1;                                  // cost 1
                        hc[1] = (A1 << 1) | (A1 >> 1);                                           // 1.4
                        hc[2] = (S[2] << 2) | (S[2] >> 2) | A0;                                  // 2
                        hc[3] = (A0 << 1) | (A0 >> 1) | (A1 << 2) | (A1 >> 2);                   // 2.1969
                        hc[4] = (A0 << 2) | (A0 >> 2);                                           // 2.8
                    };
                    uint32_t hm[5], hz[5], hsel[5];
                    classes(Mw, hm);
                    classes(Zw, hz);
#pragma unroll
                    for (int j = 0; j < 5; ++j) hsel[j] = (hz[j] & M2[k]) | (hm[j] & ~M2[k]);   // inside pixels look for zeros
                    // priority encode: level = first class hit (1..5), 0 if none -> 3 bit planes
                    const uint32_t p1 = hsel[0];
                    const uint32_t p2 = hsel[1] & ~p1;
                    const uint32_t s12 = p1 | hsel[1];
                    const uint32_t p3 = hsel[2] & ~s12;
                    const uint32_t s123 = s12 | hsel[2];
                    const uint32_t p4 = hsel[3] & ~s123;
                    const uint32_t p5 = hsel[4] & ~(s123 | hsel[3]);
                    L0[k] = p1 | p3 | p5, L1[k] = p2 | p3, L2[k] = p4 | p5;
                    // alpha > 0: every inside pixel, and outside pixels whose first hit has alpha > 0
                    uint32_t pos = M2[k];
                    if ((lut_pos >> 1) & 1u) pos |= p1;
                    if ((lut_pos >> 2) & 1u) pos |= p2;
                    if ((lut_pos >> 3) & 1u) pos |= p3;
                    if ((lut_pos >> 4) & 1u) pos |= p4;
                    if ((lut_pos >> 5) & 1u) pos |= p5;
                    need[k] = (pos >> 8) & pxmask;
                }
            } else {
                // Generic radius: the same first-hit search as above, bit-parallel over the 16 pixels of the
                // group, driven by the sorted chamfer table: every table entry (dx, dy) shifts the bit row
                // dy by dx; hits accumulate per cost class, and at the end of a class the still undecided
                // pixels that were hit get that class (5 bit planes L0..L4).  Pixels whose window holds no
                // opposite pixel at all are pruned first (class 0: alpha 1 inside, 0 outside).
                uint32_t anyM = 0, anyZ = 0;
                for (int d = -R; d <= R; ++d) {
                    const int yy = y + d;
                    const uint32_t m = bit_window(brow + d * row_words, c0);
                    const uint32_t z = (yy >= 0 && yy < H0) ? (~m & colvalid) : 0u;
                    uint32_t acc = m, accz = z;
                    for (int j = 1; j <= R; ++j) acc |= (m << j) | (m >> j), accz |= (z << j) | (z >> j);
                    anyM |= acc, anyZ |= accz;
                }
                const uint32_t inside = M2[k];
                uint32_t und = hard ? 0u : (((inside & anyZ) | (~inside & anyM)) & (pxmask << 8));
                uint32_t hcls = 0;
                int cur = 0;
                auto close_class = [&]() {
                    const uint32_t newly = hcls & und;
                    if (cur & 1) L0[k] |= newly;
                    if (cur & 2) L1[k] |= newly;
                    if (cur & 4) L2[k] |= newly;
                    if (cur & 8) L3[k] |= newly;
                    if (cur & 16) L4[k] |= newly;
                    und &= ~hcls;
                    hcls = 0;
                };
                for (int e = 0; e < ft.n && und; ++e) {
                    const int c = ft.cls[e];
                    if (c != cur) {
                        close_class();
                        cur = c;
                    }
                    const int dy = ft.dy[e], dx = ft.dx[e], yy = y + dy;
                    const uint32_t m = bit_window(brow + dy * row_words, c0);
                    const uint32_t z = (yy >= 0 && yy < H0) ? (~m & colvalid) : 0u;
                    const uint32_t sm = dx >= 0 ? (m >> dx) : (m << -dx);      // pixel bit b looks at bit b + dx
                    const uint32_t sz = dx >= 0 ? (z >> dx) : (z << -dx);
                    hcls |= (sz & inside) | (sm & ~inside);
                }
                close_class();
                // alpha > 0: every inside pixel, and outside pixels that found a masked pixel closer than feather_px
                need[k] = ((inside | L0[k] | L1[k] | L2[k] | L3[k] | L4[k]) >> 8) & pxmask;
            }
        }

        // ---- warp-level compaction: every quad with a pixel to blend becomes one work item
        uint32_t any_need = 0, qn = 0;
#pragma unroll
        for (int k = 0; k < K3_NT; ++k) {
            any_need |= need[k];
            qn += ((need[k] & 0x000fu) != 0) + ((need[k] & 0x00f0u) != 0) + ((need[k] & 0x0f00u) != 0) +
                  ((need[k] & 0xf000u) != 0);
        }
        if (__ballot_sync(0xffffffffu, any_need != 0)) {                     // warp-uniform
        int pre = (int)qn;                               // inclusive scan over lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += v;
        }
        int pos = qcount + pre - (int)qn;
        qcount += __shfl_sync(0xffffffffu, pre, 31);
#pragma unroll
        for (int k = 0; k < K3_NT; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t n4 = (need[k] >> (4 * q)) & 15u;
                if (n4) {
                    const int b = 8 + 4 * q;
                    uint2 e;
                    e.x = (uint32_t)(x0[k] + 4 * q) | ((uint32_t)row[k] << 16) | (n4 << 20) | (((M2[k] >> b) & 15u) << 24);
                    // 4x4 bit transpose: planes (l0, l1, l2, inside) x pixels -> one LUT index nibble per pixel
                    uint32_t x4 = ((L0[k] >> b) & 15u) | (((L1[k] >> b) & 15u) << 4) | (((L2[k] >> b) & 15u) << 8) |
                                  (((M2[k] >> b) & 15u) << 12);
                    uint32_t tt = (x4 ^ (x4 >> 3)) & 0x0a0au;
                    x4 ^= tt ^ (tt << 3);
                    tt = (x4 ^ (x4 >> 6)) & 0x00ccu;
                    x4 ^= tt ^ (tt << 6);
                    e.y = x4;
                    if (!SMALL_R) {                 // generic radius: one cost-class byte per pixel
                        uint32_t cb = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int bb = b + i;
                            cb |= (((L0[k] >> bb) & 1u) | (((L1[k] >> bb) & 1u) << 1) | (((L2[k] >> bb) & 1u) << 2) |
                                   (((L3[k] >> bb) & 1u) << 3) | (((L4[k] >> bb) & 1u) << 4))
                                  << (8 * i);
                        }
                        e.y = cb;
                    }
                    queue[pos++] = e;
                }
            }
        }
        }
        __syncwarp();
        if (qcount < 32 && !(drain && qcount > 0)) continue;                 // warp-uniform
        if (TMA && !landed) {
            mbar_wait(bar, 0);
            landed = true;
        }
        while (qcount >= 32 || (drain && qcount > 0)) {
            const int take = min(qcount, 32);
            const int qi = qcount - take + lane;
            qcount -= take;
            if (lane < take) {
                const uint2 item = queue[qi];
                const int xq = item.x & 0xffff, r = (item.x >> 16) & 15;
                const uint32_t n4 = (item.x >> 20) & 15u, in4 = (item.x >> 24) & 15u;
                const int yy = y0 + r;
                if constexpr (X2) {
                    const int j = yy >> 1;                               // source row of weight 3/4
                    const int ja = (yy & 1) ? min(j + 1, h - 1) : max(j - 1, 0);   // source row of weight 1/4
                    uint32_t a0, a1, a2, b0, b1, b2;
                    x2_load_row(inp_t + ja * w * 3, xq, W0, a0, a1, a2);
                    x2_load_row(inp_t + j * w * 3, xq, W0, b0, b1, b2);
                    uint32_t *sp = reinterpret_cast<uint32_t *>(strip + (r * W0 + xq) * 3);
                    uint32_t o[3] = {sp[0], sp[1], sp[2]};
                    uint32_t ma[6], mb[6], up[6];
                    x2_hpass(a0, a1, a2, ma);
                    x2_hpass(b0, b1, b2, mb);
#pragma unroll
                    for (int k = 0; k < 6; ++k) up[k] = x2_vpass(ma[k], mb[k]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((n4 >> i) & 1u) {
                            const float a = lut[(item.y >> (4 * i)) & 15u];      // > 0 by construction of `need`
                            const float na = __fsub_rn(1.f, a);
                            // one branch per pixel: a == 1 (inside, away from the edge) copies the up-scaled bytes
                            if (a < 1.f) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int k = 3 * i + c;                      // byte k of the 12-byte quad
                                    const float v = __fadd_rn(__fmul_rn(a, u8_to_float(up[k >> 1], (k & 1) * 2)),
                                                              __fmul_rn(na, u8_to_float(o[k >> 2], k & 3)));
                                    // round-half-even like np.rint: the sum is in [0, 255], so adding 1.5 * 2^23
                                    // leaves the rounded integer in the low mantissa byte
                                    const uint32_t rb = __float_as_uint(__fadd_rn(v, 12582912.f));
                                    // nibble (k & 3) of the selector takes the new byte, the rest keep o[]
                                    const uint32_t keep = 0x3210u & ~(0xfu << (4 * (k & 3)));
                                    o[k >> 2] = __byte_perm(o[k >> 2], rb, keep | (4u << (4 * (k & 3))));
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int k = 3 * i + c;
                                    const uint32_t keep = 0x3210u & ~(0xfu << (4 * (k & 3)));
                                    o[k >> 2] = __byte_perm(o[k >> 2], up[k >> 1], keep | ((4u + (k & 1) * 2) << (4 * (k & 3))));
                                }
                            }
                        }
                    }
                    sp[0] = o[0], sp[1] = o[1], sp[2] = o[2];
                    continue;
                }
                const Tap ty = yt[yy];
                const uint32_t b0s = (uint32_t)(ty.w & 0xffff) << 16, b1s = (uint32_t)ty.w & 0xffff0000u;
                const int ya = min(max(ty.ofs, 0), h - 1), yb = min(max(ty.ofs + 1, 0), h - 1);
                const uint8_t *r0 = inp_t + ya * w * 3;
                const uint8_t *r1 = inp_t + yb * w * 3;
                const int po = TMA ? (r * W0 + xq) * 3 : (yy * W0 + xq) * 3;   // offset in the strip / in the frame
                int ofs[4];
                uint32_t wts[4], o[3];
                if (VEC) {                                   // W0 % 16 == 0: the quad is whole and 4-byte aligned
                    const uint4 t01 = ldg128(xt + xq), t23 = ldg128(xt + xq + 2);
                    ofs[0] = t01.x, wts[0] = t01.y, ofs[1] = t01.z, wts[1] = t01.w;
                    ofs[2] = t23.x, wts[2] = t23.y, ofs[3] = t23.z, wts[3] = t23.w;
                    if (TMA) {
                        const uint32_t *op = reinterpret_cast<const uint32_t *>(strip + po);
                        o[0] = op[0], o[1] = op[1], o[2] = op[2];
                    } else {
                        const uint32_t *op = reinterpret_cast<const uint32_t *>(orig_t + po);
                        o[0] = __ldg(op), o[1] = __ldg(op + 1), o[2] = __ldg(op + 2);
                    }
                } else {
                    o[0] = o[1] = o[2] = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ofs[i] = 0, wts[i] = 0;
                        if (xq + i < W0) {
                            const Tap tx = xt[xq + i];
                            ofs[i] = tx.ofs, wts[i] = (uint32_t)tx.w;
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                o[(3 * i + c) >> 2] |= (uint32_t)orig_t[po + 3 * i + c] << (8 * ((3 * i + c) & 3));
                        }
                    }
                }
                // Source pixels for all four output pixels are fetched up front (one L2 round trip per
                // quad instead of one per pixel).  ofs[] is non-decreasing, so the aligned fast path is
                // valid for the whole quad when it is valid for the last pixel.
                PixelPair pr0[4], pr1[4];
                const bool fast = inp_aligned4 && (((ofs[3] * 3) & ~3) + 12 <= w * 3);
                if (fast) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int o3 = ofs[i] * 3, a4 = o3 & ~3;
                        const uint32_t *q0 = reinterpret_cast<const uint32_t *>(r0 + a4);
                        const uint32_t *q1 = reinterpret_cast<const uint32_t *>(r1 + a4);
                        const uint32_t u0 = __ldg(q0), u1 = __ldg(q0 + 1), u2 = __ldg(q0 + 2);
                        const uint32_t v0 = __ldg(q1), v1 = __ldg(q1 + 1), v2 = __ldg(q1 + 2);
                        const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(o3 & 3);
                        pr0[i].lo = __byte_perm(u0, u1, sel), pr0[i].hi = __byte_perm(u1, u2, sel);
                        pr1[i].lo = __byte_perm(v0, v1, sel), pr1[i].hi = __byte_perm(v1, v2, sel);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        pr0[i] = load_pixel_pair(r0, ofs[i], w, false);
                        pr1[i] = load_pixel_pair(r1, ofs[i], w, false);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if ((n4 >> i) & 1u) {
                        float a;
                        if (SMALL_R) {
                            a = lut[(item.y >> (4 * i)) & 15u];      // > 0 by construction of `need`
                        } else {
                            const bool inside = (in4 >> i) & 1u;
                            if (hard) {
                                a = inside ? 1.f : 0.f;
                            } else {
                                const uint32_t c = (item.y >> (8 * i)) & 255u;      // cost class of the first hit, 0 = none
                                if (c == 0)
                                    a = inside ? 1.f : 0.f;
                                else
                                    a = inside ? alpha_from(ft.ccost[c], 0.f, ft.div) : alpha_from(0.f, ft.ccost[c], ft.div);
                            }
                        }
                        if (SMALL_R || a > 0.f) {
                            uint32_t cr = vpass(b0s, b1s, hpass<0>(pr0[i], wts[i]), hpass<0>(pr1[i], wts[i]));
                            uint32_t cg = vpass(b0s, b1s, hpass<1>(pr0[i], wts[i]), hpass<1>(pr1[i], wts[i]));
                            uint32_t cb = vpass(b0s, b1s, hpass<2>(pr0[i], wts[i]), hpass<2>(pr1[i], wts[i]));
                            if (a < 1.f) {
                                const float na = __fsub_rn(1.f, a);
                                cr = blend_u8(a, na, cr, byte_of(o[(3 * i) >> 2], (3 * i) & 3));
                                cg = blend_u8(a, na, cg, byte_of(o[(3 * i + 1) >> 2], (3 * i + 1) & 3));
                                cb = blend_u8(a, na, cb, byte_of(o[(3 * i + 2) >> 2], (3 * i + 2) & 3));
                            }
                            const uint32_t rgb = cr | (cg << 8) | (cb << 16);
                            // write the 3 bytes at byte offset 3*i of the 12-byte quad with byte permutes
                            if (i == 0) {
                                o[0] = __byte_perm(o[0], rgb, 0x3654);
                            } else if (i == 1) {
                                o[0] = __byte_perm(o[0], rgb, 0x4210);
                                o[1] = __byte_perm(o[1], rgb, 0x3265);
                            } else if (i == 2) {
                                o[1] = __byte_perm(o[1], rgb, 0x5410);
                                o[2] = __byte_perm(o[2], rgb, 0x3216);
                            } else {
                                o[2] = __byte_perm(o[2], rgb, 0x6540);
                            }
                        }
                    }
                }
                if (VEC) {
                    uint32_t *dp = TMA ? reinterpret_cast<uint32_t *>(strip + po) : reinterpret_cast<uint32_t *>(out_t + po);
                    dp[0] = o[0], dp[1] = o[1], dp[2] = o[2];
                } else {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (xq + k / 3 < W0 && ((n4 >> (k / 3)) & 1u)) out_t[po + k] = (uint8_t)byte_of(o[k >> 2], k & 3);
                }
            }
        }
        __syncwarp();      // the queue is reused by the next iteration
    }
    if (TMA) {
        if (!landed) mbar_wait(bar, 0);
        fence_proxy_async();                 // the patched quads must be visible to the bulk store
        __syncthreads();
        if (threadIdx.x == 0) {
            uint8_t *dst = out_t + (long long)y0 * W0 * 3;
            for (uint32_t off = 0; off < strip_bytes; off += 32768u)
                bulk_s2g(dst + off, strip + off, min(32768u, strip_bytes - off));
            bulk_commit_and_wait_read();     // shared memory must outlive the reads of the store
        }
    }
}

// =====================================================================================================
// k3_fast: the kernel for the production geometry - exact x2 horizontal up-scale (W0 == 2w: 1080p <- 960x540
// or 960x536), feather radius <= 2 (the GUI's feather_px = 3), 16-byte aligned frames.  Same structure as the
// kernel above (TMA-staged strip, bit rows, quad work queue), rebuilt around the instruction counts ncu's
// source view showed for it (profiles/r02_k3_source_breakdown.txt):
//   phase 1  BITS: the dilated mask arrives as the 1-bit plane K1 already produced (32x fewer bytes, no
//            byte -> bit packing); otherwise packed from the u8 mask with one row per warp (no division)
//   phase 2  ROLLING window: a thread owns one 16-pixel column group for RPT consecutive rows, so each bit row
//            is fetched and masked once instead of five times and the column masks are per-task constants
//   phase 3  work items carry position + LUT nibbles only; the blend is BRANCH-FREE over all 12 bytes of a quad in
//            packed fp32 (FMUL2 / FFMA2 / FADD2): alpha 0 and 1 reproduce orig / up exactly, so no per-pixel
//            branches.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (which would change the rounding),
//            so the sum is written as fma(p1, one, p2) with `one` = 1.0f from a kernel parameter: two rounded
//            products, one rounded sum - exactly np.float32 arithmetic.
//   VX2      vertical axis also exact x2 (closed form); otherwise table-driven vertical taps on the closed-form
//            horizontal pass (h >> 4 == 32 * (far + 3 * near), so (b * (h >> 4)) >> 16 == (b * m) >> 11).
constexpr int K3F_THREADS = 512;
constexpr int K3F_QCAP = 128 + 32;      // one classified row of the warp (32 lanes x 4 quads) + carried-over items

// Geometry the host resolves once per call (kernel-parameter constant bank: no per-CTA index arithmetic).
struct FastGeom {
    int h, w, H0, W0, th, rpt;
    int Wpc, row_words, rows_s;          // mask words per frame row; per shared-memory bit row (+2 pad words); bit rows
    int bits_off, zbits_off, lut_off, taps_off, queue_off;    // word offsets inside dynamic shared memory (after the strip); zbits: k3_fastw only
    int G, n_tasks, n_steps;
    uint32_t inv_rw;                     // floor(2^32 / row_words) + 1 (row_words >= 3)
    uint32_t inv_wpc;                    // floor(2^32 / Wpc) + 1: id / Wpc == umulhi(id, inv_wpc) for id < 65536, Wpc > 1 (k3_fastw)
    long long frame_bytes, mask_frame_bytes, inp_frame_bytes, bits_frame_words;
    float div, one;
    float alpha[16];                     // k3_fastw: alpha of LUT level class | inside << 3 (host_alpha_levels)
    uint32_t alpha_pos;                  // bit i: level i has alpha > 0
};

template <bool VX2, bool BITS, int NTH, int HR = 2>
__global__ void __launch_bounds__(NTH, 1024 / NTH)
    k3_fast(const uint8_t *__restrict__ inp, const uint8_t *__restrict__ orig, const uint8_t *__restrict__ mask,
            const uint32_t *__restrict__ mask_bits, uint8_t *__restrict__ out, const Tap *__restrict__ yt,
            const __grid_constant__ FastGeom gm) {
    extern __shared__ __align__(128) uint32_t smem_base[];
    const int h = gm.h, w = gm.w, H0 = gm.H0, W0 = gm.W0, th = gm.th, rpt = gm.rpt;
    const int Wpc = gm.Wpc, row_words = gm.row_words, rows_s = gm.rows_s;
    const float div = gm.div, one = gm.one;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_base);                        // mbarrier first, strip 16 bytes in
    uint8_t *strip = reinterpret_cast<uint8_t *>(smem_base + 4);
    uint32_t *bits = smem_base + gm.bits_off;                                       // [rows_s][row_words]
    float4 *lut = reinterpret_cast<float4 *>(smem_base + gm.lut_off);               // {a, a, 1-a, 1-a} x 16, then lut_pos
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *queue = reinterpret_cast<uint2 *>(smem_base + gm.queue_off) + warp * K3F_QCAP;

    const long long t = blockIdx.y;                  // grid = (strips, frames): no division
    const int y0 = (int)blockIdx.x * th;
    const uint8_t *orig_t = orig + t * gm.frame_bytes;
    uint8_t *out_t = out + t * gm.frame_bytes;
    const uint32_t strip_bytes = (uint32_t)(min(th, H0 - y0) * W0 * 3);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, strip_bytes);
        const uint8_t *src = orig_t + (long long)y0 * W0 * 3;
        for (uint32_t off = 0; off < strip_bytes; off += 32768u)
            bulk_g2s(strip + off, src + off, min(32768u, strip_bytes - off), bar);
    }

    // ---------------- phase 1: bit rows [y0 - 2, y0 + th + 2), one zero pad word on each side
    if (BITS) {
        for (int i = warp; i < rows_s; i += NTH / 32) {
            const int y = y0 - 2 + i;
            const bool row_ok = y >= 0 && y < H0;
            const uint32_t *src = mask_bits + t * gm.bits_frame_words + (long long)y * Wpc;
            for (int k = lane; k < row_words; k += 32)
                bits[i * row_words + k] = (row_ok && k >= 1 && k <= Wpc) ? __ldg(src + (k - 1)) : 0u;
        }
    } else {
        uint16_t *b16 = reinterpret_cast<uint16_t *>(bits);
        const int halves = 2 * row_words;
        const uint8_t *mask_t = mask + t * gm.mask_frame_bytes;
        for (int id = threadIdx.x; id < rows_s * halves; id += NTH) {
            const int i = id / halves, hw = id - i * halves;
            const int y = y0 - 2 + i, x0 = (hw - 2) * 16;
            uint32_t v = 0;
            if (y >= 0 && y < H0 && x0 >= 0 && x0 < W0) v = nonzero_bits16(ldg128(mask_t + (long long)y * W0 + x0));
            b16[id] = (uint16_t)v;
        }
    }
    if (warp == 0) {
        // alpha levels: index = class (0 = no hit within the window, 1..5 = cost classes 1, 1.4, 2, 2.1969, 2.8) | inside << 3
        const int cls = lane & 7, inside = (lane >> 3) & 1;
        const float cost = cls == 1 ? 1.0f : cls == 2 ? 1.4f : cls == 3 ? 2.0f : cls == 4 ? 2.1969f : __fadd_rn(1.4f, 1.4f);
        float a = inside ? 1.f : 0.f;
        if (cls >= 1 && cls <= 5) a = inside ? alpha_from(cost, 0.f, div) : alpha_from(0.f, cost, div);
        const float na = __fsub_rn(1.f, a);
        if (lane < 16) lut[lane] = make_float4(a, a, na, na);
        const uint32_t posmask = __ballot_sync(0xffffffffu, lane < 16 && a > 0.f);     // which LUT levels have alpha > 0
        if (lane == 0) reinterpret_cast<uint32_t *>(lut + 16)[0] = posmask;
    }
    __syncthreads();

    const uint32_t lut_pos = reinterpret_cast<const uint32_t *>(lut + 16)[0];
    // outside classes 1..5 that still blend (alpha > 0), as masks applied to p1..p5
    const uint32_t e1 = (lut_pos >> 1) & 1u ? ~0u : 0u, e2 = (lut_pos >> 2) & 1u ? ~0u : 0u, e3 = (lut_pos >> 3) & 1u ? ~0u : 0u,
                   e4 = (lut_pos >> 4) & 1u ? ~0u : 0u, e5 = (lut_pos >> 5) & 1u ? ~0u : 0u;

    const uint8_t *inp_t = inp + t * gm.inp_frame_bytes;
    const f32x2 one2 = pack2(one, one), magic2 = pack2(12582912.f, 12582912.f), unbias2 = pack2(-8388608.f, -8388608.f);

    // ---- phase 3 worker: one 4-pixel quad per lane
    auto work = [&](const uint2 item) {
        const int xq = item.x & 0xffff, r = (int)(item.x >> 16);
        const int yy = y0 + r;
        uint32_t a0, a1, a2, b0, b1, b2, wa = 0, wb = 0;
        if (VX2) {
            const int j = yy >> 1;                                           // source row of weight 3/4
            const int ja = (yy & 1) ? min(j + 1, h - 1) : max(j - 1, 0);     // source row of weight 1/4
            x2_load_row(inp_t + ja * w * 3, xq, W0, a0, a1, a2);
            x2_load_row(inp_t + j * w * 3, xq, W0, b0, b1, b2);
        } else {
            const Tap ty = yt[yy];
            wa = (uint32_t)(ty.w & 0xffff) << 20, wb = ((uint32_t)ty.w >> 16) << 20;
            const int ya = min(max(ty.ofs, 0), h - 1), yb = min(max(ty.ofs + 1, 0), h - 1);
            if (HR == 4) {
                x4_load_row(inp_t + ya * w * 3, xq, W0, a0, a1, a2);
                x4_load_row(inp_t + yb * w * 3, xq, W0, b0, b1, b2);
            } else {
                x2_load_row(inp_t + ya * w * 3, xq, W0, a0, a1, a2);
                x2_load_row(inp_t + yb * w * 3, xq, W0, b0, b1, b2);
            }
        }
        uint32_t *sp = reinterpret_cast<uint32_t *>(strip + (r * W0 + xq) * 3);
        const uint32_t o0 = sp[0], o1 = sp[1], o2 = sp[2];
        // item.y = plane nibbles of the quad: L0 @ bits 0-3, L2 @ 4-7, L1 @ 16-19, inside @ 20-23 (bit i = pixel i).
        // LUT byte offset of pixel i = 16 * (L0 | L1 << 1 | L2 << 2 | inside << 3): one multiply gathers the four
        // bits (0x140028 = 2^20 + 2^18 + 2^5 + 2^3 sends bits 0, 16, 4, 20 to bits 20..23; no two partial products
        // share a bit position, so there are no carries).
        const uint8_t *lutb = reinterpret_cast<const uint8_t *>(lut);
        auto lut_at = [&](int i) {
            const uint32_t tsel = (item.y >> i) & 0x00110011u;
            return *reinterpret_cast<const float4 *>(lutb + (((tsel * 0x140028u) >> 16) & 0xf0u));
        };
        const float4 l0 = lut_at(0), l1 = lut_at(1), l2 = lut_at(2), l3 = lut_at(3);
        uint32_t ma[6], mb[6], up[6];
        if (HR == 4) {
            x4_hpass(a0, a1, a2, ma);
            x4_hpass(b0, b1, b2, mb);
        } else {
            x2_hpass(a0, a1, a2, ma);
            x2_hpass(b0, b1, b2, mb);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (VX2) {
                up[k] = x2_vpass(ma[k], mb[k]);
            } else {
                // (b * (h >> 4)) >> 16 with h >> 4 == 32 m (x2) or 16 m (x4): umulhi(b << 20, m << 1 or m)
                const uint32_t am = HR == 4 ? ma[k] : ma[k] << 1, bm = HR == 4 ? mb[k] : mb[k] << 1;   // lanes <= 2040 / 4080: no carry
                const uint32_t lo = (__umulhi(wa, am & 0xffffu) + __umulhi(wb, bm & 0xffffu) + 2u) >> 2;
                const uint32_t hi = (__umulhi(wa, am >> 16) + __umulhi(wb, bm >> 16) + 2u) >> 2;
                up[k] = lo | (hi << 16);
            }
        }
        // byte pair p = bytes (2p, 2p+1) of the 12-byte quad; byte k belongs to pixel k / 3
        const f32x2 al[6] = {pack2(l0.x, l0.y), pack2(l0.x, l1.x), pack2(l1.x, l1.y),
                             pack2(l2.x, l2.y), pack2(l2.x, l3.x), pack2(l3.x, l3.y)};
        const f32x2 nl[6] = {pack2(l0.z, l0.w), pack2(l0.z, l1.z), pack2(l1.z, l1.w),
                             pack2(l2.z, l2.w), pack2(l2.z, l3.z), pack2(l3.z, l3.w)};
        const uint32_t ow[3] = {o0, o1, o2};
        uint32_t rb[12];
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            // u8 -> f32 through the mantissa: bits(2^23 + b) - 2^23 (both lanes at once)
            const f32x2 uf = fadd2(pack2u(__byte_perm(up[p], 0x4b000000u, 0x7540u), __byte_perm(up[p], 0x4b000000u, 0x7542u)),
                                   unbias2);
            const uint32_t wsrc = ow[p >> 1];
            const uint32_t s0 = 0x7540u | (uint32_t)((2 * p) & 3), s1 = 0x7540u | (uint32_t)((2 * p + 1) & 3);
            const f32x2 of = fadd2(pack2u(__byte_perm(wsrc, 0x4b000000u, s0), __byte_perm(wsrc, 0x4b000000u, s1)), unbias2);
            // f32(a * up) + f32((1 - a) * orig), then round half to even through the mantissa
            const f32x2 v = fadd2(ffma2(fmul2(al[p], uf), one2, fmul2(nl[p], of)), magic2);
            unpack2u(v, rb[2 * p], rb[2 * p + 1]);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            sp[j] = __byte_perm(__byte_perm(rb[4 * j], rb[4 * j + 1], 0x0040), __byte_perm(rb[4 * j + 2], rb[4 * j + 3], 0x0040),
                                0x5410);
    };

    // ---------------- phase 2: classification, rolling over `rpt` rows per thread
    // One flat, rolled loop (a step = one row of every thread's task) so that the classification and the worker
    // exist once in the instruction stream; the last step only drains the queue.
    const int G = gm.G, n_tasks = gm.n_tasks, n_steps = gm.n_steps;
    int qcount = 0, it = 0, j = 0;
    bool landed = false, task_ok = false;
    int c0 = 0, r0 = 0;
    uint32_t colvalid = 0, xbase = 0;
    uint32_t Mw[5] = {0, 0, 0, 0, 0}, Zw[5] = {0, 0, 0, 0, 0};
#pragma unroll 1
    for (int step = 0; step <= n_steps; ++step) {
        const bool drain = step == n_steps;
        uint32_t need = 0, L0 = 0, L1 = 0, L2 = 0, M2 = 0;
        int r = 0;
        if (!drain) {
            if (j == 0) {                                                    // new task: (column group, row block)
                // Each warp's 32 lanes sample the strip evenly (8 runs of 4 consecutive tasks, NTH / 8 tasks apart), so that
                // the warps of a CTA get the same amount of blend work whatever the mask looks like: they all meet
                // at the final barrier, and ncu showed barrier stalls on top when a warp owned 512 contiguous pixels.
                const int id = it * NTH + (lane >> 2) * (NTH / 8) + (warp << 2) + (lane & 3);
                task_ok = id < n_tasks;
                const int rbk = task_ok ? id / G : 0, g = task_ok ? id - rbk * G : 0;
                r0 = rbk * rpt, c0 = g * 16 - 8, xbase = (uint32_t)(g * 16);
                const int lo = max(0, -c0), hi = min(32, W0 - c0);
                colvalid = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                // window rows: strip row r - 2 .. r + 2  <->  bit row index r .. r + 4
#pragma unroll
                for (int d = 1; d < 5; ++d) {
                    const int bi = min(r0 + d - 1, rows_s - 1), yy = y0 + r0 + d - 3;
                    Mw[d] = bit_window(bits + bi * row_words, c0);
                    Zw[d] = (yy >= 0 && yy < H0) ? (~Mw[d] & colvalid) : 0u;
                }
            }
            const int rr = r0 + j;
            r = rr;
#pragma unroll
            for (int d = 0; d < 4; ++d) Mw[d] = Mw[d + 1], Zw[d] = Zw[d + 1];
            {
                const int bi = min(rr + 4, rows_s - 1), yy = y0 + rr + 2;
                Mw[4] = bit_window(bits + bi * row_words, c0);
                Zw[4] = (yy >= 0 && yy < H0) ? (~Mw[4] & colvalid) : 0u;
            }
            M2 = Mw[2];
            if (task_ok && rr < th && y0 + rr < H0) {
                const uint32_t anyM = Mw[0] | Mw[1] | Mw[2] | Mw[3] | Mw[4];
                // the 5x5 window of pixel i covers bits i+6 .. i+10, i.e. bits 6..25 for the whole group
                if (anyM & 0x03ffffc0u) {
                    auto classes = [](const uint32_t *S, uint32_t *hc) {
                        const uint32_t A1 = S[1] | S[3], A0 = S[0] | S[4];
                        hc[0] = (S[2] << 1) | (S[2] >> 1) | A1;                                  // cost 1
                        hc[1] = (A1 << 1) | (A1 >> 1);                                           // 1.4
                        hc[2] = (S[2] << 2) | (S[2] >> 2) | A0;                                  // 2
                        hc[3] = (A0 << 1) | (A0 >> 1) | (A1 << 2) | (A1 >> 2);                   // 2.1969
                        hc[4] = (A0 << 2) | (A0 >> 2);                                           // 2.8
                    };
                    uint32_t hm[5], hz[5], hsel[5];
                    classes(Mw, hm);
                    classes(Zw, hz);
#pragma unroll
                    for (int k = 0; k < 5; ++k) hsel[k] = (hz[k] & M2) | (hm[k] & ~M2);   // inside pixels look for zeros
                    const uint32_t p1 = hsel[0];
                    const uint32_t p2 = hsel[1] & ~p1;
                    const uint32_t s12 = p1 | hsel[1];
                    const uint32_t p3 = hsel[2] & ~s12;
                    const uint32_t s123 = s12 | hsel[2];
                    const uint32_t p4 = hsel[3] & ~s123;
                    const uint32_t p5 = hsel[4] & ~(s123 | hsel[3]);
                    L0 = p1 | p3 | p5, L1 = p2 | p3, L2 = p4 | p5;
                    // alpha > 0: every inside pixel, and outside pixels whose first hit has alpha > 0
                    const uint32_t pos = M2 | (p1 & e1) | (p2 & e2) | (p3 & e3) | (p4 & e4) | (p5 & e5);
                    need = (pos >> 8) & 0xffffu;
                }
            }
        }
        // ---- warp-level compaction: every quad with a pixel to blend becomes one work item
        if (__ballot_sync(0xffffffffu, need != 0)) {                         // warp-uniform
            const uint32_t nzq = (need | (need >> 1) | (need >> 2) | (need >> 3)) & 0x1111u;   // bit 4q = quad q has work
            const int qn = __popc(nzq);
            int pre = qn;                                                    // inclusive scan over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += v;
            }
            int pos = qcount + pre - qn;
            qcount += __shfl_sync(0xffffffffu, pre, 31);
            if (need) {
                const uint32_t xr = xbase | ((uint32_t)r << 16);
                // plane nibbles of the four quads (the pixels of the group are bits 8..23 = bytes 1, 2 of each plane):
                //   W = [L0 px 0..15 | L1 px 0..15], V = [L2 | inside];  E / O interleave the nibbles of the even / odd
                //   quads as bytes [L0q L2q] ... [L1q Mq], and one byte permute per quad isolates its two bytes
                const uint32_t W = __byte_perm(L0, L1, 0x6521), V = __byte_perm(L2, M2, 0x6521);
                const uint32_t E = (W & 0x0f0f0f0fu) | ((V << 4) & 0xf0f0f0f0u);
                const uint32_t O = ((W >> 4) & 0x0f0f0f0fu) | (V & 0xf0f0f0f0u);
                const uint32_t yq[4] = {__byte_perm(E, 0u, 0x4240), __byte_perm(O, 0u, 0x4240), __byte_perm(E, 0u, 0x4341),
                                        __byte_perm(O, 0u, 0x4341)};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((nzq >> (4 * q)) & 1u) queue[pos++] = make_uint2(xr + 4u * q, yq[q]);
            }
            __syncwarp();
        }
        if (qcount >= 32 || (drain && qcount > 0)) {                         // warp-uniform
            if (!landed) {
                mbar_wait(bar, 0);
                landed = true;
            }
            do {
                const int take = min(qcount, 32);
                qcount -= take;
                if (lane < take) work(queue[qcount + lane]);
            } while (qcount >= 32);
            __syncwarp();              // the queue tail is overwritten by the next pushes
        }
        if (++j == rpt) j = 0, ++it;
    }
    // Warps that patched the strip have waited for it; the thread that stores it must have, too (the others
    // never touch the strip and need not look at the mbarrier).
    fence_proxy_async();                 // the patched quads must be visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!landed) mbar_wait(bar, 0);
        uint8_t *dst = out_t + (long long)y0 * W0 * 3;
        for (uint32_t off = 0; off < strip_bytes; off += 32768u)
            bulk_s2g(dst + off, strip + off, min(32768u, strip_bytes - off));
        bulk_commit_and_wait_read();     // shared memory must outlive the reads of the store
    }
}

// =====================================================================================================
// k3_fastw: k3_fast with WORD tasks.  ncu's source view of k3_fast (profiles/r02_k3_source_breakdown.txt) put 36 %
// of the warp instructions into classification + compaction, because a lane classified 16 useful pixels out of the
// 32 bits it shifted around (8 halo bits on each side) and carried a rolling 5-row window in registers.  Here a
// task is one whole 32-bit word of one strip row: the +-1 / +-2 column shifts take their carry-in from the
// neighbour words through funnel shifts (one SHF each, like a plain shift), so all 32 bits are useful; the five
// window rows of both polarities (mask M and in-frame complement Z, both kept in shared memory) are simply
// re-loaded per task (LSU pipe, which idles) instead of being rolled through registers (ALU pipe, which is the
// bound).  A lane-step yields up to 8 quads; work items shrink to ONE word
//     bits 0..9 quad index in the row (x / 4), 10..15 row in strip, 16..31 plane nibbles [L0 | L1 | L2 | inside]
// and the worker's LUT index is again one multiply (0x1248 sends bits 0, 4, 8, 12 to bits 12..15 without carries).
// Requires W0 <= 4096.  Worker arithmetic is k3_fast's.
// Two-ended per-warp queue: quads to BLEND grow from the bottom; from the top, one entry per fully INTERIOR word
// (every pixel of the word and of its 5 x 36 window is masked: alpha == 1, the result is the up-scaled pixel itself,
// no LUT, no blend).  A lane-step adds at most 256 entries, at most 31 + 3 are carried over between steps.
constexpr int K3W_QCAP = 256 + 48;

template <bool VX2, bool BITS, int NTH, int HR = 2>
__global__ void __launch_bounds__(NTH, NTH == 384 ? 3 : 1024 / NTH)
    k3_fastw(const uint8_t *__restrict__ inp, const uint8_t *__restrict__ orig, const uint8_t *__restrict__ mask,
             const uint32_t *__restrict__ mask_bits, uint8_t *__restrict__ out, const Tap *__restrict__ yt,
             const __grid_constant__ FastGeom gm) {
    extern __shared__ __align__(128) uint32_t smem_base[];
    // Launches of this kernel never depend on each other (different frames).  A following launch that is CHAINED
    // (option k3_chain: programmatic stream serialisation) may therefore start as soon as every CTA of this one has
    // started - back-to-back launches over parts of a clip then leave no drain / ramp bubble between them.
    asm volatile("griddepcontrol.launch_dependents;");
    const int h = gm.h, w = gm.w, H0 = gm.H0, W0 = gm.W0, th = gm.th;
    const int Wpc = gm.Wpc, row_words = gm.row_words, rows_s = gm.rows_s;
    const float div = gm.div, one = gm.one;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_base);                        // mbarrier first, strip 16 bytes in
    uint8_t *strip = reinterpret_cast<uint8_t *>(smem_base + 4);
    uint32_t *bitsM = smem_base + gm.bits_off;                                      // [rows_s][row_words] mask
    uint32_t *bitsZ = smem_base + gm.zbits_off;                                     // [rows_s][row_words] in-frame & ~mask
    float4 *lut = reinterpret_cast<float4 *>(smem_base + gm.lut_off);               // {a, a, 1-a, 1-a} x 16, then lut_pos
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *queue = smem_base + gm.queue_off + warp * K3W_QCAP;

    const long long t = blockIdx.y;                  // grid = (strips, frames): no division
    const int y0 = (int)blockIdx.x * th;
    const uint8_t *orig_t = orig + t * gm.frame_bytes;
    uint8_t *out_t = out + t * gm.frame_bytes;
    const uint32_t strip_bytes = (uint32_t)(min(th, H0 - y0) * W0 * 3);
    if (threadIdx.x == 0) {              // the other threads first look at the mbarrier after the __syncthreads below
        mbar_init(bar, 1);
        mbar_arrive_expect_tx(bar, strip_bytes);
        const uint8_t *src = orig_t + (long long)y0 * W0 * 3;
        for (uint32_t off = 0; off < strip_bytes; off += 32768u)
            bulk_g2s(strip + off, src + off, min(32768u, strip_bytes - off), bar);
    }

    // ---------------- phase 1: bit rows [y0 - 2, y0 + th + 2) of both polarities, one zero pad word on each side
    const uint32_t last_valid = (W0 & 31) ? ((1u << (W0 & 31)) - 1u) : 0xffffffffu;   // valid bits of frame word Wpc - 1
    if (BITS) {
        // a warp owns rows warp, warp + NTH/32, ..., a lane the words lane, lane + 32, ...; the (up to) four loads of
        // two rows x two words are all issued before the first store waits for one (ncu: a row-by-row loop spent 9 % of
        // the kernel's stall samples on its load round trips), and nothing divides
        const uint32_t *fbits = mask_bits + t * gm.bits_frame_words;
        for (int i0 = warp; i0 < rows_s; i0 += 2 * (NTH / 32))
            for (int k0 = lane; k0 < row_words; k0 += 64) {
                uint32_t mv[2][2], vv_[2][2];
#pragma unroll
                for (int ra = 0; ra < 2; ++ra) {
                    const int i = i0 + ra * (NTH / 32), y = y0 - 2 + i;
                    const bool row_ok = i < rows_s && y >= 0 && y < H0;
                    const int rowbase = y * Wpc - 1;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const int k = k0 + 32 * kb;
                        const bool ok = row_ok && k >= 1 && k <= Wpc;
                        vv_[ra][kb] = !ok ? 0u : (k == Wpc ? last_valid : 0xffffffffu);
                        mv[ra][kb] = ok ? __ldg(fbits + (rowbase + k)) : 0u;
                    }
                }
#pragma unroll
                for (int ra = 0; ra < 2; ++ra) {
                    const int i = i0 + ra * (NTH / 32);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const int k = k0 + 32 * kb;
                        if (i < rows_s && k < row_words) {
                            const uint32_t m = mv[ra][kb] & vv_[ra][kb];
                            bitsM[i * row_words + k] = m;
                            bitsZ[i * row_words + k] = ~m & vv_[ra][kb];
                        }
                    }
                }
            }
    } else {
        uint16_t *m16 = reinterpret_cast<uint16_t *>(bitsM), *z16 = reinterpret_cast<uint16_t *>(bitsZ);
        const int halves = 2 * row_words;
        const uint8_t *mask_t = mask + t * gm.mask_frame_bytes;
        for (int id = threadIdx.x; id < rows_s * halves; id += NTH) {
            const int i = id / halves, hw = id - i * halves;
            const int y = y0 - 2 + i, x0 = (hw - 2) * 16;
            uint32_t v = 0, z = 0;
            if (y >= 0 && y < H0 && x0 >= 0 && x0 < W0) {
                v = nonzero_bits16(ldg128(mask_t + (long long)y * W0 + x0));
                z = ~v & 0xffffu;
            }
            m16[id] = (uint16_t)v;
            z16[id] = (uint16_t)z;
        }
    }
    if (warp == 0) {
        // alpha levels (index = class | inside << 3) come from the host (FastGeom): computing them here put two IEEE
        // divisions per level on warp 0 - about a thousand instructions the other 15 warps waited for at the barrier
        if (lane < 16) {
            const float a = gm.alpha[lane], na = __fsub_rn(1.f, a);
            lut[lane] = make_float4(a, a, na, na);
        }
        if (lane == 0) reinterpret_cast<uint32_t *>(lut + 16)[0] = gm.alpha_pos;
    } else if (!VX2 && warp == 1 && lane < th) {
        reinterpret_cast<Tap *>(smem_base + gm.taps_off)[lane] = yt[min(y0 + lane, H0 - 1)];
    }
    __syncthreads();

    const uint32_t lut_pos = reinterpret_cast<const uint32_t *>(lut + 16)[0];
    // outside classes 1..5 that still blend (alpha > 0), as masks applied to p1..p5
    const uint32_t e1 = (lut_pos >> 1) & 1u ? ~0u : 0u, e2 = (lut_pos >> 2) & 1u ? ~0u : 0u, e3 = (lut_pos >> 3) & 1u ? ~0u : 0u,
                   e4 = (lut_pos >> 4) & 1u ? ~0u : 0u, e5 = (lut_pos >> 5) & 1u ? ~0u : 0u;

    const uint8_t *inp_t = inp + t * gm.inp_frame_bytes;
    const uint32_t *inw = reinterpret_cast<const uint32_t *>(inp_t);               // 4-byte aligned rows (host check)
    const int wq = (w * 3) >> 2;                                                   // words per source row
    const f32x2 one2 = pack2(one, one), magic2 = pack2(12582912.f, 12582912.f), unbias2 = pack2(-8388608.f, -8388608.f);

    // ---- worker, software pipelined: `fetch` issues the loads of a quad's two source rows (raw words, nothing waits
    // on them), `blend` up-scales and blends the quad fetched one round earlier - ncu put 25 % of all stall samples on
    // the first use of these loads when a round fetched and blended the same quad.
    const Tap *taps_s = reinterpret_cast<const Tap *>(smem_base + gm.taps_off);    // the strip's vertical taps (!VX2)
    struct Fetched {
        uint32_t item, a[4], b[4];
    };
    auto fetch = [&](const uint32_t item, Fetched &f) {
        const int xq = (int)(item & 0x3ffu) << 2, r = (int)(item >> 10) & 0x3f;
        const int yy = y0 + r;
        int ra, rb;
        if (VX2) {
            rb = yy >> 1;                                                    // source row of weight 3/4
            ra = (yy & 1) ? min(rb + 1, h - 1) : max(rb - 1, 0);             // source row of weight 1/4
        } else {
            const int ofs = taps_s[r].ofs;
            ra = min(max(ofs, 0), h - 1), rb = min(max(ofs + 1, 0), h - 1);
        }
        f.item = item;
        if (HR == 4) {
            x4_fetch_row(inw, ra * wq, xq, W0, f.a);
            x4_fetch_row(inw, rb * wq, xq, W0, f.b);
        } else {
            x2_fetch_row(inw, ra * wq, xq, W0, f.a);
            x2_fetch_row(inw, rb * wq, xq, W0, f.b);
        }
    };
    // `assemble` is the first use of a round's loads (the only place that waits for them); it runs BEFORE the next
    // round's fetch re-uses the raw-word registers, `finish` after it - no register copies between rounds.
    struct Assembled {
        uint32_t item, a0, a1, a2, b0, b1, b2;
    };
    auto assemble = [&](const Fetched &f, Assembled &q) {
        const int xq = (int)(f.item & 0x3ffu) << 2;
        q.item = f.item;
        if (HR == 4) {
            x4_assemble_row(f.a, xq, W0, q.a0, q.a1, q.a2);
            x4_assemble_row(f.b, xq, W0, q.b0, q.b1, q.b2);
        } else {
            x2_assemble_row(f.a, xq, W0, q.a0, q.a1, q.a2);
            x2_assemble_row(f.b, xq, W0, q.b0, q.b1, q.b2);
        }
    };
    auto finish = [&](const Assembled &q, const bool interior) {     // `interior` is warp-uniform
        const uint32_t item = q.item;
        const int xq = (int)(item & 0x3ffu) << 2, r = (int)(item >> 10) & 0x3f;
        const uint32_t a0 = q.a0, a1 = q.a1, a2 = q.a2, b0 = q.b0, b1 = q.b1, b2 = q.b2;
        uint32_t wa = 0, wb = 0;
        if (!VX2) {
            const int tw = taps_s[r].w;
            wa = (uint32_t)(tw & 0xffff) << 20, wb = ((uint32_t)tw >> 16) << 20;
        }
        uint32_t *sp = reinterpret_cast<uint32_t *>(strip + (r * W0 + xq) * 3);
        uint32_t ma[6], mb[6], up[6];
        if (HR == 4) {
            x4_hpass(a0, a1, a2, ma);
            x4_hpass(b0, b1, b2, mb);
        } else {
            x2_hpass(a0, a1, a2, ma);
            x2_hpass(b0, b1, b2, mb);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (VX2) {
                up[k] = x2_vpass_b13(ma[k], mb[k]);                 // results in bytes 1 and 3
            } else {
                // (b * (h >> 4)) >> 16 with h >> 4 == 32 m (x2) or 16 m (x4): umulhi(b << 20, m << 1 or m)
                const uint32_t am = HR == 4 ? ma[k] : ma[k] << 1, bm = HR == 4 ? mb[k] : mb[k] << 1;   // lanes <= 2040 / 4080: no carry
                const uint32_t lo = (__umulhi(wa, am & 0xffffu) + __umulhi(wb, bm & 0xffffu) + 2u) >> 2;
                const uint32_t hi = (__umulhi(wa, am >> 16) + __umulhi(wb, bm >> 16) + 2u) >> 2;
                up[k] = lo | (hi << 16);
            }
        }
        if (interior) {                  // alpha == 1 on all four pixels: the 12 up-scaled bytes as they are
#pragma unroll
            for (int j = 0; j < 3; ++j) sp[j] = __byte_perm(up[2 * j], up[2 * j + 1], VX2 ? 0x7531u : 0x6420u);
            return;
        }
        const uint32_t o0 = sp[0], o1 = sp[1], o2 = sp[2];
        // plane nibbles of the quad in the item: L0 @ bits 16-19, L1 @ 20-23, L2 @ 24-27, inside @ 28-31 (bit i = pixel i).
        // LUT byte offset of pixel i = 16 * (L0 | L1 << 1 | L2 << 2 | inside << 3): the multiplier 0x1248 >> i =
        // 2^(12-i) + 2^(9-i) + 2^(6-i) + 2^(3-i) sends bits 16+i, 20+i, 24+i, 28+i to bits 28..31; no two partial
        // products share a bit position, so there are no carries (products above bit 31 drop out).
        const uint8_t *lutb = reinterpret_cast<const uint8_t *>(lut);
        auto lut_at = [&](int i) {
            const uint32_t tsel = item & (0x11110000u << i);
            return *reinterpret_cast<const float4 *>(lutb + (((tsel * (0x1248u >> i)) >> 24) & 0xf0u));
        };
        const float4 l0 = lut_at(0), l1 = lut_at(1), l2 = lut_at(2), l3 = lut_at(3);
        // byte pair p = bytes (2p, 2p+1) of the 12-byte quad; byte k belongs to pixel k / 3
        const f32x2 al[6] = {pack2(l0.x, l0.y), pack2(l0.x, l1.x), pack2(l1.x, l1.y),
                             pack2(l2.x, l2.y), pack2(l2.x, l3.x), pack2(l3.x, l3.y)};
        const f32x2 nl[6] = {pack2(l0.z, l0.w), pack2(l0.z, l1.z), pack2(l1.z, l1.w),
                             pack2(l2.z, l2.w), pack2(l2.z, l3.z), pack2(l3.z, l3.w)};
        const uint32_t ow[3] = {o0, o1, o2};
        uint32_t rb[12];
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            // u8 -> f32 through the mantissa: bits(2^23 + b) - 2^23 (both lanes at once)
            const f32x2 uf = fadd2(pack2u(__byte_perm(up[p], 0x4b000000u, VX2 ? 0x7541u : 0x7540u),
                                          __byte_perm(up[p], 0x4b000000u, VX2 ? 0x7543u : 0x7542u)), unbias2);
            const uint32_t wsrc = ow[p >> 1];
            const uint32_t s0 = 0x7540u | (uint32_t)((2 * p) & 3), s1 = 0x7540u | (uint32_t)((2 * p + 1) & 3);
            const f32x2 of = fadd2(pack2u(__byte_perm(wsrc, 0x4b000000u, s0), __byte_perm(wsrc, 0x4b000000u, s1)), unbias2);
            // f32(a * up) + f32((1 - a) * orig), then round half to even through the mantissa
            const f32x2 v = fadd2(ffma2(fmul2(al[p], uf), one2, fmul2(nl[p], of)), magic2);
            unpack2u(v, rb[2 * p], rb[2 * p + 1]);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            sp[j] = __byte_perm(__byte_perm(rb[4 * j], rb[4 * j + 1], 0x0040), __byte_perm(rb[4 * j + 2], rb[4 * j + 3], 0x0040),
                                0x5410);
    };

    // ---------------- phase 2: classification, one 32-pixel word of one strip row per lane and step
    const int n_tasks = gm.n_tasks, n_steps = gm.n_steps;
    int qcount = 0, qwords = 0;          // quads to blend (bottom of the queue) / interior words (top)
    bool landed = false, pending = false, pending_interior = false;
    Fetched cur = {};
#pragma unroll 1
    for (int step = 0; step <= n_steps; ++step) {
        const bool drain = step == n_steps;
        uint32_t need = 0, L0 = 0, L1 = 0, L2 = 0, M2 = 0, base = 0;
        bool intr = false;               // this lane's word is fully interior
        // Each warp's 32 lanes sample the strip evenly (8 runs of 4 consecutive tasks, NTH / 8 tasks apart), so that
        // the warps of a CTA get the same amount of blend work whatever the mask looks like.
        const int id = step * NTH + (lane >> 2) * (NTH / 8) + (warp << 2) + (lane & 3);
        if (!drain && id < n_tasks) {
            const int rr = Wpc == 1 ? id : (int)__umulhi((uint32_t)id, gm.inv_wpc), k = id - rr * Wpc;   // id / Wpc (id < 65536)
            if (y0 + rr < H0) {
                // bit row rr + d <-> frame row y0 + rr - 2 + d; word index k + 1 in the padded row, pm points at word k - 1
                const uint32_t *pm = bitsM + rr * row_words + k;
                uint32_t Sp[5], Sc[5], Sn[5];
#pragma unroll
                for (int d = 0; d < 5; ++d) Sp[d] = pm[d * row_words], Sc[d] = pm[d * row_words + 1], Sn[d] = pm[d * row_words + 2];
                M2 = Sc[2];
                uint32_t A1p = Sp[1] | Sp[3], A1c = Sc[1] | Sc[3], A1n = Sn[1] | Sn[3];
                uint32_t A0p = Sp[0] | Sp[4], A0c = Sc[0] | Sc[4], A0n = Sn[0] | Sn[4];
                // any mask pixel in the 5 x 36 window of this word?
                const uint32_t anyM = (M2 | A1c | A0c) | ((Sp[2] | A1p | A0p) >> 30) | ((Sn[2] | A1n | A0n) << 30);
                if (anyM) {
                    auto classes = [](uint32_t s2p, uint32_t s2c, uint32_t s2n, uint32_t a1p, uint32_t a1c, uint32_t a1n,
                                      uint32_t a0p, uint32_t a0c, uint32_t a0n, uint32_t *hc) {
                        // column shifts with the carry-in of the neighbour word: x-1 / x-2 from the previous word's top
                        // bits (funnel shift left), x+1 / x+2 from the next word's low bits (funnel shift right)
                        hc[0] = __funnelshift_l(s2p, s2c, 1) | __funnelshift_r(s2c, s2n, 1) | a1c;              // cost 1
                        hc[1] = __funnelshift_l(a1p, a1c, 1) | __funnelshift_r(a1c, a1n, 1);                    // 1.4
                        hc[2] = __funnelshift_l(s2p, s2c, 2) | __funnelshift_r(s2c, s2n, 2) | a0c;              // 2
                        hc[3] = __funnelshift_l(a0p, a0c, 1) | __funnelshift_r(a0c, a0n, 1) |
                                __funnelshift_l(a1p, a1c, 2) | __funnelshift_r(a1c, a1n, 2);                    // 2.1969
                        hc[4] = __funnelshift_l(a0p, a0c, 2) | __funnelshift_r(a0c, a0n, 2);                    // 2.8
                    };
                    uint32_t hm[5], hz[5];
                    classes(Sp[2], Sc[2], Sn[2], A1p, A1c, A1n, A0p, A0c, A0n, hm);
                    const uint32_t *pz = bitsZ + rr * row_words + k;
#pragma unroll
                    for (int d = 0; d < 5; ++d) Sp[d] = pz[d * row_words], Sc[d] = pz[d * row_words + 1], Sn[d] = pz[d * row_words + 2];
                    A1p = Sp[1] | Sp[3], A1c = Sc[1] | Sc[3], A1n = Sn[1] | Sn[3];
                    A0p = Sp[0] | Sp[4], A0c = Sc[0] | Sc[4], A0n = Sn[0] | Sn[4];
                    // no unmasked frame pixel in the 5 x 36 window: every pixel of the word has alpha == 1
                    intr = ((Sc[2] | A1c | A0c) | ((Sp[2] | A1p | A0p) >> 30) | ((Sn[2] | A1n | A0n) << 30)) == 0u;
                    need = M2;           // interior word: all of its (frame) pixels, straight from the up-scaled image
                    if (!intr) {
                        classes(Sp[2], Sc[2], Sn[2], A1p, A1c, A1n, A0p, A0c, A0n, hz);
                        uint32_t hsel[5];
#pragma unroll
                        for (int c = 0; c < 5; ++c) hsel[c] = (hz[c] & M2) | (hm[c] & ~M2);   // inside pixels look for zeros
                        const uint32_t p1 = hsel[0];
                        const uint32_t p2 = hsel[1] & ~p1;
                        const uint32_t s12 = p1 | hsel[1];
                        const uint32_t p3 = hsel[2] & ~s12;
                        const uint32_t s123 = s12 | hsel[2];
                        const uint32_t p4 = hsel[3] & ~s123;
                        const uint32_t p5 = hsel[4] & ~(s123 | hsel[3]);
                        L0 = p1 | p3 | p5, L1 = p2 | p3, L2 = p4 | p5;
                        // alpha > 0: every inside pixel, and outside pixels whose first hit has alpha > 0; only pixels of the frame
                        const uint32_t pos = M2 | (p1 & e1) | (p2 & e2) | (p3 & e3) | (p4 & e4) | (p5 & e5);
                        need = pos & (k == Wpc - 1 ? last_valid : 0xffffffffu);
                    }
                    base = (uint32_t)(k << 3) | ((uint32_t)rr << 10);
                }
            }
        }
        // ---- warp-level compaction: every quad with a pixel to blend becomes one work item, every interior word one entry
        const uint32_t bal_b = __ballot_sync(0xffffffffu, need != 0 && !intr), bal_i = __ballot_sync(0xffffffffu, need != 0 && intr);
        const uint32_t nzq = (need | (need >> 1) | (need >> 2) | (need >> 3)) & 0x11111111u;       // bit 4q = quad q has work
        if (bal_b) {                                                         // warp-uniform
            const int qn = intr ? 0 : __popc(nzq);
            int pre = qn;                                                    // inclusive scan over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += v;
            }
            uint32_t *qp = queue + (qcount + pre - qn);
            qcount += __shfl_sync(0xffffffffu, pre, 31);
            if (qn) {
                // byte j of P01e / P23e = [L1 : L0] / [inside : L2] nibbles of quad 2j, of P01o / P23o those of quad 2j + 1
                const uint32_t P01e = (L0 & 0x0f0f0f0fu) | ((L1 << 4) & 0xf0f0f0f0u), P01o = ((L0 >> 4) & 0x0f0f0f0fu) | (L1 & 0xf0f0f0f0u);
                const uint32_t P23e = (L2 & 0x0f0f0f0fu) | ((M2 << 4) & 0xf0f0f0f0u), P23o = ((L2 >> 4) & 0x0f0f0f0fu) | (M2 & 0xf0f0f0f0u);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = q >> 1;
                    const uint32_t sel = (uint32_t)(((4 + j) << 12) | (j << 8) | ((4 + j) << 4) | j);   // bytes 2, 3 = P01.j, P23.j
                    const uint32_t nib = __byte_perm((q & 1) ? P01o : P01e, (q & 1) ? P23o : P23e, sel);
                    if ((nzq >> (4 * q)) & 1u) *qp++ = (nib & 0xffff0000u) | (base + (uint32_t)q);
                }
            }
            __syncwarp();
        }
        if (bal_i) {                     // entry = position of the word's first quad | number of its quads << 16 (frame pixels are
                                         // a prefix of the word), pushed downwards from the top of the queue
            if (need != 0 && intr)
                queue[K3W_QCAP - 1 - (qwords + __popc(bal_i & ((1u << lane) - 1u)))] = base | ((uint32_t)__popc(nzq) << 16);
            qwords += __popc(bal_i);
            __syncwarp();
        }
        if (qcount >= 32 || (drain && qcount > 0)) {                         // warp-uniform
            if (!landed) {
                mbar_wait(bar, 0);
                landed = true;
            }
            do {
                const int take = min(qcount, 32);
                qcount -= take;
                const bool has = lane < take;
                Assembled q;
                if (pending) assemble(cur, q);    // the round fetched before: its loads had a whole round (or step) to land
                if (has) fetch(queue[qcount + lane], cur);
                if (pending) finish(q, pending_interior);
                pending = has, pending_interior = false;
            } while (qcount >= 32);
            __syncwarp();              // the queue tail is overwritten by the next pushes
        }
        if (qwords >= 4 || (drain && qwords > 0)) {                          // warp-uniform: four words = 32 quads per round
            if (!landed) {
                mbar_wait(bar, 0);
                landed = true;
            }
            do {
                const int take = min(qwords, 4);
                qwords -= take;
                uint32_t e = 0;
                if ((lane >> 3) < take) e = queue[K3W_QCAP - 1 - (qwords + (lane >> 3))];
                const bool has = (uint32_t)(lane & 7) < (e >> 16);
                Assembled q;
                if (pending) assemble(cur, q);
                if (has) fetch((e & 0xffffu) + (uint32_t)(lane & 7), cur);
                if (pending) finish(q, pending_interior);
                pending = has, pending_interior = true;
            } while (qwords >= 4);
            __syncwarp();
        }
    }
    if (pending) {
        Assembled q;
        assemble(cur, q);
        finish(q, pending_interior);
    }
    fence_proxy_async();                 // the patched quads must be visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!landed) mbar_wait(bar, 0);
        uint8_t *dst = out_t + (long long)y0 * W0 * 3;
        for (uint32_t off = 0; off < strip_bytes; off += 32768u)
            bulk_s2g(dst + off, strip + off, min(32768u, strip_bytes - off));
        bulk_commit_and_wait_read();     // shared memory must outlive the reads of the store
    }
}

// =====================================================================================================
// k3_bigfeather: feather_px in (8, 32] (window radius 8..31; the GUI always passes 3, the reference accepts any float).
// Correctness first, one simple pass per strip of 8 rows:
//   phase 1  mask strip + R halo rows -> bit rows in shared memory (K1's 1-bit plane when the caller has it)
//   phase 2  per 16-pixel group, bit-parallel: walk the cost-sorted offset table (device memory; up to ~3 100 offsets with
//            two-pass chamfer cost < feather_px) until every undecided pixel of the group has met its first opposite
//            pixel = its distance; groups without any opposite pixel in reach are pruned first.  The entry index of a
//            pixel's first hit goes to a u16 plane in shared memory
//   phase 3  per 4-pixel quad: alpha from the entry's cost, generic tap-table up-scale, non-FMA fp32 blend
// The table is the single-zero-pixel result of the two raster passes (build_chamfer_table): NOT symmetric, the entry of
// offset (dx, dy) is what cv2.distanceTransform yields at a pixel whose only zero pixel sits at (x + dx, y + dy).
struct BigEntry {
    float cost;
    short dx, dy;
};
constexpr int K3_BIG_MAX_ENTRIES = 4096;
constexpr int K3_BIG_TH = 8;
constexpr int K3_BIG_THREADS = 256;

__global__ void __launch_bounds__(K3_BIG_THREADS)
    k3_bigfeather(const uint8_t *__restrict__ inp, const uint8_t *__restrict__ orig, const uint8_t *__restrict__ mask,
                  const uint32_t *__restrict__ mask_bits, uint8_t *__restrict__ out, const Tap *__restrict__ xt,
                  const Tap *__restrict__ yt, const BigEntry *__restrict__ tab, int n_entries, int R, float div, int h, int w,
                  int H0, int W0, int words_ok, int mask_vec) {
    extern __shared__ __align__(16) uint32_t big_smem[];
    const int Wp = (W0 + 31) >> 5, row_words = Wp + 3;            // one zero pad word on the left, two on the right
    const int rows_s = K3_BIG_TH + 2 * R;
    uint32_t *bits = big_smem;                                    // [rows_s][row_words], bit p of a row = column p - 32
    uint16_t *cls = reinterpret_cast<uint16_t *>(big_smem + rows_s * row_words);   // [K3_BIG_TH][W0]: first-hit entry + 1
    const long long t = blockIdx.y;
    const int y0 = (int)blockIdx.x * K3_BIG_TH;
    const uint8_t *mask_t = mask + t * H0 * (long long)W0;
    const uint8_t *orig_t = orig + t * H0 * (long long)W0 * 3;
    const uint8_t *inp_t = inp + t * h * (long long)w * 3;
    uint8_t *out_t = out + t * H0 * (long long)W0 * 3;

    // ---------------- phase 1
    for (int id = threadIdx.x; id < rows_s * row_words; id += K3_BIG_THREADS) {
        const int i = id / row_words, k = id - i * row_words - 1, y = y0 - R + i;
        uint32_t v = 0;
        if (y >= 0 && y < H0 && k >= 0 && k < Wp) {
            if (mask_bits != nullptr) {
                v = __ldg(mask_bits + (t * H0 + y) * Wp + k);
            } else {
                const uint8_t *p = mask_t + (long long)y * W0 + k * 32;
                const int n = min(32, W0 - k * 32);
                if (mask_vec && n == 32) {
                    v = nonzero_bits16(ldg128(p)) | (nonzero_bits16(ldg128(p + 16)) << 16);
                } else {
                    for (int q = 0; q < n; ++q) v |= (uint32_t)(p[q] != 0) << q;
                }
            }
            if (k == Wp - 1 && (W0 & 31)) v &= (1u << (W0 & 31)) - 1u;
        }
        bits[id] = v;
    }
    for (int id = threadIdx.x; id < (K3_BIG_TH * W0 + 1) / 2; id += K3_BIG_THREADS) reinterpret_cast<uint32_t *>(cls)[id] = 0;
    __syncthreads();

    // 16 mask bits of bit row `row` starting at frame column `col` (-32 <= col <= W0 + 31)
    auto window16 = [&](const uint32_t *row, int col) -> uint32_t {
        const int p = col + 32;
        return __funnelshift_r(row[p >> 5], row[(p >> 5) + 1], p & 31) & 0xffffu;
    };
    // bits i of a 16-pixel window at column `col` that lie inside the frame
    auto valid16 = [&](int col) -> uint32_t {
        const int lo = max(0, -col), hi = min(16, W0 - col);
        return hi <= lo ? 0u : (((1u << hi) - 1u) & ~((1u << lo) - 1u));
    };

    // ---------------- phase 2
    const int G = (W0 + 15) >> 4;
    for (int task = threadIdx.x; task < K3_BIG_TH * G; task += K3_BIG_THREADS) {
        const int r = task / G, x0 = (task - r * G) * 16, y = y0 + r;
        if (y >= H0) continue;
        const uint32_t *brow = bits + (r + R) * row_words;
        const uint32_t px = valid16(x0);
        const uint32_t inside = window16(brow, x0) & px;
        // any masked / any unmasked frame pixel within reach of the group (columns x0 - R .. x0 + 15 + R, rows y - R .. y + R)?
        const int c_lo = max(0, x0 - R), c_hi = min(W0 - 1, x0 + 15 + R);
        const int w_lo = (c_lo + 32) >> 5, w_hi = (c_hi + 32) >> 5;
        uint32_t anyM = 0, anyZ = 0;
        for (int d = -R; d <= R && !(anyM && anyZ); ++d) {
            const int yy = y + d;
            if (yy < 0 || yy >= H0) continue;
            const uint32_t *row = brow + d * row_words;
            for (int wi = w_lo; wi <= w_hi; ++wi) {
                uint32_t rng = 0xffffffffu;
                if (wi == w_lo) rng &= 0xffffffffu << ((c_lo + 32) & 31);
                if (wi == w_hi) rng &= 0xffffffffu >> (31 - ((c_hi + 32) & 31));
                const uint32_t m = row[wi];
                anyM |= m & rng, anyZ |= ~m & rng;
            }
        }
        uint32_t und = ((anyZ ? inside : 0u) | (anyM ? ~inside : 0u)) & px;
        // four entries per trip: their table words and bit-row words are requested together (the walk is bound by the
        // latency of these dependent loads, not by instructions); hits are still taken in table order
        for (int e0 = 0; e0 < n_entries && und; e0 += 4) {
            BigEntry en[4];
            uint32_t m[4], v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) en[u] = tab[min(e0 + u, n_entries - 1)];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int yy = y + en[u].dy, col = x0 + en[u].dx;
                // out-of-image pixels are "far" for both transforms; the bit rows of such rows exist in shared memory (zeros)
                const bool ok = e0 + u < n_entries && yy >= 0 && yy < H0;
                v[u] = ok ? valid16(col) : 0u;
                m[u] = window16(brow + en[u].dy * row_words, col) & v[u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t hit = (((~m[u] & v[u]) & inside) | (m[u] & ~inside)) & und;   // inside pixels look for unmasked ones
                if (hit) {
                    uint32_t hh = hit;
                    while (hh) {
                        const int i = __ffs(hh) - 1;
                        hh &= hh - 1;
                        cls[r * W0 + x0 + i] = (uint16_t)(e0 + u + 1);
                    }
                    und &= ~hit;
                }
            }
        }
    }
    __syncthreads();

    // ---------------- phase 3
    const int Q = (W0 + 3) >> 2;
    const bool inp_aligned4 = ((w * 3) % 4 == 0) && ((uintptr_t)inp_t % 4 == 0);
    for (int task = threadIdx.x; task < K3_BIG_TH * Q; task += K3_BIG_THREADS) {
        const int r = task / Q, xq = (task - r * Q) * 4, y = y0 + r;
        if (y >= H0) continue;
        const int n = min(4, W0 - xq);
        const long long po = ((long long)y * W0 + xq) * 3;
        uint32_t o[3] = {0, 0, 0};
        const bool whole = words_ok && n == 4;                    // 12 bytes, 4-byte aligned
        if (whole) {
            const uint32_t *op = reinterpret_cast<const uint32_t *>(orig_t + po);
            o[0] = __ldg(op), o[1] = __ldg(op + 1), o[2] = __ldg(op + 2);
        } else {
            for (int k = 0; k < 3 * n; ++k) o[k >> 2] |= (uint32_t)orig_t[po + k] << (8 * (k & 3));
        }
        const uint32_t in4 = window16(bits + (r + R) * row_words, xq) & 15u;
        const Tap ty = yt[y];
        const uint32_t b0s = (uint32_t)(ty.w & 0xffff) << 16, b1s = (uint32_t)ty.w & 0xffff0000u;
        const int ya = min(max(ty.ofs, 0), h - 1), yb = min(max(ty.ofs + 1, 0), h - 1);
        const uint8_t *r0 = inp_t + (long long)ya * w * 3, *r1 = inp_t + (long long)yb * w * 3;
        for (int i = 0; i < n; ++i) {
            const bool inside = (in4 >> i) & 1u;
            const uint32_t c = cls[r * W0 + xq + i];
            float a = inside ? 1.f : 0.f;
            if (c != 0) {
                const float cost = tab[c - 1].cost;
                a = inside ? alpha_from(cost, 0.f, div) : alpha_from(0.f, cost, div);
            }
            if (a > 0.f) {
                const Tap tx = xt[xq + i];
                const PixelPair p0 = load_pixel_pair(r0, tx.ofs, w, inp_aligned4), p1 = load_pixel_pair(r1, tx.ofs, w, inp_aligned4);
                const uint32_t wts = (uint32_t)tx.w;
                uint32_t cr = vpass(b0s, b1s, hpass<0>(p0, wts), hpass<0>(p1, wts));
                uint32_t cg = vpass(b0s, b1s, hpass<1>(p0, wts), hpass<1>(p1, wts));
                uint32_t cb = vpass(b0s, b1s, hpass<2>(p0, wts), hpass<2>(p1, wts));
                if (a < 1.f) {
                    const float na = __fsub_rn(1.f, a);
                    cr = blend_u8(a, na, cr, byte_of(o[(3 * i) >> 2], (3 * i) & 3));
                    cg = blend_u8(a, na, cg, byte_of(o[(3 * i + 1) >> 2], (3 * i + 1) & 3));
                    cb = blend_u8(a, na, cb, byte_of(o[(3 * i + 2) >> 2], (3 * i + 2) & 3));
                }
                const uint32_t rgb[3] = {cr, cg, cb};
                for (int ch = 0; ch < 3; ++ch) {
                    const int k = 3 * i + ch;
                    o[k >> 2] = (o[k >> 2] & ~(0xffu << (8 * (k & 3)))) | ((rgb[ch] & 0xffu) << (8 * (k & 3)));
                }
            }
        }
        if (whole) {
            uint32_t *dp = reinterpret_cast<uint32_t *>(out_t + po);
            dp[0] = o[0], dp[1] = o[1], dp[2] = o[2];
        } else {
            for (int k = 0; k < 3 * n; ++k) out_t[po + k] = (uint8_t)byte_of(o[k >> 2], k & 3);
        }
    }
}

// Host: the 16 alpha levels of the small-radius kernels, index = class (0 = no hit within the window, 1..5 = cost
// classes 1, 1.4, 2, 2.1969, 2.8) | inside << 3; the same IEEE float32 operations as alpha_from() on the device
// (diffuerase.py:99-100: 0.5 + (d_in - d_out) / (2 F), clipped).
static void host_alpha_levels(float div, float alpha[16], uint32_t *pos) {
    volatile float b14 = 1.4f;
    volatile float c28 = b14 + b14;
    const float cost[6] = {0.f, 1.0f, 1.4f, 2.0f, 2.1969f, c28};
    *pos = 0;
    for (int i = 0; i < 16; ++i) {
        const int cls = i & 7, inside = i >> 3;
        float a = inside ? 1.f : 0.f;
        if (cls >= 1 && cls <= 5) {
            volatile float diff = inside ? cost[cls] - 0.f : 0.f - cost[cls];
            volatile float q = diff / div;
            volatile float v = 0.5f + q;
            a = fminf(fmaxf(v, 0.f), 1.f);
        }
        alpha[i] = a;
        if (a > 0.f) *pos |= 1u << i;
    }
}

// Host: the chamfer table of ONE zero pixel = what cv2.distanceTransform(DIST_L2, 5) yields around it: the two raster
// passes (forward: neighbours (-1, -2..2), (-2, +-1), (0, -1); backward: their mirror images), every step one float32
// addition.  With several zero pixels the transform is the minimum of the shifted tables (float32 addition is monotone).
// Float addition is not associative, so the table is not symmetric from d ~ 12 on; same construction as
// oracle/prepost.py chamfer_cost_table, verified bit-exact against cv2 4.13 / IPP for feather_px <= 32.
// at(oy, ox) of the result = value at offset (oy, ox) FROM the zero pixel, |oy|, |ox| <= R.
static std::vector<float> build_chamfer_table(int R) {
    const float A = 1.0f, B = 1.4f, Cc = 2.1969f;
    const int n = 2 * R + 1 + 8, c = n / 2, S = n + 4;
    std::vector<float> d((size_t)S * S, INFINITY);
    auto at = [&](int i, int j) -> float & { return d[(size_t)i * S + j]; };
    at(c + 2, c + 2) = 0.f;
    const int udi[7] = {-1, -1, -1, -1, -1, -2, -2}, udj[7] = {-2, -1, 0, 1, 2, -1, 1};
    const float uc[7] = {Cc, B, A, B, Cc, Cc, Cc};
    for (int pass = 0; pass < 2; ++pass) {
        const int sgn = pass ? -1 : 1;
        for (int ii = 0; ii < n; ++ii) {
            const int i = pass ? n + 1 - ii : 2 + ii;
            for (int jj = 0; jj < n; ++jj) {
                const int j = pass ? n + 1 - jj : 2 + jj;
                float v = at(i, j);
                for (int k = 0; k < 7; ++k) {
                    volatile float t = at(i + sgn * udi[k], j + sgn * udj[k]) + uc[k];   // force a rounded fp32 sum
                    if (t < v) v = t;
                }
                volatile float t = at(i, j - sgn) + A;
                if (t < v) v = t;
                at(i, j) = v;
            }
        }
    }
    std::vector<float> tab((size_t)(2 * R + 1) * (2 * R + 1));
    for (int oy = -R; oy <= R; ++oy)
        for (int ox = -R; ox <= R; ++ox) tab[(size_t)(oy + R) * (2 * R + 1) + (ox + R)] = at(c + 2 + oy, c + 2 + ox);
    return tab;
}

// Offsets (dy, dx) at which a pixel p may find its nearest opposite pixel (at p + (dx, dy)), cost < feather_px, sorted by
// cost: the offset of p FROM that pixel is -(dx, dy).
static std::vector<std::tuple<float, int, int>> chamfer_entries(float feather_px, int R) {
    const std::vector<float> tab = build_chamfer_table(R);
    std::vector<std::tuple<float, int, int>> ent;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
            if (!dx && !dy) continue;
            const float c = tab[(size_t)(-dy + R) * (2 * R + 1) + (-dx + R)];
            if (c < feather_px) ent.push_back(std::make_tuple(c, dy, dx));
        }
    std::sort(ent.begin(), ent.end());
    return ent;
}

static void build_feather_table(float feather_px, FeatherTable *ft) {
    ft->div = (float)(2.0 * (double)feather_px);
    ft->n = 0;
    ft->n_cls = 0;
    ft->radius = 0;
    if (!(feather_px > 0.f)) return;
    const int R = (int)ceilf(feather_px) - 1;
    ft->radius = R < 0 ? 0 : R;
    if (R <= 0) return;       // every neighbour is >= 1 >= F away: alpha is the hard mask
    if (R > 7) return;        // k3_bigfeather walks its own table in device memory
    const std::vector<std::tuple<float, int, int>> ent = chamfer_entries(feather_px, R);
    ft->n_cls = 0;
    for (size_t i = 0; i < ent.size() && i < (size_t)K3_MAX_ENTRIES; ++i) {
        ft->cost[i] = std::get<0>(ent[i]);
        ft->dy[i] = (int8_t)std::get<1>(ent[i]);
        ft->dx[i] = (int8_t)std::get<2>(ent[i]);
        if (i == 0 || ft->cost[i] != ft->cost[i - 1]) {
            if (ft->n_cls == K3_MAX_CLASSES) break;          // cannot happen for feather_px <= 8 (radius <= 7)
            ft->ccost[++ft->n_cls] = ft->cost[i];
        }
        ft->cls[i] = (uint8_t)ft->n_cls;
        ft->n = (int)i + 1;
    }
}

}  // namespace vv

using namespace vv;

extern "C" int vv_chamfer_table(int R, float *table) {
    VV_CHECK_ARG(table && R >= 0 && R <= 31, "vv_chamfer_table: 0 <= R <= 31 and a table of (2R + 1)^2 floats required");
    const std::vector<float> t = build_chamfer_table(R);
    std::copy(t.begin(), t.end(), table);
    return VV_OK;
}

// tap tables of the resize-back + room for the offset table of k3_bigfeather (feather_px > 8)
extern "C" size_t vv_composite_workspace_bytes(int H0, int W0) {
    const size_t taps = vv_resize_workspace_bytes(H0, W0);
    return taps ? align_up(taps, 256) + K3_BIG_MAX_ENTRIES * sizeof(BigEntry) : 0;
}

static int composite_impl(const uint8_t *inp, int T, int h, int w, const uint8_t *orig, const uint8_t *mask,
                          const uint32_t *mask_bits, int H0, int W0, float feather_px, int keep_unmasked, uint8_t *out,
                          void *workspace, size_t workspace_bytes, void *stream) {
    VV_CHECK_ARG(inp && out && workspace, "vv_upscale_feather_composite: NULL pointer");
    VV_CHECK_ARG(T > 0 && h > 0 && w > 0 && H0 > 0 && W0 > 0, "vv_upscale_feather_composite: bad shape");
    VV_CHECK_ARG(workspace_bytes >= vv_composite_workspace_bytes(H0, W0),
                 "vv_upscale_feather_composite: workspace too small");
    if (!keep_unmasked)   // diffuerase.py:75 - resize-back only
        return vv_resize(inp, T, h, w, 3, out, H0, W0, VV_INTER_LINEAR, workspace, workspace_bytes, stream);
    VV_CHECK_ARG(orig && mask, "vv_upscale_feather_composite: orig/mask required when keep_unmasked != 0");
    if (feather_px > VV_MAX_FEATHER) {
        set_error("vv_upscale_feather_composite: feather_px %.3f > supported maximum %.1f", feather_px, VV_MAX_FEATHER);
        return VV_ERR_UNSUPPORTED;
    }
    if (W0 > 65535) {
        set_error("vv_upscale_feather_composite: frames wider than 65535 pixels are not supported");
        return VV_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const Tap *xt = nullptr, *yt = nullptr;
    int rc = VV_OK;
    bool taps_built = false;             // k3_fast / k3_fastw need no x table, and no y table for an exact x2 either

    // the table only depends on feather_px: keep the last one (the GUI never changes it from 3)
    static std::mutex ft_mu;
    static FeatherTable ft_cache;
    static float ft_cache_key = -12345.f;
    FeatherTable ft;
    {
        std::lock_guard<std::mutex> g(ft_mu);
        if (ft_cache_key != feather_px) {
            build_feather_table(feather_px, &ft_cache);
            ft_cache_key = feather_px;
        }
        ft = ft_cache;
    }
    const int Wp = ceil_div(W0, 32);
    if (ft.radius > 7 || (ft.radius > 2 && ft.radius >= max(3, get_option(OPT_K3_BIG_FROM)))) {
        // ---- feather_px in (8, 32] (and smaller radii >= "k3_big_from"): k3_bigfeather with its cost-sorted offset table in device memory (behind the taps)
        static std::vector<BigEntry> big_cache;
        static float big_key = -12345.f;
        BigEntry *dtab = (BigEntry *)((uint8_t *)workspace + align_up(vv_resize_workspace_bytes(H0, W0), 256));
        int n_entries = 0;
        {
            std::lock_guard<std::mutex> g(ft_mu);
            if (big_key != feather_px) {
                const std::vector<std::tuple<float, int, int>> ent = chamfer_entries(feather_px, ft.radius);
                big_cache.clear();
                for (const auto &e : ent) big_cache.push_back(BigEntry{std::get<0>(e), (short)std::get<2>(e), (short)std::get<1>(e)});
                big_key = feather_px;
            }
            n_entries = (int)big_cache.size();
            if (n_entries > K3_BIG_MAX_ENTRIES) {
                set_error("vv_upscale_feather_composite: offset table of feather_px %.3f too large", feather_px);
                return VV_ERR_UNSUPPORTED;
            }
            // pageable source: the runtime stages it before the call returns, so the cache may change afterwards
            cudaError_t ce = cudaMemcpyAsync(dtab, big_cache.data(), (size_t)n_entries * sizeof(BigEntry), cudaMemcpyHostToDevice, st);
            if (ce != cudaSuccess) return fail_cuda(ce, "cudaMemcpyAsync(feather table)");
        }
        if ((rc = build_linear_taps(workspace, h, w, H0, W0, &xt, &yt, st, 3))) return rc;
        const size_t smem = ((size_t)(K3_BIG_TH + 2 * ft.radius) * (Wp + 3)) * 4 + align_up((size_t)K3_BIG_TH * W0 * 2, 4);
        if (smem > 200 * 1024) {
            set_error("vv_upscale_feather_composite: feather_px > 8 supports frames up to about 10000 pixels wide");
            return VV_ERR_UNSUPPORTED;
        }
        int dev_big = 0;
        cudaGetDevice(&dev_big);
        dev_big = min(max(dev_big, 0), 63);
        static std::atomic<size_t> big_smem_set[64];
        if (smem > 48 * 1024 && smem > big_smem_set[dev_big].load()) {
            cudaError_t e = cudaFuncSetAttribute(k3_bigfeather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(k3_bigfeather)");
            big_smem_set[dev_big].store(smem);
        }
        const int words_ok = (W0 % 4 == 0) && ((uintptr_t)orig % 4 == 0) && ((uintptr_t)out % 4 == 0);
        const int mask_vec = (W0 % 16 == 0) && ((uintptr_t)mask % 16 == 0);
        const int bstrips = ceil_div(H0, K3_BIG_TH);
        for (int t0 = 0; t0 < T; t0 += 32768) {          // grid.y <= 65535 frames per launch
            const int tn = min(32768, T - t0);
            const size_t fo = (size_t)t0 * H0 * W0;
            k3_bigfeather<<<dim3((unsigned)bstrips, (unsigned)tn), K3_BIG_THREADS, smem, st>>>(
                inp + (size_t)t0 * h * w * 3, orig + fo * 3, mask + fo, mask_bits ? mask_bits + (size_t)t0 * H0 * Wp : nullptr,
                out + fo * 3, xt, yt, dtab, n_entries, ft.radius, ft.div, h, w, H0, W0, words_ok, mask_vec);
            VV_POST_LAUNCH("k3_bigfeather");
        }
        return VV_OK;
    }
    const bool small_r = feather_px > 0.f && ft.radius <= 2;
    const int R = small_r ? 2 : ft.radius;
    const bool vec = (W0 % 16 == 0) && ((uintptr_t)orig % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)mask % 16 == 0);
    // TMA-staged variant: strip rows chosen so that strip + bit rows + queues fit twice per SM
    int th = K3_TH;
    bool tma = vec && get_option(OPT_K3_TMA) != 0;
    int tma_threads = K3_THREADS_TMA;
    const size_t common_words = 16 + 2;
    size_t smem = 0;
    if (tma) {
        const int th_max = min(16, max(2, get_option(OPT_K3_TMA_ROWS)));
        tma_threads = get_option(OPT_K3_TMA_THREADS) >= 512 ? 512 : get_option(OPT_K3_TMA_THREADS) >= 384 ? 384 : 256;
        for (th = th_max; th >= 2; --th) {      // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2
            smem = (size_t)th * W0 * 3 + 16 +
                   ((size_t)(th + 2 * R) * (Wp + 2) + common_words + (tma_threads / 32) * (K3_QUEUE1 + 32) * 2) * 4;
            if (smem <= 113 * 1024) break;
        }
        if (th < 2) tma = false;
    }
    const int nt = tma ? 1 : (get_option(OPT_K3_NT) == 1 ? 1 : 2);
    if (!tma) {
        th = K3_TH;
        smem = ((size_t)(th + 2 * R) * (Wp + 2) + common_words + (K3_THREADS / 32) * (K3_QUEUE1 * nt + 32) * 2) * 4;
    }
    const int strips = ceil_div(H0, th);
    const long long grid = (long long)T * strips;
    VV_CHECK_ARG(grid < 2147483647LL, "vv_upscale_feather_composite: too many strips");

    int dev_id = 0;
    cudaGetDevice(&dev_id);
    dev_id = min(max(dev_id, 0), 63);
    // cudaFuncSetAttribute is per device: remember the opt-in shared-memory size per (kernel, device)
#define VV_K3_SMEM(kfn)                                                                                         \
    do {                                                                                                        \
        static std::atomic<size_t> smem_set[64];                                                                \
        if (smem > 48 * 1024 && smem > smem_set[dev_id].load()) {                                               \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(k3)");                              \
            smem_set[dev_id].store(smem);                                                                       \
        }                                                                                                       \
    } while (0)

    // ---- k3_fast: exact x2 horizontal up-scale, feather radius <= 2, TMA-able frames (the production geometry)
    const int x2opt = get_option(OPT_K3_X2);
    const int hr = W0 == 2 * w ? 2 : (W0 == 4 * w ? 4 : 0);
    if (vec && small_r && x2opt >= 2 && get_option(OPT_K3_TMA) != 0 && hr != 0 && w >= 4 && (w * 3) % 4 == 0 &&
        ((uintptr_t)inp % 4 == 0)) {
        // dynamic shared memory (words): [mbarrier 4][strip][bit rows (k3_fastw: M rows, Z rows)][pad to 4]
        //                                [lut 16 x float4 + lut_pos 4][queues]
        int fth = 0;
        size_t fsm = 0;
        FastGeom gm = {};
        // word tasks (k3_fastw, default) or 16-pixel rolling tasks (k3_fast, option k3_x2 = 2)
        const bool wordtasks = x2opt >= 3 && W0 <= 4096;
        // 512 threads x 2 CTAs per SM (default) or 256 threads x 4 CTAs per SM with shorter strips
        const int thr_opt = get_option(OPT_K3_TMA_THREADS);
        const int nth = (hr == 2 && thr_opt <= 256) ? 256 : (hr == 2 && wordtasks && thr_opt == 384) ? 384 : 512;
        const size_t smem_cap = nth == 256 ? 56 * 1024 : nth == 384 ? 75 * 1024 : 113 * 1024;
        for (fth = min(16, max(2, get_option(OPT_K3_TMA_ROWS))); fth >= 2; --fth) {
            gm.bits_off = fth * W0 * 3 / 4 + 4;
            gm.zbits_off = gm.bits_off + (fth + 4) * (Wp + 2);
            gm.lut_off = ((wordtasks ? gm.zbits_off : gm.bits_off) + (fth + 4) * (Wp + 2) + 3) & ~3;
            gm.taps_off = gm.lut_off + 16 * 4 + 4;                       // k3_fastw: 16 x Tap
            gm.queue_off = gm.taps_off + (wordtasks ? 32 : 0);
            fsm = ((size_t)gm.queue_off + (nth / 32) * (wordtasks ? K3W_QCAP : K3F_QCAP * 2)) * 4;
            if (fsm <= smem_cap) break;
        }
        if (fth >= 2) {
            const int fstrips = ceil_div(H0, fth);
            const bool vx2 = hr == 2 && H0 == 2 * h;
            const bool bits = mask_bits != nullptr && get_option(OPT_K3_BITS) != 0;
            if ((rc = build_linear_taps(workspace, h, w, H0, W0, &xt, &yt, st, vx2 ? 0 : 2))) return rc;
            taps_built = true;
            // rows per classification task: 4 when that still gives most threads a task, else 2
            int rpt = 4;
            for (int cand : {4, 3, 2})                      // fewest idle threads in the first pass over the tasks
                if ((W0 / 16) * ceil_div(fth, cand) <= nth && (W0 / 16) * ceil_div(fth, cand) * 10 >= nth * 7) {
                    rpt = cand;
                    break;
                }
            if ((W0 / 16) * ceil_div(fth, 4) > nth) rpt = 4;
            if (get_option(OPT_K3_NT) > 2) rpt = min(16, get_option(OPT_K3_NT));      // A/B switch
            smem = fsm;
            gm.h = h, gm.w = w, gm.H0 = H0, gm.W0 = W0, gm.th = fth, gm.rpt = rpt;
            gm.Wpc = Wp, gm.row_words = Wp + 2, gm.rows_s = fth + 4;
            gm.G = W0 / 16, gm.n_tasks = gm.G * ceil_div(fth, rpt);
            gm.n_steps = ceil_div(gm.n_tasks, nth) * rpt;
            gm.frame_bytes = (long long)H0 * W0 * 3, gm.mask_frame_bytes = (long long)H0 * W0;
            gm.inp_frame_bytes = (long long)h * w * 3, gm.bits_frame_words = (long long)H0 * Wp;
            gm.div = ft.div, gm.one = 1.0f;
            if (wordtasks) {      // one 32-pixel word of one strip row per lane and step
                gm.n_tasks = Wp * fth;
                gm.n_steps = ceil_div(gm.n_tasks, nth);
                gm.inv_wpc = (uint32_t)(0x100000000ULL / (unsigned)Wp) + 1u;
                gm.inv_rw = (uint32_t)(0x100000000ULL / (unsigned)(Wp + 2)) + 1u;
                host_alpha_levels(ft.div, gm.alpha, &gm.alpha_pos);
            }
#define VV_K3_FAST(KERNEL, V, B, N, H)                                                                          \
    do {                                                                                                        \
        auto kfn = KERNEL<V, B, N, H>;                                                                          \
        VV_K3_SMEM(kfn);                                                                                        \
        for (int t0 = 0; t0 < T; t0 += 32768) {        /* grid.y <= 65535 frames per launch */                  \
            const int tn = min(32768, T - t0);                                                                  \
            const size_t fo = (size_t)t0 * H0 * W0;                                                             \
            cudaLaunchConfig_t cfg = {};                                                                        \
            cfg.gridDim = dim3((unsigned)fstrips, (unsigned)tn);                                                \
            cfg.blockDim = dim3(N);                                                                             \
            cfg.dynamicSmemBytes = smem;                                                                        \
            cfg.stream = st;                                                                                    \
            cudaLaunchAttribute attr[1];                                                                        \
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                    \
            attr[0].val.programmaticStreamSerializationAllowed = 1;                                             \
            cfg.attrs = attr;                                                                                   \
            cfg.numAttrs = chain ? 1 : 0;                                                                       \
            cudaError_t le = cudaLaunchKernelEx(&cfg, kfn, inp + (size_t)t0 * h * w * 3, orig + fo * 3, mask + fo,  \
                                                mask_bits ? mask_bits + (size_t)t0 * H0 * Wp : (const uint32_t *)nullptr, \
                                                out + fo * 3, yt, gm);                                          \
            if (le != cudaSuccess) return fail_cuda(le, "cudaLaunchKernelEx(" #KERNEL ")");                     \
            VV_POST_LAUNCH(#KERNEL);                                                                            \
        }                                                                                                       \
    } while (0)
#define VV_K3_FAST_B(KERNEL, V, N, H)      \
    do {                                   \
        if (bits)                          \
            VV_K3_FAST(KERNEL, V, true, N, H);  \
        else                               \
            VV_K3_FAST(KERNEL, V, false, N, H); \
    } while (0)
#define VV_K3_FAST_ALL(KERNEL)                       \
    do {                                             \
        if (hr == 4)  /* x4: table-driven vertical taps, 512 threads */ \
            VV_K3_FAST_B(KERNEL, false, 512, 4);     \
        else if (vx2 && nth == 256)                  \
            VV_K3_FAST_B(KERNEL, true, 256, 2);      \
        else if (vx2 && nth == 384)                  \
            VV_K3_FAST_B(k3_fastw, true, 384, 2);    \
        else if (nth == 384)                         \
            VV_K3_FAST_B(k3_fastw, false, 384, 2);   \
        else if (vx2)                                \
            VV_K3_FAST_B(KERNEL, true, 512, 2);      \
        else if (nth == 256)                         \
            VV_K3_FAST_B(KERNEL, false, 256, 2);     \
        else                                         \
            VV_K3_FAST_B(KERNEL, false, 512, 2);     \
    } while (0)
            // chained launch: only ever set for a k3_fastw launch that directly follows another one in the stream
            const bool chain = wordtasks && get_option(OPT_K3_CHAIN) != 0;
            if (wordtasks)
                VV_K3_FAST_ALL(k3_fastw);
            else
                VV_K3_FAST_ALL(k3_fast);
#undef VV_K3_FAST_ALL
#undef VV_K3_FAST_B
#undef VV_K3_FAST
            return VV_OK;
        }
    }

    if (!taps_built && (rc = build_linear_taps(workspace, h, w, H0, W0, &xt, &yt, st, 3))) return rc;
#define VV_K3_LAUNCH(V, S, N, M, ...)                                                                                \
    do {                                                                                                        \
        auto kfn = k3_upscale_feather_composite<V, S, N, M, ##__VA_ARGS__>;                                                    \
        VV_K3_SMEM(kfn);                                                                                        \
        kfn<<<(unsigned)grid, (M) ? tma_threads : K3_THREADS, smem, st>>>(inp, orig, mask, out, xt, yt, h, w, H0,   \
                                                                            W0, strips, th, ft);                \
    } while (0)
#define VV_K3_DISPATCH(V, S)            \
    do {                                \
        if (nt == 1)                    \
            VV_K3_LAUNCH(V, S, 1, false); \
        else                            \
            VV_K3_LAUNCH(V, S, 2, false); \
    } while (0)
    const bool x2 = tma && small_r && W0 == 2 * w && H0 == 2 * h && ((uintptr_t)inp % 4 == 0) && x2opt != 0;
    if (x2)
        VV_K3_LAUNCH(true, true, 1, true, true);
    else if (tma && small_r)
        VV_K3_LAUNCH(true, true, 1, true);
    else if (tma)
        VV_K3_LAUNCH(true, false, 1, true);
    else if (vec && small_r)
        VV_K3_DISPATCH(true, true);
    else if (vec)
        VV_K3_DISPATCH(true, false);
    else if (small_r)
        VV_K3_DISPATCH(false, true);
    else
        VV_K3_DISPATCH(false, false);
#undef VV_K3_DISPATCH
#undef VV_K3_LAUNCH
#undef VV_K3_SMEM
    VV_POST_LAUNCH("k3_upscale_feather_composite");
    return VV_OK;
}

extern "C" int vv_upscale_feather_composite(const uint8_t *inp, int T, int h, int w, const uint8_t *orig,
                                            const uint8_t *mask, int H0, int W0, float feather_px, int keep_unmasked,
                                            uint8_t *out, void *workspace, size_t workspace_bytes, void *stream) {
    return composite_impl(inp, T, h, w, orig, mask, nullptr, H0, W0, feather_px, keep_unmasked, out, workspace,
                          workspace_bytes, stream);
}

extern "C" int vv_upscale_feather_composite_bits(const uint8_t *inp, int T, int h, int w, const uint8_t *orig,
                                                 const uint8_t *mask, const uint32_t *mask_bits, int H0, int W0,
                                                 float feather_px, int keep_unmasked, uint8_t *out, void *workspace,
                                                 size_t workspace_bytes, void *stream) {
    return composite_impl(inp, T, h, w, orig, mask, mask_bits, H0, W0, feather_px, keep_unmasked, out, workspace,
                          workspace_bytes, stream);
}

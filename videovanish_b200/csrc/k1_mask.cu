// K1: mask binarise + L1 dilation.   Replaces /root/reference/diffuerase.py:28-31
//
//   m = np.any(m > 0, axis=2)                                   (:29)
//   m = scipy.ndimage.binary_dilation(m, iterations=N) * 255    (:30)
//
// Two streaming passes over HBM with a bit plane (1 bit / pixel) in between:
//   k1a binarize_pack   u8 [T,H,W,C] -> u32 bit plane [T*H,Wp]          reads C B/px, writes 1/8 B/px
//   k1b dilate_expand   bit plane -> u8 [T,H,W] in {0,255}               reads ~1/8 B/px, writes 1 B/px
// so the C-byte mask is read exactly once whatever the dilation radius (no halo re-reads of
// the wide input); halos are paid on the bit plane, which is 24x smaller and L2 resident.
// The dilation itself is N rounds of the 4-connected cross done entirely in registers:
// one warp owns a 32-word x ROWS tile (lane <-> word column, rows unrolled in registers),
// horizontal carries come from the neighbour lanes by warp shuffle; no shared memory and no
// block barriers.  Cells outside the frame behave as free space, which cannot change any
// in-frame result because L1 shortest paths between in-frame pixels stay inside the frame.
//
// Bit plane layout: row r (= t*H + y) holds Wp = ceil(W/32) little-endian u32 words, pixel x
// is bit (x & 31) of word (x >> 5); bits at x >= W are zero.
#include "common.cuh"

namespace vv {

// ------------------------------------------------------------------ k1a: binarise + pack
// One thread packs 16 pixels into one u16 of the bit plane (u16 index = 2*word + half).
template <int C, bool VEC>
__global__ void __launch_bounds__(256) k1a_binarize_pack(const uint8_t *__restrict__ mask, uint16_t *__restrict__ bits,
                                                         int W, int halves_per_row, long long n_rows) {
    const long long total = n_rows * halves_per_row;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / halves_per_row;
        const int hw = (int)(idx - row * halves_per_row);
        const int x0 = hw * 16;
        uint32_t out = 0;
        if (x0 < W) {
            const uint8_t *p = mask + (row * W + x0) * C;
            if (VEC && x0 + 16 <= W) {
                if (C == 1) {
                    out = nonzero_bits16(ldg128(p));
                } else if (C == 3) {
                    const uint4 a = ldg128(p), b = ldg128(p + 16), c = ldg128(p + 32);
                    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int g = 0; g < 4; ++g) {   // 4 pixels per 3 words
                        const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
                        out |= (uint32_t)((w0 & 0x00ffffffu) != 0) << (4 * g);
                        out |= (uint32_t)(((w0 & 0xff000000u) | (w1 & 0x0000ffffu)) != 0) << (4 * g + 1);
                        out |= (uint32_t)(((w1 & 0xffff0000u) | (w2 & 0x000000ffu)) != 0) << (4 * g + 2);
                        out |= (uint32_t)((w2 & 0xffffff00u) != 0) << (4 * g + 3);
                    }
                } else {   // C == 4: one word per pixel
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 v = ldg128(p + 16 * q);
                        out |= ((uint32_t)(v.x != 0) | ((uint32_t)(v.y != 0) << 1) | ((uint32_t)(v.z != 0) << 2) |
                                ((uint32_t)(v.w != 0) << 3))
                               << (4 * q);
                    }
                }
            } else {
                const int n = min(16, W - x0);
                for (int i = 0; i < n; ++i) {
                    uint32_t any = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) any |= p[i * C + c];
                    out |= (uint32_t)(any != 0) << i;
                }
            }
        }
        bits[idx] = (uint16_t)out;
    }
}

// ------------------------------------------------------------------ k1b: dilate (+ expand)
// Warp tile: lanes 1..30 own output word columns tx*30 + (lane-1); lanes 0 and 31 carry the
// 32-pixel horizontal halo (enough for n_iter <= 32; what they compute themselves is never used).
// Register row i <-> frame row y0 - NMAX + i.  After n_iter rounds rows [NMAX, NMAX+RB) of lanes
// 1..30 are exact.  EXACT (n_iter == NMAX at compile time) unrolls the rounds as well, which lets
// the compiler drop the halo-row updates that can no longer reach an output row.
// HALF: additionally write the exact x2 INTER_NEAREST down-size (source index 2*d) of the dilated
// mask: even bits of even rows, 16 low-res pixels per word.
__device__ __forceinline__ uint32_t even_bits(uint32_t x) {   // bits 0,2,4,...,30 -> bits 0..15
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

template <int NMAX, int ROWS>
__device__ __forceinline__ void cross_round(uint32_t (&r)[ROWS]) {
    uint32_t prev = 0;
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        const uint32_t c = r[i];
        const uint32_t l = __shfl_up_sync(0xffffffffu, c, 1);      // lane 0 / 31 receive their own word:
        const uint32_t rt = __shfl_down_sync(0xffffffffu, c, 1);   // only halo lanes are affected
        const uint32_t nxt = (i + 1 < ROWS) ? r[i + 1] : 0u;
        // (c << 1 | l >> 31) | (c >> 1 | rt << 31) | c | up | down
        r[i] = (__funnelshift_l(l, c, 1) | __funnelshift_r(c, rt, 1) | c) | (prev | nxt);
        prev = c;
    }
}

// Diamond of radius 2K + 1 in 10 shift-OR steps instead of 2K + 1 cross rounds.  In the rotated lattice the L1 ball is a
// square: {|x| + |y| <= 2K, x + y even} = S1 (+) S2 with the diagonal segments S1 = {i (1,1)}, S2 = {j (1,-1)}, |i|, |j| <= K,
// and one cross round on top fills the odd parity and reaches radius 2K + 1.  A segment of L = 2K + 1 points grows by
// doubling (1, 2, 4, 8, then L - 8 more: K = 4..7), one-sided towards +x; dilation commutes with translation, so both
// segments are grown one-sided and the plane is shifted back by 2K pixels once at the end.  A diagonal step costs one
// shuffle, one funnel shift and one OR per row (a cross round: two shuffles, two funnel shifts, three ORs).
// Everything an output word of lanes 1..30 depends on lies within 2K <= 14 pixels and rows of it at every stage, i.e.
// inside the halo lanes / rows; what the halo lanes pick up from their missing neighbour travels at most 4K <= 28 pixels
// and never reaches a word that is read for an output.
template <int D, int ROWS>
__device__ __forceinline__ void diag_step_down(uint32_t (&r)[ROWS]) {   // a set pixel (x, y) sets (x + D, y + D)
#pragma unroll
    for (int i = ROWS - 1; i >= D; --i) {          // rows descending: in place
        const uint32_t c = r[i - D];
        r[i] |= __funnelshift_l(__shfl_up_sync(0xffffffffu, c, 1), c, D);
    }
}
template <int D, int ROWS>
__device__ __forceinline__ void diag_step_up(uint32_t (&r)[ROWS]) {     // (x, y) sets (x + D, y - D)
#pragma unroll
    for (int i = 0; i + D < ROWS; ++i) {           // rows ascending: in place
        const uint32_t c = r[i + D];
        r[i] |= __funnelshift_l(__shfl_up_sync(0xffffffffu, c, 1), c, D);
    }
}
template <int K, int ROWS>
__device__ __forceinline__ void diamond_block(uint32_t (&r)[ROWS]) {
    diag_step_down<1, ROWS>(r), diag_step_down<2, ROWS>(r), diag_step_down<4, ROWS>(r), diag_step_down<2 * K + 1 - 8, ROWS>(r);
    diag_step_up<1, ROWS>(r), diag_step_up<2, ROWS>(r), diag_step_up<4, ROWS>(r), diag_step_up<2 * K + 1 - 8, ROWS>(r);
#pragma unroll
    for (int i = 0; i < ROWS; ++i)                 // back by 2K pixels
        r[i] = __funnelshift_r(r[i], __shfl_down_sync(0xffffffffu, r[i], 1), 2 * K);
}

// radius 6 (K = 3): segments of 7 points = 1, 2, 4, then 3 more
template <int ROWS>
__device__ __forceinline__ void diamond_block6(uint32_t (&r)[ROWS]) {
    diag_step_down<1, ROWS>(r), diag_step_down<2, ROWS>(r), diag_step_down<3, ROWS>(r);
    diag_step_up<1, ROWS>(r), diag_step_up<2, ROWS>(r), diag_step_up<3, ROWS>(r);
#pragma unroll
    for (int i = 0; i < ROWS; ++i) r[i] = __funnelshift_r(r[i], __shfl_down_sync(0xffffffffu, r[i], 1), 6);
}

template <int NMAX, int RB, bool EXACT>
__global__ void __launch_bounds__(128)
    k1b_dilate_expand(const uint32_t *__restrict__ bits_in, uint32_t *__restrict__ bits_out, uint8_t *__restrict__ out,
                      uint8_t *__restrict__ half_out, int H, int W, int Wp, int n_iter, int tiles_x, int tiles_y,
                      long long n_tiles, int vec_ok, int diag) {
    constexpr int ROWS = RB + 2 * NMAX;
    const int lane = threadIdx.x & 31;
    const long long tile = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;   // whole warp exits together
    const int tx = (int)(tile % tiles_x);
    const long long rest = tile / tiles_x;
    const int ty = (int)(rest % tiles_y);
    const long long t = rest / tiles_y;
    const int wx = tx * 30 + lane - 1;
    const int y0 = ty * RB;
    const bool col_ok = wx >= 0 && wx < Wp;
    const uint32_t *src = bits_in + t * H * (long long)Wp + wx;
    if (EXACT) n_iter = NMAX;

    uint32_t r[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        const int y = y0 - NMAX + i;
        const bool need = (i >= NMAX - n_iter) && (i < NMAX + RB + n_iter);
        r[i] = (col_ok && need && y >= 0 && y < H) ? __ldg(src + (long long)y * Wp) : 0u;
    }

    if (EXACT) {
        if (NMAX == 8 && diag == 2) {             // radius 8 = diagonal block of radius 6 + two cross rounds
            diamond_block6<ROWS>(r);
            cross_round<NMAX, ROWS>(r), cross_round<NMAX, ROWS>(r);
        } else {
#pragma unroll
            for (int it = 0; it < NMAX; ++it) cross_round<NMAX, ROWS>(r);
        }
    } else {
        int rounds = n_iter;
        if (NMAX >= 16 && diag) {                 // kernel-uniform: a radius >= 9 starts with the largest diagonal block,
                                                  // which stands for 2K rounds once at least one cross round follows it
            if (rounds >= 15) diamond_block<7, ROWS>(r), rounds -= 14;
            else if (rounds >= 13) diamond_block<6, ROWS>(r), rounds -= 12;
            else if (rounds >= 11) diamond_block<5, ROWS>(r), rounds -= 10;
            else if (rounds >= 9) diamond_block<4, ROWS>(r), rounds -= 8;
        }
#pragma unroll 1
        for (int it = 0; it < rounds; ++it) cross_round<NMAX, ROWS>(r);
    }

    // ---- output stage.  The dilated words go through shared memory so that this part is a short
    // rolled loop (fully unrolled it was ~4k instructions of straight-line code and the kernel
    // stalled on instruction fetch) and so that every 128-bit store instruction of the warp writes
    // one contiguous 512-byte run of the row.
    __shared__ uint32_t s_rows[4][RB][32];
    uint32_t(*my)[32] = s_rows[threadIdx.x >> 5];
    {
        const int valid_bits = W - wx * 32;     // bits beyond the frame edge may have been dilated into
        const uint32_t keep = !col_ok ? 0u : valid_bits >= 32 ? 0xffffffffu : ((1u << valid_bits) - 1u);
#pragma unroll
        for (int i = 0; i < RB; ++i) my[i][lane] = r[NMAX + i] & keep;
    }
    __syncwarp();
    const int lw = W >> 1;
    const int x_tile = tx * 30 * 32;            // first pixel column of the tile's useful words
    const int rows_here = min(RB, H - y0);
#pragma unroll 1
    for (int i = 0; i < rows_here; ++i) {
        const long long row = t * H + y0 + i;
        if (bits_out && lane >= 1 && lane <= 30 && col_ok) bits_out[row * Wp + wx] = my[i][lane];
        if (out) {
            uint8_t *orow = out + row * W;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int k = pass * 32 + lane;          // 16-pixel chunk of the tile row (60 chunks)
                const int px0 = x_tile + k * 16;
                if (k < 60 && px0 < W) {
                    const uint32_t v = (my[i][1 + (k >> 1)] >> (16 * (k & 1))) & 0xffffu;
                    if (vec_ok && px0 + 16 <= W) {
                        stg128_stream(orow + px0, make_uint4(expand4(v), expand4(v >> 4), expand4(v >> 8), expand4(v >> 12)));
                    } else {
                        const int n = min(16, W - px0);
#pragma unroll 1
                        for (int q = 0; q < n; ++q) orow[px0 + q] = ((v >> q) & 1u) ? 255 : 0;
                    }
                }
            }
        }
        if (half_out && !(i & 1) && lane < 30) {          // RB and y0 are even, so (y0 + i) is even iff i is
            const int lx0 = (tx * 30 + lane) * 16;        // low-res column of this lane's 16 pixels
            if (lx0 < lw) {
                const uint32_t e = even_bits(my[i][1 + lane]);
                uint8_t *o = half_out + (t * (H >> 1) + ((y0 + i) >> 1)) * lw + lx0;
                if (vec_ok && lx0 + 16 <= lw) {
                    stg128_stream(o, make_uint4(expand4(e), expand4(e >> 4), expand4(e >> 8), expand4(e >> 12)));
                } else {
                    const int n = min(16, lw - lx0);
#pragma unroll 1
                    for (int q = 0; q < n; ++q) o[q] = ((e >> q) & 1u) ? 255 : 0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ iterations < 1: fill
// scipy repeats the cross until nothing changes: a frame with any set pixel ends up full.
__global__ void __launch_bounds__(256) k1c_any_per_frame(const uint32_t *__restrict__ bits, long long words_per_frame,
                                                         uint32_t *__restrict__ flags) {
    const long long t = blockIdx.y;
    const uint32_t *p = bits + t * words_per_frame;
    uint32_t acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < words_per_frame;
         i += (long long)gridDim.x * blockDim.x)
        acc |= __ldg(p + i);
    acc = __reduce_or_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicOr(flags + t, 1u);
}

__global__ void __launch_bounds__(256) k1c_fill(const uint32_t *__restrict__ flags, uint8_t *__restrict__ out,
                                                long long bytes_per_frame) {
    const long long t = blockIdx.y;
    const uint8_t v = flags[t] ? 255 : 0;
    uint8_t *o = out + t * bytes_per_frame;
    const long long n16 = ((uintptr_t)o % 16 == 0) ? bytes_per_frame / 16 : 0;
    const uint32_t w = v * 0x01010101u;
    const uint4 v4 = make_uint4(w, w, w, w);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        stg128_stream(o + 16 * i, v4);
    for (long long i = n16 * 16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < bytes_per_frame;
         i += (long long)gridDim.x * blockDim.x)
        o[i] = v;
}

// bit plane of a flooded frame: every in-frame bit set (bits at x >= W stay zero)
__global__ void __launch_bounds__(256) k1c_fill_bits(const uint32_t *__restrict__ flags, uint32_t *__restrict__ bits, int H,
                                                     int W, int Wp) {
    const long long t = blockIdx.y;
    const uint32_t on = flags[t] ? 0xffffffffu : 0u;
    const long long n = (long long)H * Wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % Wp), valid = W - k * 32;
        bits[t * n + i] = on & (valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u));
    }
}

// ------------------------------------------------------------------ fused low-res mask
// INTER_NEAREST down-size of the dilated mask (SURVEY row A9 mask path), sampled from the
// dilated bit plane so the full-resolution u8 mask is not re-read.
__global__ void __launch_bounds__(256)
    k1d_lowres_from_bits(const uint32_t *__restrict__ bits, int H, int W, int Wp, uint8_t *__restrict__ low, int lh,
                         int lw, long long T, int vec_ok) {
    // one thread = 16 consecutive low-res pixels of one row -> one 128-bit store
    const int groups = (lw + 15) >> 4;
    const long long total = T * lh * (long long)groups;
    const double sy = __ddiv_rn(1.0, __ddiv_rn((double)lh, (double)H));
    const double sx = __ddiv_rn(1.0, __ddiv_rn((double)lw, (double)W));
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % lh);
        const long long t = q / lh;
        const int srcy = min((int)floor(__dmul_rn((double)y, sy)), H - 1);
        const uint32_t *brow = bits + (t * H + srcy) * Wp;
        const int x0 = g * 16, n = min(16, lw - x0);
        uint32_t m16 = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < n) {
                const int srcx = min((int)floor(__dmul_rn((double)(x0 + i), sx)), W - 1);
                m16 |= ((__ldg(brow + (srcx >> 5)) >> (srcx & 31)) & 1u) << i;
            }
        }
        uint8_t *o = low + (t * lh + y) * (long long)lw + x0;
        if (vec_ok && n == 16) {
            stg128_stream(o, make_uint4(expand4(m16), expand4(m16 >> 4), expand4(m16 >> 8), expand4(m16 >> 12)));
        } else {
#pragma unroll 1
            for (int i = 0; i < n; ++i) o[i] = ((m16 >> i) & 1u) ? 255 : 0;
        }
    }
}

// Same for W == 2 * lw (1080p -> 960x536, the wrapper's real inference size): the NEAREST source column of
// low-res pixel d is exactly 2d, so 16 low-res pixels are the even bits of ONE word of the source row; only the
// row mapping needs the table rule.  One thread = one 128-bit store.
__global__ void __launch_bounds__(256)
    k1d_lowres_x2_from_bits(const uint32_t *__restrict__ bits, int H, int Wp, uint8_t *__restrict__ low, int lh, int lw,
                            long long T) {
    const int groups = lw >> 4;                       // lw % 16 == 0
    const long long total = T * lh * (long long)groups;
    const double sy = __ddiv_rn(1.0, __ddiv_rn((double)lh, (double)H));
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % lh);
        const long long t = q / lh;
        const int srcy = min((int)floor(__dmul_rn((double)y, sy)), H - 1);
        const uint32_t e = even_bits(__ldg(bits + (t * H + srcy) * Wp + g));
        stg128_stream(low + (t * lh + y) * (long long)lw + g * 16,
                      make_uint4(expand4(e), expand4(e >> 4), expand4(e >> 8), expand4(e >> 12)));
    }
}

template <int C>
static int launch_k1a(const uint8_t *mask, uint16_t *bits, int W, int Wp, long long n_rows, bool vec, int grid,
                      cudaStream_t st) {
    if (vec)
        k1a_binarize_pack<C, true><<<grid, 256, 0, st>>>(mask, bits, W, 2 * Wp, n_rows);
    else
        k1a_binarize_pack<C, false><<<grid, 256, 0, st>>>(mask, bits, W, 2 * Wp, n_rows);
    VV_POST_LAUNCH("k1a_binarize_pack");
    return VV_OK;
}

template <int NMAX, int RB, bool EXACT>
static int launch_k1b(const uint32_t *in, uint32_t *bits_out, uint8_t *out, uint8_t *half_out, int T, int H, int W,
                      int Wp, int n_iter, bool vec, cudaStream_t st) {
    const int tiles_x = ceil_div(Wp, 30), tiles_y = ceil_div(H, RB);
    const long long n_tiles = (long long)T * tiles_x * tiles_y;
    const int grid = ceil_div(n_tiles, 4);
    k1b_dilate_expand<NMAX, RB, EXACT><<<grid, 128, 0, st>>>(in, bits_out, out, half_out, H, W, Wp, n_iter, tiles_x,
                                                             tiles_y, n_tiles, vec ? 1 : 0, get_option(OPT_K1B_DIAG));
    VV_POST_LAUNCH("k1b_dilate_expand");
    return VV_OK;
}

static int dilate_pass(const uint32_t *in, uint32_t *bits_out, uint8_t *out, uint8_t *half_out, int T, int H, int W,
                       int Wp, int n, bool vec, cudaStream_t st) {
    if (n == 8 && get_option(OPT_K1B_EXACT)) return launch_k1b<8, 32, true>(in, bits_out, out, half_out, T, H, W, Wp, n, vec, st);   // GUI default
    if (n <= 4) return launch_k1b<4, 32, false>(in, bits_out, out, half_out, T, H, W, Wp, n, vec, st);
    if (n <= 8) return launch_k1b<8, 32, false>(in, bits_out, out, half_out, T, H, W, Wp, n, vec, st);
    if (n <= 16) return launch_k1b<16, 32, false>(in, bits_out, out, half_out, T, H, W, Wp, n, vec, st);
    return launch_k1b<32, 16, false>(in, bits_out, out, half_out, T, H, W, Wp, n, vec, st);
}

}  // namespace vv

using namespace vv;

extern "C" size_t vv_binarize_dilate_workspace_bytes(int T, int H, int W) {
    if (T <= 0 || H <= 0 || W <= 0) return 0;
    const size_t plane = align_up((size_t)T * H * ceil_div(W, 32) * 4, 256);
    return 2 * plane + align_up((size_t)T * 4, 256);
}

extern "C" int vv_binarize_dilate(const uint8_t *mask, int T, int H, int W, int C, int iterations, uint8_t *out,
                                  uint8_t *lowres_out, int lh, int lw, void *workspace, size_t workspace_bytes,
                                  void *stream) {
    return vv_binarize_dilate_ex(mask, T, H, W, C, iterations, out, lowres_out, lh, lw, nullptr, workspace,
                                 workspace_bytes, stream);
}

extern "C" int vv_binarize_dilate_ex(const uint8_t *mask, int T, int H, int W, int C, int iterations, uint8_t *out,
                                     uint8_t *lowres_out, int lh, int lw, uint32_t *bits_out, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    VV_CHECK_ARG(mask && out && workspace, "vv_binarize_dilate: NULL pointer");
    VV_CHECK_ARG(T > 0 && H > 0 && W > 0, "vv_binarize_dilate: bad shape T=%d H=%d W=%d", T, H, W);
    VV_CHECK_ARG(C == 1 || C == 3 || C == 4, "vv_binarize_dilate: C must be 1, 3 or 4 (got %d)", C);
    VV_CHECK_ARG(workspace_bytes >= vv_binarize_dilate_workspace_bytes(T, H, W),
                 "vv_binarize_dilate: workspace too small");
    VV_CHECK_ARG(!lowres_out || (lh > 0 && lw > 0), "vv_binarize_dilate: bad low-res size %dx%d", lh, lw);
    cudaStream_t st = (cudaStream_t)stream;
    const int Wp = ceil_div(W, 32);
    const size_t plane = align_up((size_t)T * H * Wp * 4, 256);
    uint32_t *bits0 = (uint32_t *)workspace;
    uint32_t *bits1 = (uint32_t *)((uint8_t *)workspace + plane);
    uint32_t *flags = (uint32_t *)((uint8_t *)workspace + 2 * plane);
    const long long n_rows = (long long)T * H;
    const bool vec_in = (W % 16 == 0) && ((uintptr_t)mask % 16 == 0);
    const bool vec_out = (W % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const int grid_a = (int)min((long long)ceil_div(n_rows * 2 * Wp, 256), (long long)148 * 64);
    int rc = C == 1   ? launch_k1a<1>(mask, (uint16_t *)bits0, W, Wp, n_rows, vec_in, grid_a, st)
             : C == 3 ? launch_k1a<3>(mask, (uint16_t *)bits0, W, Wp, n_rows, vec_in, grid_a, st)
                      : launch_k1a<4>(mask, (uint16_t *)bits0, W, Wp, n_rows, vec_in, grid_a, st);
    if (rc) return rc;

    if (iterations < 1) {   // until convergence == flood the frame if anything is set
        cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)T * 4, st);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
        const long long wpf = (long long)H * Wp;
        dim3 g1((unsigned)min((long long)ceil_div(wpf, 256), 64LL), (unsigned)T);
        k1c_any_per_frame<<<g1, 256, 0, st>>>(bits0, wpf, flags);
        VV_POST_LAUNCH("k1c_any_per_frame");
        const long long bpf = (long long)H * W;
        dim3 g2((unsigned)min((long long)ceil_div(bpf, 256 * 16), 256LL), (unsigned)T);
        k1c_fill<<<g2, 256, 0, st>>>(flags, out, bpf);
        VV_POST_LAUNCH("k1c_fill");
        if (lowres_out) {
            dim3 g3((unsigned)min((long long)ceil_div((long long)lh * lw, 256 * 16), 256LL), (unsigned)T);
            k1c_fill<<<g3, 256, 0, st>>>(flags, lowres_out, (long long)lh * lw);
            VV_POST_LAUNCH("k1c_fill");
        }
        if (bits_out) {
            dim3 g4((unsigned)min((long long)ceil_div(wpf, 256), 64LL), (unsigned)T);
            k1c_fill_bits<<<g4, 256, 0, st>>>(flags, bits_out, H, W, Wp);
            VV_POST_LAUNCH("k1c_fill_bits");
        }
        return VV_OK;
    }

    // Chains of <= 16 rounds on the bit planes (diamond_a (+) diamond_b == diamond_{a+b}); the last pass
    // expands to bytes.  16 keeps the register tile at 32 output rows + 2x16 halo rows: a single pass of
    // radius 25 would recompute a 50-row halo for every 16 output rows.
    uint32_t *cur = bits0, *nxt = bits1;
    int left = iterations;
    while (left > 16) {
        const int step_n = left > 32 ? 16 : (left + 1) / 2;      // split the tail evenly (25 -> 13 + 12)
        rc = dilate_pass(cur, nxt, nullptr, nullptr, T, H, W, Wp, step_n, vec_out, st);
        if (rc) return rc;
        uint32_t *tmp = cur;
        cur = nxt, nxt = tmp;
        left -= step_n;
    }
    // exact x2 down-size: the low-res mask is written by the dilation pass itself
    const bool half = lowres_out && H == 2 * lh && W == 2 * lw;
    const bool vec_half = vec_out && (lw % 16 == 0) && ((uintptr_t)lowres_out % 16 == 0);
    // the dilated bit plane goes to the caller (K3 consumes it instead of the u8 mask) and / or feeds the low-res mask
    uint32_t *dil_plane = bits_out ? bits_out : ((lowres_out && !half) ? nxt : nullptr);
    rc = dilate_pass(cur, dil_plane, out, half ? lowres_out : nullptr, T, H, W, Wp, left, half ? vec_half : vec_out, st);
    if (rc) return rc;
    if (lowres_out && !half) {
        const long long total = (long long)T * lh * ((lw + 15) / 16);
        const int vec_low = (lw % 16 == 0) && ((uintptr_t)lowres_out % 16 == 0);
        const int grid_d = (int)min((long long)ceil_div(total, 256), (long long)148 * 32);
        if (W == 2 * lw && vec_low) {
            k1d_lowres_x2_from_bits<<<grid_d, 256, 0, st>>>(dil_plane, H, Wp, lowres_out, lh, lw, T);
            VV_POST_LAUNCH("k1d_lowres_x2_from_bits");
        } else {
            k1d_lowres_from_bits<<<grid_d, 256, 0, st>>>(dil_plane, H, W, Wp, lowres_out, lh, lw, T, vec_low);
            VV_POST_LAUNCH("k1d_lowres_from_bits");
        }
    }
    return VV_OK;
}

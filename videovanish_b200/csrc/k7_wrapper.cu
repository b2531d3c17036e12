// "Next" row N4 of SURVEY.md section 8f: the DiffuEraser wrapper's pixel steps either side of the
// diffusion call made at /root/reference/diffuerase.py:62-67 [recalled-upstream
// diffueraser/diffueraser.py; PARITY UNPINNED by the reference, spec = oracle/wrapper.py]:
//   vv_wrapper_mask     read_mask:  (mask > 0) -> cv2.erode(3x3 rect, 1) -> cv2.dilate(3x3 rect, N) -> * 255
//   vv_wrapper_compose  blended compose:  alpha = u8((1 - (1 - m/255.)(1 - GaussianBlur21(m)/255.)) * 255),
//                       out = u8(img * a + frame * (1 - a)),  a = f32(alpha) / 255
// Both run on bit rows in shared memory.  OpenCV's bit-exact u8 GaussianBlur is integer arithmetic with
// Q0.8 taps (oracle/wrapper.py GAUSS21_Q8), so for a {0,255} mask the blurred value is
// (255 * S + 2^15) >> 16 with S = sum_y k_y sum_x k_x bit(x, y) over the REFLECT_101-padded image.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace vv {

constexpr int K7_THREADS = 512;
constexpr int K7_R = 10;     // blur radius (21 taps)

__device__ __forceinline__ uint32_t k7_row_bits(const uint8_t *__restrict__ row, int x0, int w, bool vec) {
    // bit i = row[x0 + i] != 0, zero beyond w
    uint32_t bits = 0;
    if (vec && x0 + 32 <= w) {
        bits = nonzero_bits16(ldg128(row + x0)) | (nonzero_bits16(ldg128(row + x0 + 16)) << 16);
    } else {
        const int n = min(32, w - x0);
        for (int i = 0; i < n; ++i) bits |= (uint32_t)(row[x0 + i] != 0) << i;
    }
    return bits;
}

// ---- read_mask: binarise, erode once, dilate N times (3x3 rect) ---------------------------------
// One CTA per strip of `th` rows (+ N+1 halo rows each side).  Out-of-image pixels do not erode (they
// count as set) and do not dilate (they count as clear), like cv2's default morphology borders.
__global__ void __launch_bounds__(K7_THREADS)
    k7_wrapper_mask(const uint8_t *__restrict__ mask, uint8_t *__restrict__ out, int h, int w, int N, int th, int vec) {
    extern __shared__ __align__(16) uint32_t k7_smem[];
    const int Wp = (w + 31) >> 5, H = N + 1, R = th + 2 * H;
    uint32_t *A = k7_smem, *B = k7_smem + R * Wp;
    const long long t = blockIdx.y;
    const int y0 = blockIdx.x * th, ybase = y0 - H;
    const uint8_t *mt = mask + t * h * (long long)w;
    const uint32_t tail = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;      // valid bits of the last word
    for (int id = threadIdx.x; id < R * Wp; id += K7_THREADS) {
        const int r = id / Wp, j = id - r * Wp, y = ybase + r;
        uint32_t v = 0xffffffffu;
        if (y >= 0 && y < h) {
            v = k7_row_bits(mt + (long long)y * w, j * 32, w, vec);
            if (j == Wp - 1) v |= ~tail;                                         // beyond the last column: "set"
        }
        A[id] = v;
    }
    __syncthreads();
    // erode: AND over the 3x3 window
    for (int id = threadIdx.x; id < R * Wp; id += K7_THREADS) {
        const int r = id / Wp, j = id - r * Wp, y = ybase + r;
        uint32_t e = 0;
        if (r >= 1 && r <= R - 2 && y >= 0 && y < h) {
            e = 0xffffffffu;
#pragma unroll
            for (int d = -1; d <= 1; ++d) {
                const uint32_t *row = A + (r + d) * Wp;
                const uint32_t m = row[j], lft = j > 0 ? row[j - 1] : 0xffffffffu, rgt = j < Wp - 1 ? row[j + 1] : 0xffffffffu;
                e &= m & ((m << 1) | (lft >> 31)) & ((m >> 1) | (rgt << 31));
            }
            if (j == Wp - 1) e &= tail;
        }
        B[id] = e;
    }
    __syncthreads();
    // dilate N times: OR over the 3x3 window; rows outside the image stay clear
    uint32_t *src = B, *dst = A;
    for (int it = 0; it < N; ++it) {
        for (int id = threadIdx.x; id < R * Wp; id += K7_THREADS) {
            const int r = id / Wp, j = id - r * Wp, y = ybase + r;
            uint32_t v = 0;
            if (r >= 1 && r <= R - 2 && y >= 0 && y < h) {
#pragma unroll
                for (int d = -1; d <= 1; ++d) {
                    const uint32_t *row = src + (r + d) * Wp;
                    const uint32_t m = row[j], lft = j > 0 ? row[j - 1] : 0u, rgt = j < Wp - 1 ? row[j + 1] : 0u;
                    v |= m | (m << 1) | (lft >> 31) | (m >> 1) | (rgt << 31);
                }
                if (j == Wp - 1) v &= tail;
            }
            dst[id] = v;
        }
        __syncthreads();
        uint32_t *tmp = src;
        src = dst, dst = tmp;
    }
    // strip rows -> bytes
    uint8_t *ot = out + t * h * (long long)w;
    for (int id = threadIdx.x; id < th * Wp; id += K7_THREADS) {
        const int r = id / Wp, j = id - r * Wp, y = y0 + r;
        if (y >= h) continue;
        const uint32_t v = src[(r + H) * Wp + j];
        uint8_t *o = ot + (long long)y * w + j * 32;
        if (vec && j * 32 + 32 <= w) {
            stg128_stream(o, make_uint4(expand4(v), expand4(v >> 4), expand4(v >> 8), expand4(v >> 12)));
            stg128_stream(o + 16, make_uint4(expand4(v >> 16), expand4(v >> 20), expand4(v >> 24), expand4(v >> 28)));
        } else {
            const int n = min(32, w - j * 32);
            for (int i = 0; i < n; ++i) o[i] = ((v >> i) & 1u) ? 255 : 0;
        }
    }
}

// ---- blended compose ---------------------------------------------------------------------------
struct ComposeTables {
    uint16_t part[3][128];     // partial horizontal sums of 7-bit slices of the 21-bit window
    uint8_t alpha[256];        // alpha of an unmasked pixel as a function of the blurred value (float64 on the host)
};

__device__ __forceinline__ int k7_reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

// One CTA per strip of `th` rows.  Shared memory: padded bit rows of the strip + 10 halo rows each side,
// the horizontal pass of all of them as u16 (<= 256), the lookup tables.  4 pixels per thread in the
// vertical pass + blend; 8-pixel column groups that see no mask bit anywhere in the strip's window are
// plain copies of the original frame.
template <bool BLENDED>
__global__ void __launch_bounds__(K7_THREADS)
    k7_wrapper_compose(const uint8_t *__restrict__ img, const uint8_t *__restrict__ frames, const uint8_t *__restrict__ mask,
                       uint8_t *__restrict__ out, int h, int w, int th, int hs_stride, int vec, float one,
                       const __grid_constant__ ComposeTables tab) {
    extern __shared__ __align__(16) uint32_t k7_smem[];
    const int Wp = (w + 31) >> 5, roww = Wp + 2, R = th + 2 * K7_R, G8 = (w + 7) >> 3;
    // layout (host: smem_for): hs first so that its rows are 16-byte aligned
    uint16_t *hs = reinterpret_cast<uint16_t *>(k7_smem);                  // [R][hs_stride] (BLENDED only)
    uint32_t *bits = k7_smem + (BLENDED ? (R * hs_stride) / 2 : 0);        // [R][roww], bit p = x + 32
    uint32_t *colflag = bits + R * roww;                                   // [G8][2] bit r: row r of the buffer has a mask bit in the group's window
    float *fa = reinterpret_cast<float *>(colflag + 2 * G8);               // [256] a = alpha / 255
    uint16_t *part = reinterpret_cast<uint16_t *>(fa + 256);               // [3][128]
    uint8_t *alut = reinterpret_cast<uint8_t *>(part + 3 * 128);           // [256]
    const long long t = blockIdx.y;
    const int y0 = blockIdx.x * th;
    const long long npx = (long long)h * w;
    const uint8_t *mt = mask + t * npx;

    for (int i = threadIdx.x; i < 256; i += K7_THREADS) fa[i] = __fdiv_rn((float)i, 255.f);
    for (int i = threadIdx.x; i < 3 * 128; i += K7_THREADS) part[i] = tab.part[i >> 7][i & 127];
    for (int i = threadIdx.x; i < 256; i += K7_THREADS) alut[i] = tab.alpha[i];
    for (int i = threadIdx.x; i < 2 * G8; i += K7_THREADS) colflag[i] = 0;
    if (BLENDED) {
        for (int id = threadIdx.x; id < R * roww; id += K7_THREADS) {
            const int r = id / roww, j = id - r * roww - 1;
            uint32_t v = 0;
            if (j >= 0 && j < Wp) v = k7_row_bits(mt + (long long)k7_reflect101(y0 - K7_R + r, h) * w, j * 32, w, vec);
            bits[id] = v;
        }
        __syncthreads();
        // REFLECT_101 columns: bit(-k) = bit(k), bit(w-1+k) = bit(w-1-k), k = 1..10
        for (int id = threadIdx.x; id < R * 2 * K7_R; id += K7_THREADS) {
            const int r = id / (2 * K7_R), q = id - r * (2 * K7_R), k = (q >> 1) + 1;
            const int xs = (q & 1) ? w - 1 - k : k, xd = (q & 1) ? w - 1 + k : -k;
            uint32_t *row = bits + r * roww;
            if ((row[(xs + 32) >> 5] >> ((xs + 32) & 31)) & 1u) atomicOr(row + ((xd + 32) >> 5), 1u << ((xd + 32) & 31));
        }
        __syncthreads();
        // horizontal pass, 8 pixels per thread: window bit i+d <-> pixel x0+i, tap d
        for (int id = threadIdx.x; id < R * G8; id += K7_THREADS) {
            const int r = id / G8, g = id - r * G8, x0 = g * 8;
            const uint32_t *row = bits + r * roww;
            const int p0 = x0 + 32 - K7_R;
            const uint32_t win = __funnelshift_r(row[p0 >> 5], row[(p0 >> 5) + 1], p0 & 31);
            uint32_t hv[4] = {0, 0, 0, 0};
            if ((win & 0x0fffffffu) == 0x0fffffffu) {          // all 28 bits the group looks at are set: 256 each
                hv[0] = hv[1] = hv[2] = hv[3] = 0x01000100u;
                atomicOr(colflag + 2 * g + (r >> 5), 1u << (r & 31));
            } else if (win) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t wv = win >> i;
                    const uint32_t s = part[wv & 127u] + part[128 + ((wv >> 7) & 127u)] + part[256 + ((wv >> 14) & 127u)];
                    hv[i >> 1] |= s << (16 * (i & 1));
                }
                atomicOr(colflag + 2 * g + (r >> 5), 1u << (r & 31));
            }
            *reinterpret_cast<uint4 *>(hs + r * hs_stride + x0) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
        }
    }
    __syncthreads();

    // vertical pass + blend, 4 pixels per thread: a warp walks the rows, its lanes the 4-pixel groups of a row (no index
    // division).  The frame words (and the model-output words wherever a mask bit is near) are requested BEFORE the
    // vertical pass, so that their latency hides behind its ~100 instructions (ncu: 47 % of the stall samples sat on the
    // first use of these loads); the blend runs in packed fp32 and truncates through the mantissa (add.rm with 2^23)
    // instead of the conversion pipe.
    const int G4 = (w + 3) >> 2;
    const uint8_t *it_ = img + t * npx * 3, *ft = frames + t * npx * 3;
    uint8_t *ot = out + t * npx * 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const f32x2 one2 = pack2(one, one), big2 = pack2(8388608.f, 8388608.f), unbias2 = pack2(-8388608.f, -8388608.f);
    for (int r = warp; r < th; r += K7_THREADS / 32) {
        const int y = y0 + r;
        if (y >= h) break;
        // the mask word and the frame words of a group are requested one trip ahead (the branch on the mask word is the
        // first thing a trip does)
        uint32_t m4n = 0, fwn[3] = {0, 0, 0};
        auto request = [&](int gg) {
            if (vec && gg < G4) {
                const long long pn = ((long long)y * w + gg * 4) * 3;
                m4n = __ldg(reinterpret_cast<const uint32_t *>(mt + (long long)y * w + gg * 4));
                const uint32_t *f32p = reinterpret_cast<const uint32_t *>(ft + pn);
                fwn[0] = __ldg(f32p), fwn[1] = __ldg(f32p + 1), fwn[2] = __ldg(f32p + 2);
            }
        };
        request(lane);
        for (int g = lane; g < G4; g += 32) {
            const int x0 = g * 4;
            const int n = min(4, w - x0);
            const long long po = ((long long)y * w + x0) * 3;
            uint32_t alpha[4] = {0, 0, 0, 0};
            uint32_t m4 = m4n;
            uint32_t fw[3] = {fwn[0], fwn[1], fwn[2]}, iw[3] = {0, 0, 0};
            request(g + 32);
            if (!vec) {
                for (int i = 0; i < n; ++i) m4 |= (uint32_t)mt[(long long)y * w + x0 + i] << (8 * i);
            }
            // some buffer row r .. r+20 (the 21 taps of output row r) sees a mask bit
            const unsigned long long rowflags = ((unsigned long long)colflag[2 * (x0 >> 3) + 1] << 32) | colflag[2 * (x0 >> 3)];
            const bool near_mask = BLENDED && ((rowflags >> r) & 0x1fffffull) != 0;
            if (vec && (near_mask || m4)) {
                const uint32_t *i32 = reinterpret_cast<const uint32_t *>(it_ + po);
                iw[0] = __ldg(i32), iw[1] = __ldg(i32 + 1), iw[2] = __ldg(i32 + 2);
            }
            if (!BLENDED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) alpha[i] = byte_of(m4, i);
            } else if (near_mask && !(byte_of(m4, 0) && byte_of(m4, 1) && byte_of(m4, 2) && byte_of(m4, 3))) {
                // not every pixel is masked: taps 0 and 20 are zero and the kernel is symmetric: rows d and 20-d are added
                // first, two u16 lanes per word.  sum_{d=1..9} K[d] * 512 = 57 856 < 2^16, so the lanes cannot carry into
                // each other; the centre tap is added after unpacking.
                constexpr uint32_t K[11] = {0, 2, 2, 4, 6, 11, 15, 20, 25, 28, 30};
                uint2 acc = make_uint2(0u, 0u);
#pragma unroll
                for (int d = 1; d < 10; ++d) {
                    const uint2 u = *reinterpret_cast<const uint2 *>(hs + (r + d) * hs_stride + x0);
                    const uint2 v = *reinterpret_cast<const uint2 *>(hs + (r + 20 - d) * hs_stride + x0);
                    acc.x += K[d] * (u.x + v.x), acc.y += K[d] * (u.y + v.y);
                }
                const uint2 c = *reinterpret_cast<const uint2 *>(hs + (r + 10) * hs_stride + x0);
                const uint32_t s[4] = {(acc.x & 0xffffu) + K[10] * (c.x & 0xffffu), (acc.x >> 16) + K[10] * (c.x >> 16),
                                       (acc.y & 0xffffu) + K[10] * (c.y & 0xffffu), (acc.y >> 16) + K[10] * (c.y >> 16)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t blur = (255u * s[i] + 32768u) >> 16;
                    alpha[i] = byte_of(m4, i) ? 255u : alut[blur];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) alpha[i] = byte_of(m4, i) ? 255u : 0u;
            }
            const uint32_t amin = min(min(alpha[0], alpha[1]), min(alpha[2], alpha[3]));
            const uint32_t amax = max(max(alpha[0], alpha[1]), max(alpha[2], alpha[3]));
            if (vec) {
                uint32_t *o32 = reinterpret_cast<uint32_t *>(ot + po);
                if (amax == 0u) {                      // a == 0 -> the frame, exactly
                    o32[0] = fw[0], o32[1] = fw[1], o32[2] = fw[2];
                    continue;
                }
                if (amin == 255u) {                    // a == 1 -> the model output, exactly (every pixel masked: loaded above)
                    o32[0] = iw[0], o32[1] = iw[1], o32[2] = iw[2];
                    continue;
                }
                // 0 < a somewhere: blurred alphas only occur near a mask bit, so the model-output words were requested
                float fa4[4], na4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) fa4[i] = fa[alpha[i]], na4[i] = __fsub_rn(1.f, fa4[i]);
                uint32_t rb[12];
#pragma unroll
                for (int p = 0; p < 6; ++p) {          // byte pair p = bytes (2p, 2p+1) of the 12-byte group; byte k belongs to pixel k / 3
                    const int k0 = 2 * p, k1 = 2 * p + 1;
                    const f32x2 a2 = pack2(fa4[k0 / 3], fa4[k1 / 3]), n2 = pack2(na4[k0 / 3], na4[k1 / 3]);
                    const uint32_t wi = iw[p >> 1], wf = fw[p >> 1];
                    const uint32_t s0 = 0x7540u | (uint32_t)(k0 & 3), s1 = 0x7540u | (uint32_t)(k1 & 3);
                    const f32x2 xi = fadd2(pack2u(__byte_perm(wi, 0x4b000000u, s0), __byte_perm(wi, 0x4b000000u, s1)), unbias2);
                    const f32x2 xf = fadd2(pack2u(__byte_perm(wf, 0x4b000000u, s0), __byte_perm(wf, 0x4b000000u, s1)), unbias2);
                    // f32(img * a) + f32(frame * (1 - a)) (two rounded products, one rounded sum), then astype(uint8) =
                    // truncation: 2^23 + v rounded DOWN leaves floor(v) in the mantissa
                    const f32x2 v = fadd2_rm(ffma2(fmul2(xi, a2), one2, fmul2(xf, n2)), big2);
                    unpack2u(v, rb[k0], rb[k1]);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    o32[j] = __byte_perm(__byte_perm(rb[4 * j], rb[4 * j + 1], 0x0040), __byte_perm(rb[4 * j + 2], rb[4 * j + 3], 0x0040),
                                         0x5410);
            } else {
                for (int k = 0; k < 3 * n; ++k) {
                    const float a = fa[alpha[k / 3]], na = __fsub_rn(1.f, a);
                    const float v = __fadd_rn(__fmul_rn((float)it_[po + k], a), __fmul_rn((float)ft[po + k], na));
                    ot[po + k] = (uint8_t)__float2uint_rz(v);
                }
            }
        }
    }
}

static void build_compose_tables(ComposeTables *t) {
    static const int K[21] = {0, 2, 2, 4, 6, 11, 15, 20, 25, 28, 30, 28, 25, 20, 15, 11, 6, 4, 2, 2, 0};
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 128; ++v) {
            int s = 0;
            for (int b = 0; b < 7; ++b)
                if ((v >> b) & 1) s += K[7 * c + b];
            t->part[c][v] = (uint16_t)s;
        }
    // numpy: ((1 - (1 - 0/255.) * (1 - b/255.)) * 255).astype(uint8), all in float64
    for (int b = 0; b < 256; ++b) {
        volatile double blurred = (double)b / 255.0;
        volatile double keep = (1.0 - 0.0 / 255.0) * (1.0 - blurred);
        volatile double soft = 1.0 - keep;
        volatile double scaled = soft * 255.0;
        t->alpha[b] = (uint8_t)scaled;
    }
}

}  // namespace vv

using namespace vv;

extern "C" int vv_wrapper_mask(const uint8_t *mask, int T, int h, int w, int dilation_iter, uint8_t *out, void *stream) {
    VV_CHECK_ARG(mask && out, "vv_wrapper_mask: NULL pointer");
    VV_CHECK_ARG(T > 0 && h > 0 && w > 0 && T <= 65535, "vv_wrapper_mask: bad shape");
    VV_CHECK_ARG(dilation_iter >= 0, "vv_wrapper_mask: negative dilation_iter");
    const int Wp = ceil_div(w, 32), H = dilation_iter + 1;
    // strip height: two bit-row buffers of (th + 2H) rows within 160 KB of shared memory
    const long long max_rows = (160 * 1024) / (2LL * 4 * Wp);
    long long th = min(64LL, max_rows - 2 * H);
    if (th > h) th = h;
    if (th < 1) {
        set_error("vv_wrapper_mask: dilation_iter %d is too large for %d-pixel-wide frames", dilation_iter, w);
        return VV_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)2 * (th + 2 * H) * Wp * 4;
    // cudaFuncSetAttribute is per device: remember the opt-in shared-memory size per device
    static std::atomic<size_t> smem_set[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    dev_id = min(max(dev_id, 0), 63);
    if (smem > 48 * 1024 && smem > smem_set[dev_id].load()) {
        cudaError_t e = cudaFuncSetAttribute(k7_wrapper_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(k7_wrapper_mask)");
        smem_set[dev_id].store(smem);
    }
    const int vec = (w % 16 == 0) && ((uintptr_t)mask % 16 == 0) && ((uintptr_t)out % 16 == 0);
    k7_wrapper_mask<<<dim3(ceil_div(h, th), T), K7_THREADS, smem, (cudaStream_t)stream>>>(mask, out, h, w, dilation_iter, (int)th,
                                                                                       vec);
    VV_POST_LAUNCH("k7_wrapper_mask");
    return VV_OK;
}

extern "C" int vv_wrapper_compose(const uint8_t *img, const uint8_t *frames, const uint8_t *mask255, int T, int h, int w,
                                  int blended, uint8_t *out, void *stream) {
    VV_CHECK_ARG(img && frames && mask255 && out, "vv_wrapper_compose: NULL pointer");
    VV_CHECK_ARG(T > 0 && h > 0 && w > 0 && T <= 65535, "vv_wrapper_compose: bad shape");
    if (blended && (h < K7_R + 1 || w < K7_R + 1)) {
        set_error("vv_wrapper_compose: frames smaller than %d pixels are not supported with blending", K7_R + 1);
        return VV_ERR_UNSUPPORTED;
    }
    static ComposeTables tab;
    static std::once_flag once;
    std::call_once(once, [] { build_compose_tables(&tab); });
    const int Wp = ceil_div(w, 32), G8 = ceil_div(w, 8);
    const int hs_stride = G8 * 8;                       // u16 elements per row, 16-byte aligned rows
    auto smem_for = [&](int th) {
        const size_t R = (size_t)th + 2 * K7_R;
        return (blended ? R * hs_stride * 2 : 0) + R * (Wp + 2) * 4 + (size_t)G8 * 8 + 256 * 4 + 3 * 128 * 2 + 256;
    };
    int th = min(32, h);                                // th + 20 buffer rows must fit the 64-bit row flags
    while (th > 1 && smem_for(th) > 100 * 1024) --th;   // two CTAs per SM when possible ...
    if (smem_for(th) > 100 * 1024) {
        th = min(8, h);
        while (th > 1 && smem_for(th) > 200 * 1024) --th;   // ... one CTA per SM for very wide frames
    }
    const size_t smem = smem_for(th);
    if (smem > 200 * 1024) {
        set_error("vv_wrapper_compose: %d-pixel-wide frames are not supported", w);
        return VV_ERR_UNSUPPORTED;
    }
    const int vec = (w % 4 == 0) && ((uintptr_t)img % 4 == 0) && ((uintptr_t)frames % 4 == 0) && ((uintptr_t)out % 4 == 0) &&
                    ((uintptr_t)mask255 % 16 == 0) && (w % 16 == 0);
    dim3 grid(ceil_div(h, th), T);
    static std::atomic<size_t> smem_set_c[2][64];      // per (variant, device): cudaFuncSetAttribute is per device
    int dev_c = 0;
    cudaGetDevice(&dev_c);
    std::atomic<size_t> *smem_set = &smem_set_c[0][min(max(dev_c, 0), 63)];
    constexpr int SMEM_VARIANT_STRIDE = 64;
    if (smem > 48 * 1024 && smem > smem_set[blended ? SMEM_VARIANT_STRIDE : 0].load()) {
        cudaError_t e = blended ? cudaFuncSetAttribute(k7_wrapper_compose<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(k7_wrapper_compose<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(k7_wrapper_compose)");
        smem_set[blended ? SMEM_VARIANT_STRIDE : 0].store(smem);
    }
    if (blended)
        k7_wrapper_compose<true><<<grid, K7_THREADS, smem, (cudaStream_t)stream>>>(img, frames, mask255, out, h, w, th, hs_stride, vec, 1.0f, tab);
    else
        k7_wrapper_compose<false><<<grid, K7_THREADS, smem, (cudaStream_t)stream>>>(img, frames, mask255, out, h, w, th, hs_stride, vec, 1.0f, tab);
    VV_POST_LAUNCH("k7_wrapper_compose");
    return VV_OK;
}

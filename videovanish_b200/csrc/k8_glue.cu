// Pixel glue between the hot-path kernels and the two networks.
//
// N2  neighbour-window merge of the ProPainter network output (call site /root/reference/diffuerase.py:52-57;
//     [recalled-upstream propainter/inference.py]: for every sliding window of `neighbor_length` frames
//         pred = (net_out + 1) / 2 ;  pred = pred * 255                    (float32)
//         img  = u8(pred) * m + ori * (1 - m)                               (m in {0,1}, truncation)
//         comp = img                      the first time a frame is seen
//         comp = u8(f32(comp) * 0.5 + f32(img) * 0.5)    afterwards         (truncation == (comp + img) >> 1)
//     oracle: oracle/propagation.py ref_neighbor_merge.  PARITY UNPINNED (un-vendored upstream).
// N4  masked frames of the DiffuEraser wrapper's read_mask: frame * (1 - m)  [recalled-upstream
//     diffueraser/diffueraser.py]; oracle/wrapper.py ref_masked_frame.
// N1  channel swap of the frame I/O: cv2.cvtColor(bgr, COLOR_BGR2RGB) of /root/reference/tools.py:21 and the
//     RGB -> BGR swap before VideoWriter.write (:43); the same byte permutation both ways.
#include "common.cuh"

namespace vv {

__device__ __forceinline__ uint32_t unit_to_u8(float p) {
    // ((p + 1) / 2) * 255 in float32, then numpy's astype(uint8) of an in-range value: truncation
    const float v = __fmul_rn(__fmul_rn(__fadd_rn(p, 1.0f), 0.5f), 255.0f);
    return (uint32_t)__float2int_rz(v) & 0xffu;
}

// One thread = 4 consecutive pixels of one frame (VEC: npx % 4 == 0 and 16-byte aligned planes).
template <bool VEC>
__global__ void __launch_bounds__(256)
    k8_neighbor_merge(const float *__restrict__ pred, const uint8_t *__restrict__ mask, const uint8_t *__restrict__ ori,
                      uint8_t *__restrict__ comp, long long npx, int L, unsigned long long first_mask) {
    const long long groups = (npx + 3) >> 2;
    const long long total = groups * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(idx / groups);
        const long long p0 = (idx - l * groups) * 4;
        const bool first = (first_mask >> l) & 1ull;
        const float *pr = pred + (long long)l * 3 * npx;
        const uint8_t *mk = mask + (long long)l * npx + p0;
        const uint8_t *op = ori + ((long long)l * npx + p0) * 3;
        uint8_t *cp = comp + ((long long)l * npx + p0) * 3;
        if (VEC) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(pr + p0));
            const float4 g = __ldg(reinterpret_cast<const float4 *>(pr + npx + p0));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(pr + 2 * npx + p0));
            const uint32_t m4 = __ldg(reinterpret_cast<const uint32_t *>(mk));
            const uint32_t *o3 = reinterpret_cast<const uint32_t *>(op);
            uint32_t o[3] = {__ldg(o3), __ldg(o3 + 1), __ldg(o3 + 2)};
            const float pv[12] = {r.x, g.x, b.x, r.y, g.y, b.y, r.z, g.z, b.z, r.w, g.w, b.w};
            uint32_t img[3] = {0, 0, 0};
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const uint32_t v = byte_of(m4, k / 3) ? unit_to_u8(pv[k]) : byte_of(o[k >> 2], k & 3);
                img[k >> 2] |= v << (8 * (k & 3));
            }
            uint32_t *c3 = reinterpret_cast<uint32_t *>(cp);
            if (!first) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    // per-byte (a + b) >> 1 without carries between bytes: (a & b) + ((a ^ b) >> 1)
                    const uint32_t a = c3[j], bb = img[j];
                    img[j] = (a & bb) + (((a ^ bb) & 0xfefefefeu) >> 1);
                }
            }
            c3[0] = img[0], c3[1] = img[1], c3[2] = img[2];
        } else {
            const int n = (int)min(4LL, npx - p0);
            for (int i = 0; i < n; ++i) {
                for (int c = 0; c < 3; ++c) {
                    uint32_t v = mk[i] ? unit_to_u8(pr[c * npx + p0 + i]) : op[3 * i + c];
                    if (!first) v = ((uint32_t)cp[3 * i + c] + v) >> 1;
                    cp[3 * i + c] = (uint8_t)v;
                }
            }
        }
    }
}

// frames * (1 - m): one thread = 4 pixels.
template <bool VEC>
__global__ void __launch_bounds__(256)
    k8_apply_mask(const uint8_t *__restrict__ frames, const uint8_t *__restrict__ mask, uint8_t *__restrict__ out,
                  long long n_px) {
    const long long groups = (n_px + 3) >> 2;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
        const long long p0 = g * 4;
        if (VEC) {
            const uint32_t m4 = __ldg(reinterpret_cast<const uint32_t *>(mask + p0));
            const uint32_t *f3 = reinterpret_cast<const uint32_t *>(frames + p0 * 3);
            uint32_t a = __ldg(f3), b = __ldg(f3 + 1), c = __ldg(f3 + 2);
            // byte k of the 12 belongs to pixel k / 3
            const uint32_t z0 = byte_of(m4, 0) ? 0u : 0xffu, z1 = byte_of(m4, 1) ? 0u : 0xffu;
            const uint32_t z2 = byte_of(m4, 2) ? 0u : 0xffu, z3 = byte_of(m4, 3) ? 0u : 0xffu;
            a &= z0 * 0x00010101u | z1 << 24;
            b &= z1 * 0x00000101u | z2 * 0x01010000u;
            c &= z2 | z3 * 0x01010100u;
            uint32_t *o3 = reinterpret_cast<uint32_t *>(out + p0 * 3);
            o3[0] = a, o3[1] = b, o3[2] = c;
        } else {
            const int n = (int)min(4LL, n_px - p0);
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) out[(p0 + i) * 3 + c] = mask[p0 + i] ? 0 : frames[(p0 + i) * 3 + c];
        }
    }
}

// R <-> B swap of packed 3-byte pixels: one thread = 16 pixels = 48 bytes (three 128-bit accesses).
__global__ void __launch_bounds__(256)
    k8_swap_rb(const uint8_t *src, uint8_t *dst, long long n_px, int vec_ok) {      // src == dst allowed: plain loads
    const long long groups = (n_px + 15) >> 4;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
        const long long p0 = g * 16;
        if (vec_ok && p0 + 16 <= n_px) {
            const uint8_t *s = src + p0 * 3;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
            const uint4 a = s4[0], b = s4[1], c = s4[2];
            const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            uint32_t o[12];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
                // input bytes : w0 = [R0 G0 B0 R1]  w1 = [G1 B1 R2 G2]  w2 = [B2 R3 G3 B3]
                // output bytes: o0 = [B0 G0 R0 B1]  o1 = [G1 R1 B2 G2]  o2 = [R2 B3 G3 R3]
                o[3 * q] = __byte_perm(w0, w1, 0x5012);           // sel nibbles (lsb first): 2,1,0,5
                o[3 * q + 1] = __byte_perm(__byte_perm(w1, w0, 0x3070), w2, 0x3410);   // [G1 R1 . G2] then B2
                o[3 * q + 2] = __byte_perm(w2, w1, 0x1236);       // 6 = byte 2 of w1 (R2), 3 = B3, 2 = G3, 1 = R3
            }
            uint8_t *d = dst + p0 * 3;
            stg128_stream(d, make_uint4(o[0], o[1], o[2], o[3]));
            stg128_stream(d + 16, make_uint4(o[4], o[5], o[6], o[7]));
            stg128_stream(d + 32, make_uint4(o[8], o[9], o[10], o[11]));
        } else {
            const int n = (int)min(16LL, n_px - p0);
            for (int i = 0; i < n; ++i) {
                const uint8_t r = src[(p0 + i) * 3], gg = src[(p0 + i) * 3 + 1], bb = src[(p0 + i) * 3 + 2];
                dst[(p0 + i) * 3] = bb, dst[(p0 + i) * 3 + 1] = gg, dst[(p0 + i) * 3 + 2] = r;
            }
        }
    }
}

// Rows of each frame that the composite can change: [first row with a mask bit - margin, last such row + margin],
// clipped to the frame, as (lo, hi) with hi exclusive; (0, 0) for an empty mask.  One CTA per frame over K1's 1-bit plane.
// The host-list front end then moves only these rows of the finished frames across PCIe (hostpipe / diffuerase).
__global__ void __launch_bounds__(256) k8_mask_row_bounds(const uint32_t *__restrict__ bits, int H, int Wp, int margin,
                                                          int *__restrict__ bounds) {
    __shared__ int s_lo, s_hi;
    if (threadIdx.x == 0) s_lo = H, s_hi = -1;
    __syncthreads();
    const uint32_t *fb = bits + (long long)blockIdx.x * H * Wp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int lo = H, hi = -1;
    for (int y = warp; y < H; y += 8) {
        uint32_t any = 0;
        for (int k = lane; k < Wp; k += 32) any |= __ldg(fb + (long long)y * Wp + k);
        if (__ballot_sync(0xffffffffu, any != 0)) lo = min(lo, y), hi = max(hi, y);
    }
    if (lane == 0) {
        atomicMin(&s_lo, lo);
        atomicMax(&s_hi, hi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool empty = s_hi < 0;
        bounds[2 * blockIdx.x] = empty ? 0 : max(0, s_lo - margin);
        bounds[2 * blockIdx.x + 1] = empty ? 0 : min(H, s_hi + 1 + margin);
    }
}

// The same from a dilated u8 mask [n,H,W] (the host-list pipeline keeps these resident between pre and post).
__global__ void __launch_bounds__(256) k8_mask_row_bounds_u8(const uint8_t *__restrict__ mask, int H, int W, int *__restrict__ bounds) {
    __shared__ int s_lo, s_hi;
    if (threadIdx.x == 0) s_lo = H, s_hi = -1;
    __syncthreads();
    const uint8_t *fm = mask + (long long)blockIdx.x * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool words = (W % 4 == 0) && ((uintptr_t)mask % 4 == 0);
    int lo = H, hi = -1;
    for (int y = warp; y < H; y += 8) {
        uint32_t any = 0;
        const uint8_t *row = fm + (long long)y * W;
        if (words) {
            for (int k = lane; k < W / 4; k += 32) any |= __ldg(reinterpret_cast<const uint32_t *>(row) + k);
        } else {
            for (int k = lane; k < W; k += 32) any |= row[k];
        }
        if (__ballot_sync(0xffffffffu, any != 0)) lo = min(lo, y), hi = max(hi, y);
    }
    if (lane == 0) {
        atomicMin(&s_lo, lo);
        atomicMax(&s_hi, hi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool empty = s_hi < 0;
        bounds[2 * blockIdx.x] = empty ? 0 : s_lo;
        bounds[2 * blockIdx.x + 1] = empty ? 0 : s_hi + 1;
    }
}

int mask_row_bounds_u8(const uint8_t *mask, int n, int H, int W, int *bounds, cudaStream_t st) {
    k8_mask_row_bounds_u8<<<n, 256, 0, st>>>(mask, H, W, bounds);
    VV_POST_LAUNCH("k8_mask_row_bounds_u8");
    return VV_OK;
}

}  // namespace vv

using namespace vv;

extern "C" int vv_neighbor_merge(const float *pred_chw, const uint8_t *mask, const uint8_t *ori, uint8_t *comp, int L,
                                 int h, int w, unsigned long long first_mask, void *stream) {
    VV_CHECK_ARG(pred_chw && mask && ori && comp, "vv_neighbor_merge: NULL pointer");
    VV_CHECK_ARG(L > 0 && L <= 64 && h > 0 && w > 0, "vv_neighbor_merge: bad shape (1 <= L <= 64 frames per window)");
    const long long npx = (long long)h * w;
    const bool vec = (npx % 4 == 0) && ((uintptr_t)pred_chw % 16 == 0) && ((uintptr_t)mask % 4 == 0) &&
                     ((uintptr_t)ori % 4 == 0) && ((uintptr_t)comp % 4 == 0);
    const int grid = (int)min((long long)ceil_div(((npx + 3) / 4) * L, 256), (long long)148 * 32);
    if (vec)
        k8_neighbor_merge<true><<<grid, 256, 0, (cudaStream_t)stream>>>(pred_chw, mask, ori, comp, npx, L, first_mask);
    else
        k8_neighbor_merge<false><<<grid, 256, 0, (cudaStream_t)stream>>>(pred_chw, mask, ori, comp, npx, L, first_mask);
    VV_POST_LAUNCH("k8_neighbor_merge");
    return VV_OK;
}

extern "C" int vv_apply_mask(const uint8_t *frames, const uint8_t *mask, int T, int h, int w, uint8_t *out, void *stream) {
    VV_CHECK_ARG(frames && mask && out, "vv_apply_mask: NULL pointer");
    VV_CHECK_ARG(T > 0 && h > 0 && w > 0, "vv_apply_mask: bad shape");
    const long long n_px = (long long)T * h * w;
    const bool vec = (n_px % 4 == 0) && ((uintptr_t)frames % 4 == 0) && ((uintptr_t)mask % 4 == 0) && ((uintptr_t)out % 4 == 0);
    const int grid = (int)min((long long)ceil_div((n_px + 3) / 4, 256), (long long)148 * 32);
    if (vec)
        k8_apply_mask<true><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, mask, out, n_px);
    else
        k8_apply_mask<false><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, mask, out, n_px);
    VV_POST_LAUNCH("k8_apply_mask");
    return VV_OK;
}

extern "C" int vv_mask_row_bounds(const uint32_t *mask_bits, int T, int H, int Wp, int margin, int *bounds, void *stream) {
    VV_CHECK_ARG(mask_bits && bounds && T > 0 && H > 0 && Wp > 0 && margin >= 0, "vv_mask_row_bounds: bad argument");
    k8_mask_row_bounds<<<T, 256, 0, (cudaStream_t)stream>>>(mask_bits, H, Wp, margin, bounds);
    VV_POST_LAUNCH("k8_mask_row_bounds");
    return VV_OK;
}

extern "C" int vv_swap_rb(const uint8_t *src, uint8_t *dst, size_t n_pixels, void *stream) {
    VV_CHECK_ARG(src && dst && n_pixels > 0, "vv_swap_rb: bad argument");
    const int vec_ok = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    const int grid = (int)min((long long)ceil_div(((long long)n_pixels + 15) / 16, 256), (long long)148 * 32);
    k8_swap_rb<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, (long long)n_pixels, vec_ok);
    VV_POST_LAUNCH("k8_swap_rb");
    return VV_OK;
}

// K2: resize (u8, NHWC).   Replaces cv2.resize at /root/reference/diffuerase.py:73 (:86 and
// tools.py:42 for NEAREST) and the un-vendored down-size to inference resolution that
// diffuerase.py:62-64 triggers through `max_img_size`.
//
// LINEAR reproduces OpenCV's u8 fixed-point path bit for bit (oracle/prepost.py
// model_resize_linear): per-axis tap tables with 11-bit coefficients, int32 horizontal pass
// S[sx]*a0 + S[sx+1]*a1, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
// The exact x2 down-scale that cv2 reroutes to INTER_AREA (2x2 box) is the same arithmetic
// with a0 = a1 = b0 = b1 = 1024.
//
// Tap tables are built on the device by a tiny kernel using the same IEEE double / float
// operations as OpenCV's host code, so the library keeps no host state between calls.
#include "common.cuh"

namespace vv {

struct Tap {          // one destination index along one axis
    int ofs;          // first source index (may be -1 on the y axis: rows are clipped at use)
    int w;            // w0 | (w1 << 16), 11-bit fixed point
};

// dst taps for cv2 INTER_LINEAR.  clamp_coeff = 1 on the x axis (s<0 -> s=0,f=0; s>=src-1 ->
// s=src-1,f=0), 0 on the y axis (coefficients kept, row indices clipped by the consumer).
__global__ void k2_make_linear_taps(Tap *__restrict__ taps, int dst, int src, int clamp_coeff) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dst) return;
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_coeff) {
        if (s < 0) s = 0, f = 0.f;
        if (s >= src - 1) s = src - 1, f = 0.f;
    }
    const int w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    const int w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    taps[d].ofs = s;
    taps[d].w = (w0 & 0xffff) | (w1 << 16);
}

__global__ void k2_make_nearest_taps(int *__restrict__ ofs, int dst, int src) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dst) return;
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    ofs[d] = min((int)floor(__dmul_rn((double)d, scale)), src - 1);
}

__device__ __forceinline__ int vlin(int b0, int b1, int h0, int h1) {
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

// ------------------------------------------------------------------ generic LINEAR
// One thread = 4 consecutive destination pixels of one row (all channels).
template <int C>
__global__ void __launch_bounds__(256)
    k2_resize_linear_generic(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const Tap *__restrict__ xt,
                             const Tap *__restrict__ yt, int H, int W, int h, int w, long long T, int words_ok) {
    const int groups = (w + 3) >> 2;
    const long long total = T * h * groups;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % h);
        const long long t = q / h;
        const Tap ty = yt[y];
        const int b0 = (short)(ty.w & 0xffff), b1 = ty.w >> 16;
        const int y0 = min(max(ty.ofs, 0), H - 1), y1 = min(max(ty.ofs + 1, 0), H - 1);
        const uint8_t *r0 = src + (t * H + y0) * (long long)W * C;
        const uint8_t *r1 = src + (t * H + y1) * (long long)W * C;
        uint8_t px[4 * C];
        const int x0 = g * 4, n = min(4, w - x0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < n) {
                const Tap tx = xt[x0 + i];
                const int a0 = (short)(tx.w & 0xffff), a1 = tx.w >> 16;
                const int s0 = tx.ofs * C, s1 = min(tx.ofs + 1, W - 1) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int h0 = __ldg(r0 + s0 + c) * a0 + __ldg(r0 + s1 + c) * a1;
                    const int h1 = __ldg(r1 + s0 + c) * a0 + __ldg(r1 + s1 + c) * a1;
                    px[i * C + c] = (uint8_t)vlin(b0, b1, h0, h1);
                }
            }
        }
        uint8_t *o = dst + ((t * h + y) * (long long)w + x0) * C;
        if (words_ok && n == 4) {
            uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
#pragma unroll
            for (int k = 0; k < C; ++k)
                o32[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
        } else {
            for (int k = 0; k < n * C; ++k) o[k] = px[k];
        }
    }
}

// ------------------------------------------------------------------ LINEAR, W == R*w, R = 2 or 4 (RGB)
// OpenCV's horizontal taps are then the pair (2x, 2x+1) resp. (4x+1, 4x+2) with equal weights (1024, 1024) - the
// sample position R*(x + 0.5) - 0.5 lies exactly between them; the vertical axis stays table driven, so this covers
// 1080p -> 960x540 (box), 1080p -> 960x536 and 4K -> 960x536 alike.
// One thread = 32 source pixels of two source rows (2 x 96 B, all 128-bit loads) -> 32/R destination pixels.  For
// R = 4 the generic gather kernel touched the same sectors pixel by pixel (77 % of the formula bytes); only two of
// every ~four source rows are read at all, which is why this path can exceed the formula roofline.
template <int R>
__global__ void __launch_bounds__(256)
    k2_resize_linear_ratio_rgb(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const Tap *__restrict__ yt,
                               int H, int W, int h, int w, long long T) {
    constexpr int NPX = 32 / R, NW = 3 * NPX / 4;        // destination pixels / words per thread (12 or 6)
    const int groups = w / NPX;
    const long long total = T * h * groups;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % h);
        const long long t = q / h;
        const Tap ty = yt[y];
        const int b0 = (short)(ty.w & 0xffff), b1 = ty.w >> 16;
        const int y0 = min(max(ty.ofs, 0), H - 1), y1 = min(max(ty.ofs + 1, 0), H - 1);
        const uint8_t *r0 = src + ((t * H + y0) * (long long)W + g * 32) * 3;
        const uint8_t *r1 = src + ((t * H + y1) * (long long)W + g * 32) * 3;
        uint32_t a[24], b[24];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint4 va = ldg128(r0 + 16 * k), vb = ldg128(r1 + 16 * k);
            a[4 * k] = va.x, a[4 * k + 1] = va.y, a[4 * k + 2] = va.z, a[4 * k + 3] = va.w;
            b[4 * k] = vb.x, b[4 * k + 1] = vb.y, b[4 * k + 2] = vb.z, b[4 * k + 3] = vb.w;
        }
        uint32_t o[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) o[k] = 0;
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int i0 = 3 * (R * j + R / 2 - 1) + c, i1 = i0 + 3;      // source byte indices in the 96-byte span
                const int h0 = (int)(byte_of(a[i0 >> 2], i0 & 3) + byte_of(a[i1 >> 2], i1 & 3)) << 10;
                const int h1 = (int)(byte_of(b[i0 >> 2], i0 & 3) + byte_of(b[i1 >> 2], i1 & 3)) << 10;
                const int ob = 3 * j + c;
                o[ob >> 2] |= (uint32_t)vlin(b0, b1, h0, h1) << (8 * (ob & 3));
            }
        }
        uint8_t *op = dst + ((t * h + y) * (long long)w + g * NPX) * 3;
        if (R == 2) {
            stg128_stream(op, make_uint4(o[0], o[1], o[2], o[3]));
            stg128_stream(op + 16, make_uint4(o[4], o[5], o[6], o[7]));
            stg128_stream(op + 32, make_uint4(o[NW - 4], o[NW - 3], o[NW - 2], o[NW - 1]));
        } else {
#pragma unroll
            for (int k = 0; k < NW; k += 2) *reinterpret_cast<uint2 *>(op + 4 * k) = make_uint2(o[k], o[k + 1]);
        }
    }
}

// ------------------------------------------------------------------ NEAREST
template <int C>
__global__ void __launch_bounds__(256)
    k2_resize_nearest(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const int *__restrict__ xo,
                      const int *__restrict__ yo, int H, int W, int h, int w, long long T) {
    const long long total = T * h * (long long)w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w);
        const long long q = idx / w;
        const int y = (int)(q % h);
        const long long t = q / h;
        const uint8_t *s = src + ((t * H + yo[y]) * (long long)W + xo[x]) * C;
        uint8_t *o = dst + idx * C;
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = __ldg(s + c);
    }
}

// NEAREST for single-channel images (masks): 16 destination pixels per thread, one 128-bit store.
__global__ void __launch_bounds__(256)
    k2_resize_nearest_c1_x16(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const int *__restrict__ xo,
                             const int *__restrict__ yo, int H, int W, int h, int w, long long T) {
    const int groups = w >> 4;
    const long long total = T * h * (long long)groups;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % h);
        const long long t = q / h;
        const uint8_t *s = src + (t * H + yo[y]) * (long long)W;
        const int4 *xq = reinterpret_cast<const int4 *>(xo + g * 16);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int4 xi = __ldg(xq + k);
            o[k] = __ldg(s + xi.x) | (__ldg(s + xi.y) << 8) | (__ldg(s + xi.z) << 16) | ((uint32_t)__ldg(s + xi.w) << 24);
        }
        stg128_stream(dst + (t * h + y) * (long long)w + g * 16, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// NEAREST for single-channel images at an exact ratio W == R * w (R = 2 or 4): OpenCV's source column of pixel x is
// exactly R * x (floor(x * (1 / (w / W))) with an exactly representable scale), so 16 destination pixels are every
// R-th byte of ONE aligned span of 16 R source bytes: 128-bit loads and byte permutes instead of 16 byte gathers
// through the column table (1080p -> 540p mask: 52 % of the HBM roofline with the gather kernel).
template <int R>
__global__ void __launch_bounds__(256)
    k2_resize_nearest_c1_ratio(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const int *__restrict__ yo, int H,
                               int W, int h, int w, long long T) {
    const int groups = w >> 4;
    const long long total = T * h * (long long)groups;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % h);
        const long long t = q / h;
        const uint8_t *s = src + (t * H + yo[y]) * (long long)W + (long long)g * 16 * R;
        uint32_t o[4];
        if (R == 2) {
            const uint4 a = ldg128(s), b = ldg128(s + 16);
            o[0] = __byte_perm(a.x, a.y, 0x6420), o[1] = __byte_perm(a.z, a.w, 0x6420);
            o[2] = __byte_perm(b.x, b.y, 0x6420), o[3] = __byte_perm(b.z, b.w, 0x6420);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 a = ldg128(s + 16 * k);
                o[k] = __byte_perm(__byte_perm(a.x, a.y, 0x0040), __byte_perm(a.z, a.w, 0x0040), 0x5410);
            }
        }
        stg128_stream(dst + (t * h + y) * (long long)w + g * 16, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

}  // namespace vv

using namespace vv;

extern "C" size_t vv_resize_workspace_bytes(int h, int w) {
    if (h <= 0 || w <= 0) return 0;
    return align_up((size_t)w * sizeof(Tap), 256) + align_up((size_t)h * sizeof(Tap), 256);
}

extern "C" int vv_inference_size(int H0, int W0, int max_img_size, int *h, int *w) {
    VV_CHECK_ARG(h && w && H0 > 0 && W0 > 0 && max_img_size > 0, "vv_inference_size: bad argument");
    int ww = W0, hh = H0;
    const int mx = H0 > W0 ? H0 : W0;
    if (mx > max_img_size) {
        const double r = (double)mx / (double)max_img_size;
        ww = (int)((double)W0 / r);
        hh = (int)((double)H0 / r);
    }
    *w = ww - ww % 8;
    *h = hh - hh % 8;
    return VV_OK;
}

namespace vv {
// Shared with K3: builds x/y LINEAR tap tables for (H,W) -> (h,w) into `workspace`.
// which: 1 = the x table, 2 = the y table, 3 = both (a caller that uses one axis in closed form skips that launch)
int build_linear_taps(void *workspace, int H, int W, int h, int w, const Tap **xt, const Tap **yt, cudaStream_t st, int which) {
    Tap *x = (Tap *)workspace;
    Tap *y = (Tap *)((uint8_t *)workspace + align_up((size_t)w * sizeof(Tap), 256));
    if (which & 1) {
        k2_make_linear_taps<<<ceil_div(w, 256), 256, 0, st>>>(x, w, W, 1);
        VV_POST_LAUNCH("k2_make_linear_taps(x)");
    }
    if (which & 2) {
        k2_make_linear_taps<<<ceil_div(h, 256), 256, 0, st>>>(y, h, H, 0);
        VV_POST_LAUNCH("k2_make_linear_taps(y)");
    }
    *xt = x, *yt = y;
    return VV_OK;
}
}  // namespace vv

extern "C" int vv_resize(const uint8_t *src, int T, int H, int W, int C, uint8_t *dst, int h, int w, int interp,
                         void *workspace, size_t workspace_bytes, void *stream) {
    VV_CHECK_ARG(src && dst && workspace, "vv_resize: NULL pointer");
    VV_CHECK_ARG(T > 0 && H > 0 && W > 0 && h > 0 && w > 0, "vv_resize: bad shape");
    VV_CHECK_ARG(C == 1 || C == 3 || C == 4, "vv_resize: C must be 1, 3 or 4 (got %d)", C);
    VV_CHECK_ARG(interp == VV_INTER_LINEAR || interp == VV_INTER_NEAREST, "vv_resize: unknown interpolation %d", interp);
    VV_CHECK_ARG(workspace_bytes >= vv_resize_workspace_bytes(h, w), "vv_resize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (H == h && W == w) {   // the reference only resizes when the shape differs (diffuerase.py:72)
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)T * H * W * C, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync");
        return VV_OK;
    }
    const int max_grid = 148 * 32;
    if (interp == VV_INTER_NEAREST) {
        int *xo = (int *)workspace;
        int *yo = (int *)((uint8_t *)workspace + align_up((size_t)w * sizeof(Tap), 256));
        k2_make_nearest_taps<<<ceil_div(w, 256), 256, 0, st>>>(xo, w, W);
        VV_POST_LAUNCH("k2_make_nearest_taps(x)");
        k2_make_nearest_taps<<<ceil_div(h, 256), 256, 0, st>>>(yo, h, H);
        VV_POST_LAUNCH("k2_make_nearest_taps(y)");
        const int grid = (int)min((long long)ceil_div((long long)T * h * w, 256), (long long)max_grid);
        const bool c1v = C == 1 && w % 16 == 0 && (uintptr_t)dst % 16 == 0;
        const int g16 = (int)min((long long)ceil_div((long long)T * h * (w / 16 + 1), 256), (long long)max_grid);
        if (c1v && W == 2 * w && (uintptr_t)src % 16 == 0) {
            k2_resize_nearest_c1_ratio<2><<<g16, 256, 0, st>>>(src, dst, yo, H, W, h, w, T);
        } else if (c1v && W == 4 * w && (uintptr_t)src % 16 == 0) {
            k2_resize_nearest_c1_ratio<4><<<g16, 256, 0, st>>>(src, dst, yo, H, W, h, w, T);
        } else if (c1v) {
            k2_resize_nearest_c1_x16<<<g16, 256, 0, st>>>(src, dst, xo, yo, H, W, h, w, T);
        } else if (C == 1)
            k2_resize_nearest<1><<<grid, 256, 0, st>>>(src, dst, xo, yo, H, W, h, w, T);
        else if (C == 3)
            k2_resize_nearest<3><<<grid, 256, 0, st>>>(src, dst, xo, yo, H, W, h, w, T);
        else
            k2_resize_nearest<4><<<grid, 256, 0, st>>>(src, dst, xo, yo, H, W, h, w, T);
        VV_POST_LAUNCH("k2_resize_nearest");
        return VV_OK;
    }
    const Tap *xt, *yt;
    int rc = build_linear_taps(workspace, H, W, h, w, &xt, &yt, st, 3);
    if (rc) return rc;
    if (C == 3 && W == 2 * w && w % 16 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
        const int grid = (int)min((long long)ceil_div((long long)T * h * (w / 16), 256), (long long)max_grid);
        k2_resize_linear_ratio_rgb<2><<<grid, 256, 0, st>>>(src, dst, yt, H, W, h, w, T);
        VV_POST_LAUNCH("k2_resize_linear_ratio_rgb<2>");
        return VV_OK;
    }
    if (C == 3 && W == 4 * w && w % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 8 == 0) {
        const int grid = (int)min((long long)ceil_div((long long)T * h * (w / 8), 256), (long long)max_grid);
        k2_resize_linear_ratio_rgb<4><<<grid, 256, 0, st>>>(src, dst, yt, H, W, h, w, T);
        VV_POST_LAUNCH("k2_resize_linear_ratio_rgb<4>");
        return VV_OK;
    }
    const int words_ok = ((w * C) % 4 == 0) && ((uintptr_t)dst % 4 == 0);
    const int grid = (int)min((long long)ceil_div((long long)T * h * ((w + 3) / 4), 256), (long long)max_grid);
    if (C == 1)
        k2_resize_linear_generic<1><<<grid, 256, 0, st>>>(src, dst, xt, yt, H, W, h, w, T, words_ok);
    else if (C == 3)
        k2_resize_linear_generic<3><<<grid, 256, 0, st>>>(src, dst, xt, yt, H, W, h, w, T, words_ok);
    else
        k2_resize_linear_generic<4><<<grid, 256, 0, st>>>(src, dst, xt, yt, H, W, h, w, T, words_ok);
    VV_POST_LAUNCH("k2_resize_linear_generic");
    return VV_OK;
}

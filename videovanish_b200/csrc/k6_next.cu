// "Next" rows of SURVEY.md section 8f: the pixel glue immediately either side of the hot path.
//
// N3  SAM2 mask colour painter.  Replaces /root/reference/sam2_masker.py:151-175: a black canvas per
//     frame, every object's mask (logits > 0, NEAREST-resized to the frame if needed, :167) painted in
//     ascending object order so that the highest id wins (:159-173).
// N2  K4 output -> the float tensors the ProPainter network consumes: `to_tensors()*2-1` normalisation
//     of the propagated pixels (0.0 where the state says so) in CHW order + the updated hole mask
//     [recalled-upstream propainter/inference.py; decode = oracle/propagation.py decode_state].
#include <string.h>

#include "common.cuh"

namespace vv {

constexpr int K6_MAX_OBJECTS = 255;
struct PaintColors {
    uint8_t rgb[K6_MAX_OBJECTS * 3 + 3];
};

// One thread = 4 consecutive canvas pixels (12 bytes = 3 words).  Objects are tested from the last
// (highest priority) to the first; the first hit decides the colour.
template <typename MaskT>
__global__ void __launch_bounds__(256)
    k6_paint_masks(const MaskT *__restrict__ masks, int K, int mh, int mw, uint8_t *__restrict__ out, int H0, int W0,
                   long long T, const int *__restrict__ xo, const int *__restrict__ yo, int words_ok,
                   const __grid_constant__ PaintColors colors) {
    const int groups = (W0 + 3) >> 2;
    const long long total = T * H0 * (long long)groups;
    const long long plane = (long long)mh * mw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = (int)(q % H0);
        const long long t = q / H0;
        const int x0 = g * 4, n = min(4, W0 - x0);
        const MaskT *frame = masks + t * K * plane + (long long)yo[y] * mw;
        uint32_t px[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < n) {
                const int sx = xo[x0 + i];
                int top = -1;                         // highest object whose mask is set here
                if (K <= 8) {                         // all object planes loaded at once (independent loads)
                    uint32_t hit = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < K) hit |= (uint32_t)(frame[k * plane + sx] > (MaskT)0) << k;
                    top = 31 - __clz((int)hit);       // -1 when nothing is set
                } else {
                    for (int k = K - 1; k >= 0; --k)
                        if (frame[k * plane + sx] > (MaskT)0) {
                            top = k;
                            break;
                        }
                }
                if (top >= 0) px[i] = colors.rgb[3 * top] | (colors.rgb[3 * top + 1] << 8) | (colors.rgb[3 * top + 2] << 16);
            }
        }
        uint8_t *o = out + ((t * H0 + y) * (long long)W0 + x0) * 3;
        if (words_ok && n == 4) {
            uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
            o32[0] = px[0] | (px[1] << 24);
            o32[1] = (px[1] >> 8) | (px[2] << 16);
            o32[2] = (px[2] >> 16) | (px[3] << 8);
        } else {
            for (int i = 0; i < n; ++i) o[3 * i] = (uint8_t)px[i], o[3 * i + 1] = (uint8_t)(px[i] >> 8), o[3 * i + 2] = (uint8_t)(px[i] >> 16);
        }
    }
}

// Fast path for the two horizontal ratios the GUI produces (SAM2 returns masks at the video resolution:
// W0 == mw; half-resolution predictors: W0 == 2 * mw, where OpenCV's NEAREST index is x >> 1): one thread
// = 16 consecutive canvas pixels of a row, the object planes are read with 16-byte (8-byte) loads and the
// 48 output bytes leave as three 16-byte stores.  Any vertical ratio (row table `yo`).
template <typename MaskT, bool X2>
__device__ __forceinline__ uint32_t k6_hits16(const MaskT *__restrict__ p) {   // bit i = canvas pixel i is set
    uint32_t bits = 0;
    if (sizeof(MaskT) == 4) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int j = 0; j < (X2 ? 2 : 4); ++j) {
            const float4 v = __ldg(q + j);
            bits |= ((uint32_t)(v.x > 0.f) | ((uint32_t)(v.y > 0.f) << 1) | ((uint32_t)(v.z > 0.f) << 2) |
                     ((uint32_t)(v.w > 0.f) << 3))
                    << (4 * j);
        }
    } else if (X2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        bits = nonzero_bits16(make_uint4(v.x, v.y, 0u, 0u));
    } else {
        bits = nonzero_bits16(ldg128(p));
    }
    if (X2) {             // 8 source bits -> every bit doubled
        bits = (bits | (bits << 4)) & 0x0f0fu;
        bits = (bits | (bits << 2)) & 0x3333u;
        bits = (bits | (bits << 1)) & 0x5555u;
        bits |= bits << 1;
    }
    return bits;
}

template <typename MaskT, bool X2>
__global__ void __launch_bounds__(256)
    k6_paint_masks_rows(const MaskT *__restrict__ masks, int K, int mh, int mw, uint8_t *__restrict__ out, int H0, int W0,
                        long long T, const int *__restrict__ yo, int pair_rows, const __grid_constant__ PaintColors colors) {
    // pair_rows (H0 == 2 * mh: output rows 2j and 2j+1 both come from source row j): one work item paints both, so the
    // mask planes are read and tested once per source row instead of twice
    const int groups = W0 >> 4;
    const int rows = pair_rows ? H0 >> 1 : H0;
    const long long total = T * rows * (long long)groups;
    const long long plane = (long long)mh * mw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const long long q = idx / groups;
        const int y = pair_rows ? 2 * (int)(q % rows) : (int)(q % rows);
        const long long t = q / rows;
        const MaskT *src = masks + t * K * plane + (long long)yo[y] * mw + (X2 ? g * 8 : g * 16);
        // ascending object order, later objects overwrite (sam2_masker.py:159-173); up to 8 planes in flight
        uint32_t px[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) px[i] = 0;
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t hits[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) hits[k] = (k0 + k < K) ? k6_hits16<MaskT, X2>(src + (k0 + k) * plane) : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k0 + k < K && hits[k]) {
                    const uint8_t *c = colors.rgb + 3 * (k0 + k);
                    const uint32_t col = c[0] | (c[1] << 8) | (c[2] << 16);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((hits[k] >> i) & 1u) px[i] = col;
                }
            }
        }
        uint4 *o = reinterpret_cast<uint4 *>(out + ((t * H0 + y) * (long long)W0 + g * 16) * 3);
        uint32_t ow[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {          // 4 pixels -> 3 words
            const uint32_t a = px[4 * j], b = px[4 * j + 1], c = px[4 * j + 2], d = px[4 * j + 3];
            ow[3 * j] = a | (b << 24), ow[3 * j + 1] = (b >> 8) | (c << 16), ow[3 * j + 2] = (c >> 16) | (d << 8);
        }
        stg128_stream(o, make_uint4(ow[0], ow[1], ow[2], ow[3]));
        stg128_stream(o + 1, make_uint4(ow[4], ow[5], ow[6], ow[7]));
        stg128_stream(o + 2, make_uint4(ow[8], ow[9], ow[10], ow[11]));
        if (pair_rows) {
            uint4 *o2 = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(o) + (long long)W0 * 3);
            stg128_stream(o2, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            stg128_stream(o2 + 1, make_uint4(ow[4], ow[5], ow[6], ow[7]));
            stg128_stream(o2 + 2, make_uint4(ow[8], ow[9], ow[10], ow[11]));
        }
    }
}

__global__ void k6_make_nearest_taps(int *__restrict__ ofs, int dst, int src) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dst) return;
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    ofs[d] = min((int)floor(__dmul_rn((double)d, scale)), src - 1);
}

// packed [n_frames,h,w] -> f32 [n_frames,3,h,w] in [-1,1] (+ f32 hole mask [n_frames,h,w]); 4 pixels per thread.
__global__ void __launch_bounds__(256)
    k6_state_to_float(const uint32_t *__restrict__ packed, long long n_frames, long long npx, float *__restrict__ rgb,
                      float *__restrict__ hole, int vec_ok) {
    const long long groups = (npx + 3) >> 2;
    const long long total = n_frames * groups;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long t = idx / groups, p0 = (idx - t * groups) * 4;
        const int n = (int)min(4LL, npx - p0);
        uint32_t v[4] = {0, 0, 0, 0};
        if (vec_ok) {
            const uint4 u = *reinterpret_cast<const uint4 *>(packed + t * npx + p0);
            v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w;
        } else {
            for (int i = 0; i < n; ++i) v[i] = packed[t * npx + p0 + i];
        }
        float c[3][4], m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool zero = v[i] & (2u << 24);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                // (u8 / 255) * 2 - 1 in fp32, like to_tensors() followed by *2-1
                const float f = __fsub_rn(__fmul_rn(__fdiv_rn((float)byte_of(v[i], ch), 255.f), 2.f), 1.f);
                c[ch][i] = zero ? 0.f : f;
            }
            m[i] = (v[i] & (1u << 24)) ? 1.f : 0.f;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float *o = rgb + (t * 3 + ch) * npx + p0;
            if (vec_ok) {
                *reinterpret_cast<float4 *>(o) = make_float4(c[ch][0], c[ch][1], c[ch][2], c[ch][3]);
            } else {
                for (int i = 0; i < n; ++i) o[i] = c[ch][i];
            }
        }
        if (hole) {
            float *o = hole + t * npx + p0;
            if (vec_ok) {
                *reinterpret_cast<float4 *>(o) = make_float4(m[0], m[1], m[2], m[3]);
            } else {
                for (int i = 0; i < n; ++i) o[i] = m[i];
            }
        }
    }
}

}  // namespace vv

using namespace vv;

extern "C" size_t vv_paint_masks_workspace_bytes(int H0, int W0) {
    if (H0 <= 0 || W0 <= 0) return 0;
    return align_up((size_t)W0 * 4, 256) + align_up((size_t)H0 * 4, 256);
}

extern "C" int vv_paint_masks(const void *masks, int mask_is_f32, int T, int K, int mh, int mw, const uint8_t *colors_host,
                              uint8_t *out, int H0, int W0, void *workspace, size_t workspace_bytes, void *stream) {
    VV_CHECK_ARG(masks && colors_host && out && workspace, "vv_paint_masks: NULL pointer");
    VV_CHECK_ARG(T > 0 && K > 0 && mh > 0 && mw > 0 && H0 > 0 && W0 > 0, "vv_paint_masks: bad shape");
    VV_CHECK_ARG(K <= K6_MAX_OBJECTS, "vv_paint_masks: at most %d objects (got %d)", K6_MAX_OBJECTS, K);
    VV_CHECK_ARG(workspace_bytes >= vv_paint_masks_workspace_bytes(H0, W0), "vv_paint_masks: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int *xo = (int *)workspace;
    int *yo = (int *)((uint8_t *)workspace + align_up((size_t)W0 * 4, 256));
    k6_make_nearest_taps<<<ceil_div(W0, 256), 256, 0, st>>>(xo, W0, mw);
    VV_POST_LAUNCH("k6_make_nearest_taps(x)");
    k6_make_nearest_taps<<<ceil_div(H0, 256), 256, 0, st>>>(yo, H0, mh);
    VV_POST_LAUNCH("k6_make_nearest_taps(y)");
    PaintColors pc;
    memset(&pc, 0, sizeof(pc));
    memcpy(pc.rgb, colors_host, (size_t)K * 3);
    const int words_ok = ((W0 * 3) % 4 == 0) && ((uintptr_t)out % 4 == 0);
    const long long total = (long long)T * H0 * ((W0 + 3) / 4);
    const int grid = (int)min((long long)ceil_div(total, 256), (long long)148 * 32);
    // row-vectorised fast paths: identical width or exact x2, 16-byte aligned rows on both sides
    const size_t esz = mask_is_f32 ? 4 : 1;
    const bool same_w = W0 == mw, twice_w = W0 == 2 * mw;
    const int src_per_16 = same_w ? 16 : 8;
    const bool rows_ok = (same_w || twice_w) && (W0 % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                         ((uintptr_t)masks % (mask_is_f32 ? 16 : src_per_16) == 0) &&
                         (((size_t)mw * esz) % (mask_is_f32 ? 16 : src_per_16) == 0);
    if (rows_ok) {
        const int pair_rows = (H0 == 2 * mh) ? 1 : 0;        // NEAREST row of y is floor(y / 2) exactly
        const long long tot16 = (long long)T * (pair_rows ? H0 / 2 : H0) * (W0 / 16);
        const int g16 = (int)min((long long)ceil_div(tot16, 256), (long long)148 * 32);
#define VV_K6_ROWS(MT, X2) \
    k6_paint_masks_rows<MT, X2><<<g16, 256, 0, st>>>((const MT *)masks, K, mh, mw, out, H0, W0, T, yo, pair_rows, pc)
        if (mask_is_f32 && twice_w)
            VV_K6_ROWS(float, true);
        else if (mask_is_f32)
            VV_K6_ROWS(float, false);
        else if (twice_w)
            VV_K6_ROWS(uint8_t, true);
        else
            VV_K6_ROWS(uint8_t, false);
#undef VV_K6_ROWS
        VV_POST_LAUNCH("k6_paint_masks_rows");
        return VV_OK;
    }
    if (mask_is_f32)
        k6_paint_masks<float><<<grid, 256, 0, st>>>((const float *)masks, K, mh, mw, out, H0, W0, T, xo, yo, words_ok, pc);
    else
        k6_paint_masks<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)masks, K, mh, mw, out, H0, W0, T, xo, yo, words_ok, pc);
    VV_POST_LAUNCH("k6_paint_masks");
    return VV_OK;
}

extern "C" int vv_propagate_to_float(const uint32_t *packed, int n_frames, int h, int w, float *rgb_chw, float *hole_mask,
                                     void *stream) {
    VV_CHECK_ARG(packed && rgb_chw, "vv_propagate_to_float: NULL pointer");
    VV_CHECK_ARG(n_frames > 0 && h > 0 && w > 0, "vv_propagate_to_float: bad shape");
    const long long npx = (long long)h * w;
    const int vec_ok = (npx % 4 == 0) && ((uintptr_t)packed % 16 == 0) && ((uintptr_t)rgb_chw % 16 == 0) &&
                       (!hole_mask || (uintptr_t)hole_mask % 16 == 0);
    const long long total = (long long)n_frames * ((npx + 3) / 4);
    const int grid = (int)min((long long)ceil_div(total, 256), (long long)148 * 32);
    k6_state_to_float<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, n_frames, npx, rgb_chw, hole_mask, vec_ok);
    VV_POST_LAUNCH("k6_state_to_float");
    return VV_OK;
}

// K5: chunk-overlap feather blend.  The reference claims the feature (README.md:18) but ships no
// code for it (README.md:76 lists it as a TODO), so the spec is builder-defined (SURVEY row A11,
// oracle/chunk_blend.py) and mirrors the composite arithmetic of diffuerase.py:112:
//   for overlap frame k in [0, O):  w = f32(k+1) / f32(O+1)
//   out = u8(clip(rint(f32((1-w)*A) + f32(w*B))))     A = earlier chunk's tail, B = later chunk's head
// Pure streaming: 2 x 16 B in, 16 B out per thread-iteration, no reuse.  `B` may be a peer-GPU
// pointer (cudaIpcOpenMemHandle / P2P mapping): the loads then travel over NVLink inside this
// kernel, so the halo transfer overlaps the blend instead of preceding it.
#include "common.cuh"

namespace vv {

// w in (0,1) and both inputs in [0,255]: the rounded sum stays in [0,255], so the clip of the spec is a no-op.
__device__ __forceinline__ uint32_t blend4(uint32_t a, uint32_t b, float w, float nw) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float v = __fadd_rn(__fmul_rn(nw, u8_to_float(a, i)), __fmul_rn(w, u8_to_float(b, i)));
        r[i] = rint_u8_bits(v);            // only the low byte is used below
    }
    return __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
}

__global__ void __launch_bounds__(256)
    k5_chunk_blend(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, uint8_t *__restrict__ out,
                   long long frame_bytes, int k0, int O_total, int vec) {
    const int k = blockIdx.y;
    const float w = __fdiv_rn((float)(k0 + k + 1), (float)(O_total + 1));
    const float nw = __fsub_rn(1.f, w);
    const uint8_t *a = A + k * frame_bytes, *b = B + k * frame_bytes;
    uint8_t *o = out + k * frame_bytes;
    const long long n16 = vec ? frame_bytes / 16 : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + stride < n16; i += 2 * stride) {      // two independent 16-byte columns in flight per thread
        const uint4 va = ldg128(a + 16 * i), vb = ldg128(b + 16 * i);
        const uint4 vc = ldg128(a + 16 * (i + stride)), vd = ldg128(b + 16 * (i + stride));
        stg128_stream(o + 16 * i, make_uint4(blend4(va.x, vb.x, w, nw), blend4(va.y, vb.y, w, nw),
                                             blend4(va.z, vb.z, w, nw), blend4(va.w, vb.w, w, nw)));
        stg128_stream(o + 16 * (i + stride), make_uint4(blend4(vc.x, vd.x, w, nw), blend4(vc.y, vd.y, w, nw),
                                                        blend4(vc.z, vd.z, w, nw), blend4(vc.w, vd.w, w, nw)));
    }
    for (; i < n16; i += stride) {
        const uint4 va = ldg128(a + 16 * i), vb = ldg128(b + 16 * i);
        stg128_stream(o + 16 * i, make_uint4(blend4(va.x, vb.x, w, nw), blend4(va.y, vb.y, w, nw),
                                             blend4(va.z, vb.z, w, nw), blend4(va.w, vb.w, w, nw)));
    }
    for (long long i = n16 * 16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < frame_bytes; i += stride) {
        const float v = __fadd_rn(__fmul_rn(nw, (float)a[i]), __fmul_rn(w, (float)b[i]));
        o[i] = (uint8_t)min(max(__float2int_rn(v), 0), 255);
    }
}

}  // namespace vv

using namespace vv;

extern "C" int vv_chunk_blend(const uint8_t *A, const uint8_t *B, int O, size_t frame_bytes, int k0, int O_total,
                              uint8_t *out, void *stream) {
    VV_CHECK_ARG(A && B && out, "vv_chunk_blend: NULL pointer");
    VV_CHECK_ARG(O > 0 && frame_bytes > 0 && k0 >= 0 && O_total >= k0 + O, "vv_chunk_blend: bad overlap range");
    VV_CHECK_ARG(O <= 65535, "vv_chunk_blend: at most 65535 overlap frames per call");
    const int vec = (frame_bytes % 16 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0);
    const long long per_frame = vec ? (long long)frame_bytes / 16 : (long long)frame_bytes;
    // grid.x sized so that one launch fills the machine a few times over (148 SMs x 8 CTAs)
    long long gx = ceil_div(per_frame, 256 * 4);
    const long long cap = ceil_div(148 * 16, O);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    k5_chunk_blend<<<dim3((unsigned)gx, (unsigned)O), 256, 0, (cudaStream_t)stream>>>(A, B, out, (long long)frame_bytes,
                                                                                     k0, O_total, vec);
    VV_POST_LAUNCH("k5_chunk_blend");
    return VV_OK;
}

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

// ---- rank-boundary halo blend with a device-side handshake ------------------------------------------
// One process per GPU; consecutive ranks share `overlap` frames.  Each rank blends, in place, the half of
// each boundary it owns and reads the neighbour's ORIGINAL frames of that half in place over NVLink (CUDA-IPC
// mapping).  No host barrier: the kernels of neighbouring ranks synchronise through three flags per rank that
// live in IPC-mapped device memory and carry a monotonically increasing epoch:
//   ready       set by this rank's kernel at its start: everything this rank enqueued before (K3 ...) is done
//   consumed[2] set by the neighbours' kernels when they have finished reading this rank's frames
// A block polls the peer's `ready` (ld.acquire.sys) before it touches peer memory; the last block of the grid
// releases the neighbours' `consumed` flags and then waits for its own, so that the kernel - and with it the
// stream - only moves on when nobody reads this rank's buffer any more.  Spins are bounded (a lost peer sets
// the error flag instead of hanging the GPU).
constexpr int HF_READY = 0, HF_CONSUMED_BY_PREV = 1, HF_CONSUMED_BY_NEXT = 2, HF_DONE = 3, HF_ERROR = 4;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// true when the flag reached `epoch` (wrap-safe), false after ~2 s of polling
__device__ __forceinline__ bool wait_flag(const uint32_t *p, uint32_t epoch) {
    for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
        if ((int32_t)(ld_acquire_sys(p) - epoch) >= 0) return true;
        __nanosleep(400);
    }
    return false;
}

struct HaloJob {
    const uint8_t *A, *B;     // earlier chunk's tail frames, later chunk's head frames (one of them is peer memory)
    uint8_t *out;
    int n, k0;                // frames of this job, index of its first frame inside the overlap
    const uint32_t *peer_ready;   // flags of the rank whose memory this job reads
};

__global__ void __launch_bounds__(256)
    k5_halo_blend(const __grid_constant__ HaloJob next_job, const __grid_constant__ HaloJob prev_job, long long frame_bytes,
                  int O_total, uint32_t *my_flags, uint32_t *next_flags, uint32_t *prev_flags, uint32_t epoch) {
    __shared__ int s_ok;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(my_flags + HF_READY, epoch);
    }
    const bool is_next = (int)blockIdx.y < next_job.n;
    const HaloJob &job = is_next ? next_job : prev_job;
    const int k = is_next ? (int)blockIdx.y : (int)blockIdx.y - next_job.n;
    if (threadIdx.x == 0) {
        s_ok = wait_flag(job.peer_ready, epoch) ? 1 : 0;
        if (!s_ok) my_flags[HF_ERROR] = 1;
    }
    __syncthreads();
    if (s_ok) {
        const float w = __fdiv_rn((float)(job.k0 + k + 1), (float)(O_total + 1));
        const float nw = __fsub_rn(1.f, w);
        const uint4 *a = reinterpret_cast<const uint4 *>(job.A + k * frame_bytes);
        const uint4 *b = reinterpret_cast<const uint4 *>(job.B + k * frame_bytes);
        uint8_t *o = job.out + k * frame_bytes;
        const long long n16 = frame_bytes / 16, stride = (long long)gridDim.x * blockDim.x;
        long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        // plain (coherent) loads: one operand is another GPU's memory that was written during this epoch.  Four
        // independent pairs per trip: NVLink reads take microseconds, and with the small grid of the overlapped mode
        // (the exchange runs underneath K3 and must not take its SM slots) bytes in flight are what sets the rate.
        for (; i + 3 * stride < n16; i += 4 * stride) {
            uint4 va[4], vb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) va[j] = a[i + j * stride], vb[j] = b[i + j * stride];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                stg128_stream(o + 16 * (i + j * stride), make_uint4(blend4(va[j].x, vb[j].x, w, nw), blend4(va[j].y, vb[j].y, w, nw),
                                                                    blend4(va[j].z, vb[j].z, w, nw), blend4(va[j].w, vb[j].w, w, nw)));
        }
        for (; i < n16; i += stride) {
            const uint4 va = a[i], vb = b[i];
            stg128_stream(o + 16 * i, make_uint4(blend4(va.x, vb.x, w, nw), blend4(va.y, vb.y, w, nw),
                                                 blend4(va.z, vb.z, w, nw), blend4(va.w, vb.w, w, nw)));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t total = gridDim.x * gridDim.y;
        if (atomicAdd(my_flags + HF_DONE, 1u) == total - 1) {          // last block of this rank's grid
            my_flags[HF_DONE] = 0;
            __threadfence_system();
            if (next_flags) st_release_sys(next_flags + HF_CONSUMED_BY_PREV, epoch);
            if (prev_flags) st_release_sys(prev_flags + HF_CONSUMED_BY_NEXT, epoch);
            bool ok = true;
            if (prev_flags) ok &= wait_flag(my_flags + HF_CONSUMED_BY_PREV, epoch);
            if (next_flags) ok &= wait_flag(my_flags + HF_CONSUMED_BY_NEXT, epoch);
            if (!ok) my_flags[HF_ERROR] = 1;
        }
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

extern "C" int vv_halo_blend(uint8_t *out, int T, size_t frame_bytes, int overlap, const uint8_t *next_head,
                             const uint8_t *prev_tail, uint32_t *my_flags, uint32_t *next_flags, uint32_t *prev_flags,
                             uint32_t epoch, void *stream) {
    VV_CHECK_ARG(out && my_flags, "vv_halo_blend: NULL pointer");
    VV_CHECK_ARG(overlap >= 2, "vv_halo_blend: overlap must be at least 2 (each side blends half of it)");
    VV_CHECK_ARG(T >= 2 * overlap, "vv_halo_blend: a rank needs at least 2 x overlap frames (got %d for overlap %d)", T, overlap);
    VV_CHECK_ARG((next_head == nullptr) == (next_flags == nullptr) && (prev_tail == nullptr) == (prev_flags == nullptr),
                 "vv_halo_blend: peer frames and peer flags go together");
    VV_CHECK_ARG(frame_bytes % 16 == 0 && (uintptr_t)out % 16 == 0 && (!next_head || (uintptr_t)next_head % 16 == 0) &&
                     (!prev_tail || (uintptr_t)prev_tail % 16 == 0),
                 "vv_halo_blend: frames must be 16-byte aligned");
    const int half = overlap / 2;
    HaloJob nj = {}, pj = {};
    if (next_head && half > 0) {           // overlap indices [0, half): my tail (A, in place) x next rank's head (B, peer)
        nj.A = out + (size_t)(T - overlap) * frame_bytes, nj.B = next_head, nj.out = out + (size_t)(T - overlap) * frame_bytes;
        nj.n = half, nj.k0 = 0, nj.peer_ready = next_flags + HF_READY;
    }
    if (prev_tail && overlap - half > 0) { // [half, overlap): prev rank's tail (A, peer) x my head (B, in place)
        pj.A = prev_tail, pj.B = out + (size_t)half * frame_bytes, pj.out = out + (size_t)half * frame_bytes;
        pj.n = overlap - half, pj.k0 = half, pj.peer_ready = prev_flags + HF_READY;
    }
    const int frames = nj.n + pj.n;
    VV_CHECK_ARG(frames > 0, "vv_halo_blend: no neighbour given");
    // CTA budget: the whole machine a few times over when the exchange runs alone, a few dozen CTAs when it runs on a side
    // stream underneath another kernel (option k5_halo_ctas, set by chunking.produce_and_blend_boundaries)
    const int budget = max(frames, get_option(OPT_K5_HALO_CTAS));
    const long long gx = max(1LL, min((long long)ceil_div((long long)frame_bytes / 16, 256 * 4),
                                      (long long)(budget / frames)));
    dim3 grid((unsigned)gx, (unsigned)frames);
    k5_halo_blend<<<grid, 256, 0, (cudaStream_t)stream>>>(nj, pj, (long long)frame_bytes, overlap, my_flags, next_flags,
                                                         prev_flags, epoch);
    VV_POST_LAUNCH("k5_halo_blend");
    return VV_OK;
}

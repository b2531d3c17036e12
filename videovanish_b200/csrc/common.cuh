// Shared helpers for the sm_100a kernels behind include/vvb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/vvb200.h"

namespace vv {

void set_error(const char *fmt, ...);

// Tuning switches (vv_set_option): variants kept side by side so that one GPU run can A/B them.
enum Option { OPT_K1B_EXACT = 0, OPT_K3_NT, OPT_K3_TMA, OPT_K3_TMA_ROWS, OPT_K3_TMA_THREADS, OPT_K4_PDL, OPT_K4_NPT, OPT_K3_BITS, OPT_K3_X2, OPT_K4_PACK_CTAS, OPT_K4_PACK_OCC, OPT_K4_LEAN, OPT_K4_TAPS, OPT_K4_STEP_CTAS, OPT_K4_SPECULATE, OPT_K5_HALO_CTAS, OPT_K3_CHAIN, OPT_K4_PERSIST, OPT_PIPE_ROWS, OPT_K4_STREAMS, OPT_K4_CHAIN_CTAS, OPT_K1B_DIAG, OPT_K3_BIG_FROM, OPT_COUNT };
int get_option(int opt);
extern std::atomic<unsigned long long> g_launches;

inline int fail_cuda(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VV_ERR_CUDA;
}

#define VV_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            vv::set_error(__VA_ARGS__);         \
            return VV_ERR_INVALID;              \
        }                                       \
    } while (0)

// Call after every kernel launch: counts it and converts launch errors.
#define VV_POST_LAUNCH(name)                                            \
    do {                                                                \
        vv::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return vv::fail_cuda(e__, name);        \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device-side load/store helpers ------------------------------------------------
// Streaming 128-bit accesses.  Loads stay on the read-only path with normal L1
// allocation (the 48-byte-per-thread RGB pattern re-touches each line from three
// instructions); stores bypass L1 since nothing re-reads them in the same kernel.
__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg128_stream(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

// u8 <-> f32 without the conversion pipe (I2F / F2I issue at a quarter of the FP32 rate):
// byte `sel_byte` (PRMT index 0..3 of `v`) as an exact float = bits(2^23 + b) - 2^23 ...
__device__ __forceinline__ float u8_to_float(uint32_t v, uint32_t sel_byte) {
    return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7540u | sel_byte)), 8388608.f);
}
// ... and for v in [0, 255]: the low byte of bits(v + 1.5 * 2^23) is rint(v), ties to even (np.rint).
__device__ __forceinline__ uint32_t rint_u8_bits(float v) { return __float_as_uint(__fadd_rn(v, 12582912.f)); }

// 4 mask bits (bit i <-> pixel i) -> 4 bytes of 0x00 / 0xFF.
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
    return (((nib & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}

// 16 bytes -> 16 bits, bit i = (byte i != 0).
__device__ __forceinline__ uint32_t nonzero_bits16(const uint4 &v) {
    auto nz4 = [](uint32_t w) -> uint32_t {
        // per-byte non-zero test: set bit 7 of every non-zero byte, then gather to 4 bits
        uint32_t t = ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w;   // bit7 of each byte = byte != 0
        t = (t >> 7) & 0x01010101u;
        return (t * 0x10204080u) >> 28;                        // bytes 0..3 -> bits 0..3
    };
    return nz4(v.x) | (nz4(v.y) << 4) | (nz4(v.z) << 8) | (nz4(v.w) << 12);
}

// ---- packed fp32 (two lanes per 64-bit register: FMUL2 / FADD2 / FFMA2 on sm_100) -------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pack2u(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2u(f32x2 v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ f32x2 fadd2_rm(f32x2 a, f32x2 b) {      // round towards minus infinity
    f32x2 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- TMA (bulk async copy) + mbarrier wrappers, sm_90+ PTX -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// global -> shared bulk copy; completion is signalled on `bar` (complete_tx of `bytes`).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// make generic-proxy writes to shared memory visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace vv

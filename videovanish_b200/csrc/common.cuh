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
enum Option { OPT_K1B_EXACT = 0, OPT_COUNT };
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

}  // namespace vv

// Library-level pieces of the C ABI (include/vvb200.h): error reporting, introspection,
// launch accounting and the CUDA-IPC helpers used by the multi-GPU halo blend.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vv {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_options[OPT_COUNT] = {{1}, {2}, {1}, {16}, {512}, {1}, {1}, {1}, {3}, {128}, {4}, {5}, {0}, {5}, {1}, {148 * 8}, {0}, {0}, {1}, {2}, {8}, {2}, {8}};
static const char *const g_option_names[OPT_COUNT] = {"k1b_exact", "k3_nt", "k3_tma", "k3_tma_rows", "k3_tma_threads", "k4_pdl", "k4_npt", "k3_bits", "k3_x2", "k4_pack_ctas", "k4_pack_occ", "k4_lean", "k4_precheck", "k4_step_ctas", "k4_speculate", "k5_halo_ctas", "k3_chain", "k4_persist", "pipe_rows", "k4_streams", "k4_chain_ctas", "k1b_diag", "k3_big_from"};

int get_option(int opt) { return g_options[opt].load(std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace vv

using namespace vv;

extern "C" int vv_version(void) { return 100; }

extern "C" int vv_set_option(const char *name, int value) {
    VV_CHECK_ARG(name, "vv_set_option: NULL name");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, g_option_names[i])) {
            g_options[i].store(value);
            return VV_OK;
        }
    set_error("vv_set_option: unknown option '%s'", name);
    return VV_ERR_INVALID;
}

extern "C" int vv_get_option(const char *name, int *value) {
    VV_CHECK_ARG(name && value, "vv_get_option: NULL argument");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, g_option_names[i])) {
            *value = g_options[i].load();
            return VV_OK;
        }
    set_error("vv_get_option: unknown option '%s'", name);
    return VV_ERR_INVALID;
}

extern "C" const char *vv_last_error(void) { return g_err; }

extern "C" unsigned long long vv_launch_count(void) { return g_launches.load(); }

extern "C" void vv_reset_launch_count(void) { g_launches.store(0); }

extern "C" int vv_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceProperties");
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return VV_OK;
}

extern "C" int vv_ipc_get_handle(const void *dev_ptr, void *handle_out_64B, size_t *offset_out) {
    VV_CHECK_ARG(dev_ptr && handle_out_64B && offset_out, "vv_ipc_get_handle: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    // An IPC handle always denotes a whole allocation; `dev_ptr` may sit inside a block of a caching
    // allocator, so report its offset from the allocation base (driver API, resolved at run time).
    typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
    static GetRange get_range = nullptr;
    if (!get_range) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess || !fn) return fail_cuda(e, "cudaGetDriverEntryPoint(cuMemGetAddressRange)");
        get_range = (GetRange)fn;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0) {
        set_error("vv_ipc_get_handle: cuMemGetAddressRange failed");
        return VV_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcGetMemHandle");
    memcpy(handle_out_64B, &h, 64);
    *offset_out = (size_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return VV_OK;
}

extern "C" int vv_ipc_open_handle(const void *handle_64B, void **mapped_ptr) {
    VV_CHECK_ARG(handle_64B && mapped_ptr, "vv_ipc_open_handle: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64B, 64);
    cudaError_t e = cudaIpcOpenMemHandle(mapped_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcOpenMemHandle");
    return VV_OK;
}

extern "C" int vv_ipc_close_handle(void *mapped_ptr) {
    VV_CHECK_ARG(mapped_ptr, "vv_ipc_close_handle: NULL pointer");
    cudaError_t e = cudaIpcCloseMemHandle(mapped_ptr);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcCloseMemHandle");
    return VV_OK;
}

// Host-buffer pipeline behind the Python drop-in (include/vvb200.h, vv_pipeline_*).
//
// `diffuerase.run_infill_on_frames` (reference diffuerase.py:20-114) receives and returns Python
// lists of per-frame host numpy arrays, so the end-to-end path is PCIe bound.  This runtime keeps
// the GPU busy behind that interface: frames are processed in batches over `n_slots` streams, each
// slot owning its device buffers, so the H2D copy of batch b+1, the kernels of batch b and the
// D2H copy of batch b-1 overlap.  Host pointers that are already page-locked (the arrays our own
// `tools.load_video_frames_from_path` and output allocator hand out) are copied directly; pageable
// arrays are first gathered into a pinned staging ring by a small pool of memcpy threads.
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace vv {

int mask_row_bounds_u8(const uint8_t *mask, int n, int H, int W, int *bounds, cudaStream_t st);   // k8_glue.cu

// ------------------------------------------------------------------ parallel memcpy pool
struct CopyJob {
    void *dst;
    const void *src;
    size_t bytes;
};

class CopyPool {
   public:
    explicit CopyPool(int n) : stop_(false), next_(0), left_(0), gen_(0) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // Splits the jobs into <= 1 MiB pieces and copies them on all workers + the caller.
    void run(const std::vector<CopyJob> &jobs) {
        if (jobs.empty()) return;
        pieces_.clear();
        const size_t chunk = 1 << 20;
        for (const CopyJob &j : jobs)
            for (size_t o = 0; o < j.bytes; o += chunk)
                pieces_.push_back({(uint8_t *)j.dst + o, (const uint8_t *)j.src + o, std::min(chunk, j.bytes - o)});
        {
            std::lock_guard<std::mutex> g(m_);
            next_.store(0);
            left_ = (int)th_.size();
            ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return left_ == 0; });
    }

   private:
    void work() {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= pieces_.size()) break;
            memcpy(pieces_[i].dst, pieces_[i].src, pieces_[i].bytes);
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--left_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::vector<CopyJob> pieces_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    bool stop_;
    std::atomic<size_t> next_;
    int left_;
    unsigned long long gen_;
};

// Worker threads of a memcpy pool (the caller copies too): VV_COPY_THREADS, else min(cap, hardware threads - 1).
// Measured on the 16-thread host of a B200 box: the staging pool of the pageable route gains up to 15 workers (c5_long
// 2.0 k -> 2.35 k frames/s), the background row copies are best at 7 (more only compete with the enqueueing thread).
static int copy_pool_threads(unsigned cap = 7u) {
    const char *e = getenv("VV_COPY_THREADS");
    const unsigned hc = std::thread::hardware_concurrency();
    const int dflt = (int)std::min(cap, hc > 1 ? hc - 1 : 1u);
    if (!e || !*e) return dflt;
    const int n = atoi(e);
    return n >= 1 && n <= 64 ? n : dflt;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *dev_big_in = nullptr;   // fpb * H0*W0*4   masks (C<=4) / original frames
    uint8_t *dev_big_out = nullptr;  // fpb * H0*W0*3   composited frames
    uint8_t *dev_small = nullptr;    // fpb * H0*W0*3  inference-resolution frames / low-res masks
    uint8_t *dev_mask = nullptr;     // fpb * H0*W0     dilated masks when not resident
    void *ws = nullptr;
    uint8_t *pin_in = nullptr;       // lazily allocated staging (pageable inputs only)
    uint8_t *pin_out = nullptr;
    bool pending = false;
    std::vector<CopyJob> copyout;
};

}  // namespace vv

using namespace vv;

struct vv_pipeline {
    int device, H0, W0, fpb, n_slots;
    size_t big_in, big_out, small, maskb, ws_bytes, pin_in_bytes, pin_out_bytes;
    std::vector<Slot> slots;
    CopyPool *pool;
    CopyPool *row_pool;       // second pool for vv_pipeline_host_rows_begin (runs while uploads may use `pool`)
    std::thread *row_thread;  // the background copy of the rows outside the mask bounds, joined by the download
    uint8_t *res_masks;       // resident dilated masks of the last vv_pipeline_pre
    int res_frames, res_cap;
    long long last_rows, last_rows_total;   // rows per direction the last vv_pipeline_post moved / frames x H0
    int *res_bounds_dev;      // their per-frame row ranges (first row with a mask pixel, last + 1), device and host copy
    std::vector<int> res_bounds;
    std::mutex mu;            // one job at a time, like the reference's _job_running guard
};

#define VV_CUDA(call)                                     \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

static int slot_wait(vv_pipeline *p, Slot &s) {
    if (!s.pending) return VV_OK;
    VV_CUDA(cudaEventSynchronize(s.done));
    p->pool->run(s.copyout);
    s.copyout.clear();
    s.pending = false;
    return VV_OK;
}

static int drain(vv_pipeline *p) {
    int rc = VV_OK;
    for (Slot &s : p->slots) {
        const int r = slot_wait(p, s);
        if (r && !rc) rc = r;
    }
    return rc;
}

// Error exit of a batch loop: copies may have been enqueued on a slot stream without its `done` event
// (the batch failed half way), so every stream is synchronised before the call returns and the caller's
// host buffers can go away; staged results of failed jobs are dropped.
static void quiesce(vv_pipeline *p) {
    for (Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        s.copyout.clear();
        s.pending = false;
    }
    cudaGetLastError();
}

static int finish(vv_pipeline *p, int rc) {
    if (rc) {
        quiesce(p);
        return rc;
    }
    rc = drain(p);
    if (rc) quiesce(p);
    return rc;
}

// Copies n per-frame host buffers into consecutive device frames.
static int upload(vv_pipeline *p, Slot &s, const uint8_t *const *src, int n, size_t bytes, uint8_t *dev,
                  size_t pin_off) {
    std::vector<CopyJob> stage;
    for (int i = 0; i < n; ++i) {
        if (is_pinned(src[i])) {
            VV_CUDA(cudaMemcpyAsync(dev + i * bytes, src[i], bytes, cudaMemcpyHostToDevice, s.st));
        } else {
            if (!s.pin_in) VV_CUDA(cudaHostAlloc((void **)&s.pin_in, p->pin_in_bytes, cudaHostAllocDefault));
            stage.push_back({s.pin_in + pin_off + i * bytes, src[i], bytes});
        }
    }
    if (!stage.empty()) {
        p->pool->run(stage);
        for (const CopyJob &j : stage)
            VV_CUDA(cudaMemcpyAsync(dev + ((uint8_t *)j.dst - (s.pin_in + pin_off)), j.dst, j.bytes,
                                    cudaMemcpyHostToDevice, s.st));
    }
    return VV_OK;
}

static int download(vv_pipeline *p, Slot &s, const uint8_t *dev, uint8_t *const *dst, int n, size_t bytes,
                    size_t pin_off) {
    for (int i = 0; i < n; ++i) {
        if (is_pinned(dst[i])) {
            VV_CUDA(cudaMemcpyAsync(dst[i], dev + i * bytes, bytes, cudaMemcpyDeviceToHost, s.st));
        } else {
            if (!s.pin_out) VV_CUDA(cudaHostAlloc((void **)&s.pin_out, p->pin_out_bytes, cudaHostAllocDefault));
            uint8_t *stg = s.pin_out + pin_off + i * bytes;
            VV_CUDA(cudaMemcpyAsync(stg, dev + i * bytes, bytes, cudaMemcpyDeviceToHost, s.st));
            s.copyout.push_back({dst[i], stg, bytes});
        }
    }
    return VV_OK;
}

extern "C" int vv_pipeline_create(vv_pipeline **out, int device, int H0, int W0, int frames_per_batch, int n_slots) {
    VV_CHECK_ARG(out, "vv_pipeline_create: NULL out pointer");
    VV_CHECK_ARG(H0 > 0 && W0 > 0, "vv_pipeline_create: bad geometry");
    if (frames_per_batch <= 0) frames_per_batch = 8;
    if (n_slots <= 0) n_slots = 3;
    VV_CUDA(cudaSetDevice(device));
    vv_pipeline *p = new vv_pipeline();
    p->device = device, p->H0 = H0, p->W0 = W0, p->fpb = frames_per_batch, p->n_slots = n_slots;
    const size_t px = (size_t)H0 * W0;
    p->big_in = frames_per_batch * px * 4;
    p->big_out = frames_per_batch * px * 3;
    p->small = frames_per_batch * px * 3;   // inference-resolution frames never exceed the original size
    p->maskb = frames_per_batch * px;
    p->ws_bytes = std::max(vv_binarize_dilate_workspace_bytes(frames_per_batch, H0, W0),
                           std::max(vv_resize_workspace_bytes(H0, W0),
                                    vv_composite_workspace_bytes(H0, W0)));
    p->pin_in_bytes = p->big_in + p->small + p->maskb;
    p->pin_out_bytes = p->big_out + p->small;
    p->res_masks = nullptr, p->res_frames = 0, p->res_cap = 0;
    p->row_pool = nullptr, p->row_thread = nullptr, p->res_bounds_dev = nullptr;
    p->last_rows = p->last_rows_total = 0;
    p->pool = new CopyPool(copy_pool_threads(15u));
    p->slots.resize(n_slots);
    for (Slot &s : p->slots) {
        cudaError_t e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.dev_big_in, p->big_in);
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.dev_big_out, p->big_out);
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.dev_small, p->small);
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.dev_mask, p->maskb);
        if (e == cudaSuccess) e = cudaMalloc(&s.ws, p->ws_bytes);
        if (e != cudaSuccess) {
            fail_cuda(e, "vv_pipeline_create");
            vv_pipeline_destroy(p);
            return VV_ERR_CUDA;
        }
    }
    *out = p;
    return VV_OK;
}

extern "C" void vv_pipeline_destroy(vv_pipeline *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        cudaFree(s.dev_big_in), cudaFree(s.dev_big_out), cudaFree(s.dev_small), cudaFree(s.dev_mask), cudaFree(s.ws);
        if (s.pin_in) cudaFreeHost(s.pin_in);
        if (s.pin_out) cudaFreeHost(s.pin_out);
        if (s.done) cudaEventDestroy(s.done);
        if (s.st) cudaStreamDestroy(s.st);
    }
    if (p->res_masks) cudaFree(p->res_masks);
    if (p->res_bounds_dev) cudaFree(p->res_bounds_dev);
    if (p->row_thread) {
        p->row_thread->join();
        delete p->row_thread;
    }
    delete p->row_pool;
    delete p->pool;
    delete p;
}

extern "C" int vv_pipeline_pre(vv_pipeline *p, const uint8_t *const *masks, int T, int C, int iterations,
                               uint8_t *const *dilated_out, uint8_t *const *lowres_out, int lh, int lw) {
    VV_CHECK_ARG(p && masks && T > 0, "vv_pipeline_pre: bad argument");
    VV_CHECK_ARG(!lowres_out || (lh > 0 && lw > 0 && (size_t)lh * lw <= (size_t)p->H0 * p->W0),
                 "vv_pipeline_pre: low-res size %dx%d must be positive and not larger than the frame", lh, lw);
    VV_CHECK_ARG(C == 1 || C == 3 || C == 4, "vv_pipeline_pre: C must be 1, 3 or 4 (got %d)", C);
    std::lock_guard<std::mutex> g(p->mu);
    VV_CUDA(cudaSetDevice(p->device));
    const size_t px = (size_t)p->H0 * p->W0, spx = (size_t)lh * lw;
    // keep the dilated masks on the device for vv_pipeline_post (bounded: 16 GiB)
    const bool resident = (size_t)T * px <= ((size_t)16 << 30);
    if (resident && p->res_cap < T) {
        if (p->res_masks) VV_CUDA(cudaFree(p->res_masks));
        p->res_masks = nullptr, p->res_cap = 0;
        VV_CUDA(cudaMalloc((void **)&p->res_masks, (size_t)T * px));
        if (p->res_bounds_dev) VV_CUDA(cudaFree(p->res_bounds_dev));
        p->res_bounds_dev = nullptr;
        VV_CUDA(cudaMalloc((void **)&p->res_bounds_dev, (size_t)T * 2 * sizeof(int)));
        p->res_cap = T;
    }
    p->res_frames = resident ? T : 0;
    p->res_bounds.clear();
    int rc = VV_OK;
    for (int b = 0, t0 = 0; t0 < T && !rc; ++b, t0 += p->fpb) {
        Slot &s = p->slots[b % p->n_slots];
        const int n = std::min(p->fpb, T - t0);
        if ((rc = slot_wait(p, s))) break;
        if ((rc = upload(p, s, masks + t0, n, px * C, s.dev_big_in, 0))) break;
        uint8_t *dil = resident ? p->res_masks + (size_t)t0 * px : s.dev_mask;
        rc = vv_binarize_dilate(s.dev_big_in, n, p->H0, p->W0, C, iterations, dil, lowres_out ? s.dev_small : nullptr,
                                lh, lw, s.ws, p->ws_bytes, s.st);
        if (rc) break;
        if (resident && (rc = mask_row_bounds_u8(dil, n, p->H0, p->W0, p->res_bounds_dev + 2 * t0, s.st))) break;
        if (dilated_out && (rc = download(p, s, dil, dilated_out + t0, n, px, 0))) break;
        if (lowres_out && (rc = download(p, s, s.dev_small, lowres_out + t0, n, spx, p->big_out))) break;
        VV_CUDA(cudaEventRecord(s.done, s.st));
        s.pending = true;
    }
    rc = finish(p, rc);
    if (!rc && resident) {               // every slot stream has been waited for: the row ranges are complete
        p->res_bounds.resize((size_t)T * 2);
        VV_CUDA(cudaMemcpy(p->res_bounds.data(), p->res_bounds_dev, (size_t)T * 2 * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return rc;
}

extern "C" int vv_pipeline_downsize(vv_pipeline *p, const uint8_t *const *frames, int T, int h, int w,
                                    uint8_t *const *small_out) {
    VV_CHECK_ARG(p && frames && small_out && T > 0, "vv_pipeline_downsize: bad argument");
    VV_CHECK_ARG(h > 0 && w > 0 && (size_t)h * w <= (size_t)p->H0 * p->W0,
                 "vv_pipeline_downsize: target %dx%d must be positive and not larger than the frame", h, w);
    std::lock_guard<std::mutex> g(p->mu);
    VV_CUDA(cudaSetDevice(p->device));
    const size_t px = (size_t)p->H0 * p->W0, spx = (size_t)h * w;
    int rc = VV_OK;
    for (int b = 0, t0 = 0; t0 < T && !rc; ++b, t0 += p->fpb) {
        Slot &s = p->slots[b % p->n_slots];
        const int n = std::min(p->fpb, T - t0);
        if ((rc = slot_wait(p, s))) break;
        if ((rc = upload(p, s, frames + t0, n, px * 3, s.dev_big_in, 0))) break;
        rc = vv_resize(s.dev_big_in, n, p->H0, p->W0, 3, s.dev_small, h, w, VV_INTER_LINEAR, s.ws, p->ws_bytes, s.st);
        if (rc) break;
        if ((rc = download(p, s, s.dev_small, small_out + t0, n, spx * 3, p->big_out))) break;
        VV_CUDA(cudaEventRecord(s.done, s.st));
        s.pending = true;
    }
    return finish(p, rc);
}

extern "C" int vv_pipeline_post(vv_pipeline *p, const uint8_t *const *inpainted, int h, int w,
                                const uint8_t *const *orig, const uint8_t *const *dilated, int T, float feather_px,
                                int keep_unmasked, uint8_t *const *out) {
    VV_CHECK_ARG(p && inpainted && out && T > 0, "vv_pipeline_post: bad argument");
    VV_CHECK_ARG(h > 0 && w > 0, "vv_pipeline_post: bad inpainted size %dx%d", h, w);
    if ((size_t)h * w > (size_t)p->H0 * p->W0) {
        set_error("vv_pipeline_post: inpainted frames (%dx%d) larger than the original (%dx%d)", h, w, p->H0, p->W0);
        return VV_ERR_UNSUPPORTED;
    }
    VV_CHECK_ARG(!keep_unmasked || orig, "vv_pipeline_post: original frames required when keep_unmasked != 0");
    std::lock_guard<std::mutex> g(p->mu);
    VV_CHECK_ARG(!keep_unmasked || dilated || p->res_frames >= T,
                 "vv_pipeline_post: no dilated masks supplied and none resident from vv_pipeline_pre");
    VV_CUDA(cudaSetDevice(p->device));
    const size_t px = (size_t)p->H0 * p->W0, spx = (size_t)h * w;
    // Row-bounded mode: outside the rows a frame's (resident) dilated mask + feather radius reaches, the composite
    // returns the original frame (alpha = 0 -> rint(orig) = orig, diffuerase.py:112).  Those rows are copied host ->
    // host from `orig` to `out` by a background memcpy pool; only the rows in between are uploaded, and downloaded.
    // K3 still runs over whole frames (its time is noise next to PCIe): the device rows that were not uploaded hold
    // stale bytes, which only reach output rows that are not downloaded.
    std::vector<int> lo, hi;
    const size_t row_bytes = (size_t)p->W0 * 3;
    if (keep_unmasked && !dilated && p->res_bounds.size() >= (size_t)T * 2 && get_option(OPT_PIPE_ROWS) != 0) {
        const int margin = (feather_px > 0.f ? std::max(0, (int)ceilf(feather_px) - 1) : 0) + 1;
        long long rows = 0;
        bool ok = true;
        lo.resize(T), hi.resize(T);
        for (int i = 0; i < T && ok; ++i) {
            const int a = p->res_bounds[2 * i], z = p->res_bounds[2 * i + 1];
            lo[i] = z > a ? std::max(0, (a - margin) & ~15) : 0;                       // 16-row aligned (K3 strips), (0,0) = empty
            hi[i] = z > a ? std::min(p->H0, (z + margin + 15) & ~15) : 0;
            rows += hi[i] - lo[i];
            ok = orig[i] && out[i] && orig[i] != out[i] && is_pinned(orig[i]) && is_pinned(out[i]);
        }
        if (!ok || rows * 4 > (long long)T * p->H0 * 3) lo.clear(), hi.clear();      // worth it below 75 % of the rows
    }
    const bool by_rows = !lo.empty();
    p->last_rows_total = (long long)T * p->H0;
    p->last_rows = p->last_rows_total;
    if (by_rows) {
        p->last_rows = 0;
        for (int i = 0; i < T; ++i) p->last_rows += hi[i] - lo[i];
        if (p->row_thread) {
            p->row_thread->join();
            delete p->row_thread;
            p->row_thread = nullptr;
        }
        if (!p->row_pool) p->row_pool = new CopyPool(copy_pool_threads());
        std::vector<CopyJob> jobs;
        for (int i = 0; i < T; ++i) {
            if (lo[i] > 0) jobs.push_back({out[i], orig[i], (size_t)lo[i] * row_bytes});
            if (hi[i] < p->H0)
                jobs.push_back({out[i] + (size_t)hi[i] * row_bytes, orig[i] + (size_t)hi[i] * row_bytes, (size_t)(p->H0 - hi[i]) * row_bytes});
        }
        CopyPool *pool = p->row_pool;
        p->row_thread = new std::thread([pool, jobs] { pool->run(jobs); });
    }
    int rc = VV_OK;
    for (int b = 0, t0 = 0; t0 < T && !rc; ++b, t0 += p->fpb) {
        Slot &s = p->slots[b % p->n_slots];
        const int n = std::min(p->fpb, T - t0);
        if ((rc = slot_wait(p, s))) break;
        if ((rc = upload(p, s, inpainted + t0, n, spx * 3, s.dev_small, p->big_in))) break;
        const uint8_t *mk = nullptr;
        if (by_rows) {
            cudaError_t ce = cudaSuccess;
            for (int i = 0; i < n && ce == cudaSuccess; ++i) {
                const int f = t0 + i;
                if (hi[f] > lo[f])
                    ce = cudaMemcpyAsync(s.dev_big_in + i * px * 3 + (size_t)lo[f] * row_bytes, orig[f] + (size_t)lo[f] * row_bytes,
                                         (size_t)(hi[f] - lo[f]) * row_bytes, cudaMemcpyHostToDevice, s.st);
            }
            if (ce != cudaSuccess) {
                rc = fail_cuda(ce, "cudaMemcpyAsync(orig rows)");
                break;
            }
            mk = p->res_masks + (size_t)t0 * px;
            rc = vv_upscale_feather_composite(s.dev_small, n, h, w, s.dev_big_in, mk, p->H0, p->W0, feather_px, keep_unmasked,
                                              s.dev_big_out, s.ws, p->ws_bytes, s.st);
            if (rc) break;
            for (int i = 0; i < n && ce == cudaSuccess; ++i) {
                const int f = t0 + i;
                if (hi[f] > lo[f])
                    ce = cudaMemcpyAsync(out[f] + (size_t)lo[f] * row_bytes, s.dev_big_out + i * px * 3 + (size_t)lo[f] * row_bytes,
                                         (size_t)(hi[f] - lo[f]) * row_bytes, cudaMemcpyDeviceToHost, s.st);
            }
            if (ce != cudaSuccess) {
                rc = fail_cuda(ce, "cudaMemcpyAsync(out rows)");
                break;
            }
            VV_CUDA(cudaEventRecord(s.done, s.st));
            s.pending = true;
            continue;
        }
        if (keep_unmasked) {
            if ((rc = upload(p, s, orig + t0, n, px * 3, s.dev_big_in, 0))) break;
            if (dilated) {
                if ((rc = upload(p, s, dilated + t0, n, px, s.dev_mask, p->big_in + p->small))) break;
                mk = s.dev_mask;
            } else {
                mk = p->res_masks + (size_t)t0 * px;
            }
        }
        rc = vv_upscale_feather_composite(s.dev_small, n, h, w, s.dev_big_in, mk, p->H0, p->W0, feather_px,
                                          keep_unmasked, s.dev_big_out, s.ws, p->ws_bytes, s.st);
        if (rc) break;
        if ((rc = download(p, s, s.dev_big_out, out + t0, n, px * 3, 0))) break;
        VV_CUDA(cudaEventRecord(s.done, s.st));
        s.pending = true;
    }
    rc = finish(p, rc);
    if (p->row_thread) {                 // the rows copied from the original frames (also on the error path)
        p->row_thread->join();
        delete p->row_thread;
        p->row_thread = nullptr;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// Device-resident clips (videovanish_b200/wrappers.py): per-frame host buffers <-> one contiguous device
// array, through the same slots (direct DMA for page-locked buffers, pinned ring + memcpy pool otherwise).
extern "C" int vv_pipeline_upload(vv_pipeline *p, const uint8_t *const *src, int T, size_t frame_bytes, uint8_t *dev_dst,
                                  void *stream) {
    VV_CHECK_ARG(p && src && dev_dst && T > 0 && frame_bytes > 0, "vv_pipeline_upload: bad argument");
    VV_CHECK_ARG(frame_bytes <= (size_t)p->H0 * p->W0 * 4, "vv_pipeline_upload: frame larger than the context geometry");
    std::lock_guard<std::mutex> g(p->mu);
    VV_CUDA(cudaSetDevice(p->device));
    // `dev_dst` usually comes from a stream-ordered caching allocator: the block may have been freed by work that is
    // still PENDING on the consumer stream.  The copies run on the slot streams, so they must not start before the
    // consumer stream has got to this point (otherwise they overwrite memory that earlier kernels still use).
    cudaEvent_t fence;
    VV_CUDA(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
    cudaError_t fe = cudaEventRecord(fence, (cudaStream_t)stream);
    int rc = fe == cudaSuccess ? VV_OK : fail_cuda(fe, "cudaEventRecord");
    for (int b = 0, t0 = 0; t0 < T && !rc; ++b, t0 += p->fpb) {
        Slot &s = p->slots[b % p->n_slots];
        const int n = std::min(p->fpb, T - t0);
        if ((rc = slot_wait(p, s))) break;
        if (b < p->n_slots && (fe = cudaStreamWaitEvent(s.st, fence, 0)) != cudaSuccess) {
            rc = fail_cuda(fe, "cudaStreamWaitEvent");
            break;
        }
        if ((rc = upload(p, s, src + t0, n, frame_bytes, dev_dst + (size_t)t0 * frame_bytes, 0))) break;
        fe = cudaEventRecord(s.done, s.st);
        if (fe != cudaSuccess) {
            rc = fail_cuda(fe, "cudaEventRecord");
            break;
        }
        s.pending = true;
    }
    cudaEventDestroy(fence);
    if (rc) return finish(p, rc);
    // The consumer stream waits for every slot ON THE DEVICE; the host does not wait: pageable sources have
    // already been copied into the pinned ring, page-locked ones are read by DMA until `stream` gets there.
    for (Slot &s : p->slots)
        if (s.pending) VV_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, s.done, 0));
    return VV_OK;
}

extern "C" int vv_pipeline_download(vv_pipeline *p, const uint8_t *dev_src, int T, size_t frame_bytes, uint8_t *const *dst,
                                    void *stream) {
    VV_CHECK_ARG(p && dev_src && dst && T > 0 && frame_bytes > 0, "vv_pipeline_download: bad argument");
    VV_CHECK_ARG(frame_bytes <= (size_t)p->H0 * p->W0 * 3, "vv_pipeline_download: frame larger than the context geometry");
    std::lock_guard<std::mutex> g(p->mu);
    VV_CUDA(cudaSetDevice(p->device));
    // the producer stream's work so far must be complete before any slot copies
    cudaEvent_t ready;
    VV_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ready, (cudaStream_t)stream);
    int rc = e == cudaSuccess ? VV_OK : fail_cuda(e, "cudaEventRecord");
    for (int b = 0, t0 = 0; t0 < T && !rc; ++b, t0 += p->fpb) {
        Slot &s = p->slots[b % p->n_slots];
        const int n = std::min(p->fpb, T - t0);
        if ((rc = slot_wait(p, s))) break;
        if (b < p->n_slots && (e = cudaStreamWaitEvent(s.st, ready, 0)) != cudaSuccess) {
            rc = fail_cuda(e, "cudaStreamWaitEvent");
            break;
        }
        if ((rc = download(p, s, dev_src + (size_t)t0 * frame_bytes, dst + t0, n, frame_bytes, 0))) break;
        e = cudaEventRecord(s.done, s.st);
        if (e != cudaSuccess) {
            rc = fail_cuda(e, "cudaEventRecord");
            break;
        }
        s.pending = true;
    }
    rc = finish(p, rc);
    cudaEventDestroy(ready);
    return rc;
}

// ---- row-bounded results ---------------------------------------------------------------------------------------
// The composite (diffuerase.py:70-112) only changes pixels within the feather radius of a mask pixel: outside the row
// range [lo, hi) of a frame (vv_mask_row_bounds over K1's bit plane) the finished frame IS the input frame.  The
// host-list front end therefore copies those rows host -> host from the caller's input frames, on a memcpy pool in the
// background while the clip is uploaded and processed, and only the rows inside the range cross PCIe on the way back.
extern "C" int vv_pipeline_host_rows_begin(vv_pipeline *p, int T, int H, size_t row_bytes, uint8_t *const *dst,
                                           const uint8_t *const *src, const int *lo, const int *hi) {
    VV_CHECK_ARG(p && dst && src && lo && hi && T > 0 && H > 0 && row_bytes > 0, "vv_pipeline_host_rows_begin: bad argument");
    std::lock_guard<std::mutex> g(p->mu);
    if (p->row_thread) {                 // a previous call whose download never came: finish it first
        p->row_thread->join();
        delete p->row_thread;
        p->row_thread = nullptr;
    }
    if (!p->row_pool) p->row_pool = new CopyPool(copy_pool_threads());
    std::vector<CopyJob> jobs;
    for (int i = 0; i < T; ++i) {
        VV_CHECK_ARG(dst[i] && src[i] && lo[i] >= 0 && lo[i] <= hi[i] && hi[i] <= H,
                     "vv_pipeline_host_rows_begin: bad row range [%d,%d) of frame %d", lo[i], hi[i], i);
        if (lo[i] > 0) jobs.push_back({dst[i], src[i], (size_t)lo[i] * row_bytes});
        if (hi[i] < H) jobs.push_back({dst[i] + (size_t)hi[i] * row_bytes, src[i] + (size_t)hi[i] * row_bytes,
                                       (size_t)(H - hi[i]) * row_bytes});
    }
    CopyPool *pool = p->row_pool;
    p->row_thread = new std::thread([pool, jobs] { pool->run(jobs); });
    return VV_OK;
}

extern "C" int vv_pipeline_download_rows(vv_pipeline *p, const uint8_t *dev_src, int T, int H, size_t row_bytes,
                                         uint8_t *const *dst, const int *lo, const int *hi, void *stream) {
    VV_CHECK_ARG(p && dev_src && dst && lo && hi && T > 0 && H > 0 && row_bytes > 0, "vv_pipeline_download_rows: bad argument");
    VV_CHECK_ARG((size_t)H * row_bytes <= (size_t)p->H0 * p->W0 * 3, "vv_pipeline_download_rows: frame larger than the context geometry");
    std::lock_guard<std::mutex> g(p->mu);
    VV_CUDA(cudaSetDevice(p->device));
    for (int i = 0; i < T; ++i)
        VV_CHECK_ARG(dst[i] && lo[i] >= 0 && lo[i] <= hi[i] && hi[i] <= H && is_pinned(dst[i]),
                     "vv_pipeline_download_rows: frame %d needs a page-locked destination and a row range inside the frame", i);
    cudaEvent_t ready;
    VV_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ready, (cudaStream_t)stream);
    int rc = e == cudaSuccess ? VV_OK : fail_cuda(e, "cudaEventRecord");
    const size_t frame_bytes = (size_t)H * row_bytes;
    for (int s = 0; s < p->n_slots && !rc; ++s)
        if ((e = cudaStreamWaitEvent(p->slots[s].st, ready, 0)) != cudaSuccess) rc = fail_cuda(e, "cudaStreamWaitEvent");
    for (int i = 0; i < T && !rc; ++i) {
        if (hi[i] <= lo[i]) continue;
        Slot &s = p->slots[(i / p->fpb) % p->n_slots];
        const size_t off = (size_t)lo[i] * row_bytes;
        e = cudaMemcpyAsync(dst[i] + off, dev_src + (size_t)i * frame_bytes + off, (size_t)(hi[i] - lo[i]) * row_bytes,
                            cudaMemcpyDeviceToHost, s.st);
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaMemcpyAsync(rows)");
    }
    for (int s = 0; s < p->n_slots; ++s) {                       // also on the error path: no DMA into freed buffers
        e = cudaStreamSynchronize(p->slots[s].st);
        if (e != cudaSuccess && !rc) rc = fail_cuda(e, "cudaStreamSynchronize");
    }
    cudaEventDestroy(ready);
    if (p->row_thread) {                                         // the rows copied from the input frames
        p->row_thread->join();
        delete p->row_thread;
        p->row_thread = nullptr;
    }
    return rc;
}

extern "C" int vv_pipeline_last_rows(vv_pipeline *p, long long *rows_moved, long long *rows_total) {
    VV_CHECK_ARG(p && rows_moved && rows_total, "vv_pipeline_last_rows: NULL argument");
    std::lock_guard<std::mutex> g(p->mu);
    *rows_moved = p->last_rows, *rows_total = p->last_rows_total;
    return VV_OK;
}

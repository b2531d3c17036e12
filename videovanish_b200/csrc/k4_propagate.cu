// K4: ProPainter-style flow-guided propagation prior (image propagation, 'nearest').
// Replaces the un-vendored `propainter.forward` call at /root/reference/diffuerase.py:49-57
// [BidirectionalPropagation(learnable=False) + fbConsistencyCheck + flow_warp; upstream files
// cited in oracle/propagation.py].  PARITY UNPINNED by the reference; the oracle is the torch-CPU
// restatement and its explicit float32 model (oracle/propagation.py::_step), which this kernel
// follows operation for operation:
//   sample position  p + flow_prop(p), normalised and un-normalised exactly as flow_warp +
//                    grid_sample(align_corners=True) do in fp32
//   flow check       bilinear (zeros padding, fma chain) warp of flow_check, then
//                    |f + bw|^2 < 0.01 (|f|^2 + |bw|^2) + 0.5
//   mask validity    bilinear warp of the previous hole mask > 0.1
//   fill             hole & valid & !mask_valid -> copy the NEAREST previous pixel (or zero padding)
// With 'nearest' and binary masks the pixel path is a pure select, so pixels stay u8: the state
// is one packed word per pixel, R | G<<8 | B<<16 | state<<24 (bit0 hole, bit1 "value is the
// masked frame's 0.0").  The four bilinear taps of the previous state give both the mask warp and
// the nearest pixel, so one 4-word gather serves both.
//
// Structure: k4_pack turns frames + masks into the packed state ONCE (fully parallel over all
// frames, streaming) and records each frame's hole pixels in a list.  The scan itself is serial in
// time but touches hole pixels only: every step launch walks the hole list of one frame per
// sub-video, in place on the state buffer (backward pass, then forward pass over the same buffer).
// Up to 32 independent sub-videos (propainter/inference.py windows: 50 frames + 10 pad each side)
// advance in lock step, so the chain is 2*(window length - 1) short launches whatever the clip length.
//
// Frame placement: the frames a window KEEPS live directly in the caller's output array, its pad
// frames (which upstream computes and discards) in workspace scratch, so the result needs no
// compaction copy afterwards (`frame_ptr`).
#include "common.cuh"
#include <mutex>

namespace vv {

constexpr uint32_t ST_HOLE = 1u << 24;
constexpr uint32_t ST_ZERO = 2u << 24;
constexpr int K4_MAX_SUB = 32;

struct SubDesc {
    int start;            // first frame of the window in the clip arrays
    int len;              // frames in the window
    int pad_s;            // leading pad frames (state kept in scratch)
    int keep;             // frames [pad_s, pad_s + keep) of the window are written to the output array
    long long out_frame;  // first frame of the window in the hole-list arrays (windows concatenated)
    long long keep_out;   // output-array frame index of the first kept frame
    long long pad_out;    // scratch frame index of the window's first pad frame
};
struct SubBatch {
    SubDesc sub[K4_MAX_SUB];
    int n;
};

// State of frame `idx` of a window: kept frames sit in `out`, pad frames in `pads`.
__host__ __device__ __forceinline__ uint32_t *frame_ptr(const SubDesc &sd, int idx, uint32_t *out, uint32_t *pads,
                                                        long long npx) {
    if (idx < sd.pad_s) return pads + (sd.pad_out + idx) * npx;
    if (idx < sd.pad_s + sd.keep) return out + (sd.keep_out + (idx - sd.pad_s)) * npx;
    return pads + (sd.pad_out + (idx - sd.keep)) * npx;
}

__device__ __forceinline__ float unnormalized(float pos, int size) {
    // flow_warp: 2*g/max(size-1,1) - 1 ; grid_sample(align_corners=True): (c+1) * ((size-1)/2)
    const float n = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, pos), (float)max(size - 1, 1)), 1.0f);
    return __fmul_rn(__fadd_rn(n, 1.0f), (float)(size - 1) * 0.5f);
}

// The arithmetic of one hole pixel, split into the parts that depend on the flows only (sample position, forward /
// backward consistency - computable for every frame at once) and the part that depends on the serial scan (the
// previous frame's state).
struct Taps {
    float ix, iy, x0f, y0f, nw, ne, sw, se;
    uint32_t i00;             // linear index of tap (y0, x0): only used when the tap is inside the frame
    bool xa, xb, ya, yb;      // which taps lie inside the frame (the others are zeros padding)
};
__device__ __forceinline__ float2 sample_pos(int x, int y, float2 f, int h, int w) {
    return make_float2(unnormalized(__fadd_rn((float)x, f.x), w), unnormalized(__fadd_rn((float)y, f.y), h));
}
__device__ __forceinline__ Taps make_taps(float ix, float iy, int h, int w) {
    Taps t;
    t.ix = ix, t.iy = iy;
    t.x0f = floorf(ix), t.y0f = floorf(iy);
    const float wx = __fsub_rn(ix, t.x0f), wy = __fsub_rn(iy, t.y0f);
    const float ex = __fsub_rn(1.f, wx), sy = __fsub_rn(1.f, wy);
    t.nw = __fmul_rn(sy, ex), t.ne = __fmul_rn(sy, wx), t.sw = __fmul_rn(wy, ex), t.se = __fmul_rn(wy, wx);
    // clamp before the int conversion so absurd flows cannot overflow
    const int x0 = (int)fminf(fmaxf(t.x0f, -4.f), (float)w + 4.f);
    const int y0 = (int)fminf(fmaxf(t.y0f, -4.f), (float)h + 4.f);
    t.xa = (unsigned)x0 < (unsigned)w, t.xb = (unsigned)(x0 + 1) < (unsigned)w;
    t.ya = (unsigned)y0 < (unsigned)h, t.yb = (unsigned)(y0 + 1) < (unsigned)h;
    t.i00 = (uint32_t)(y0 * w + x0);
    return t;
}
// Tap loads are separate from the arithmetic so that a caller can issue all eight (flow + state) before anything
// waits on them: one memory round trip per hole.
struct FlowTaps {
    float2 c00, c01, c10, c11;
};
struct StateTaps {
    uint32_t p00, p01, p10, p11;
};
__device__ __forceinline__ FlowTaps load_flow_taps(const Taps &t, const float2 *__restrict__ flow_check, int w) {
    FlowTaps c;
    c.c00 = c.c01 = c.c10 = c.c11 = make_float2(0.f, 0.f);          // out-of-frame taps: zeros padding
    if (t.ya && t.xa) c.c00 = __ldg(flow_check + t.i00);
    if (t.ya && t.xb) c.c01 = __ldg(flow_check + (t.i00 + 1u));
    if (t.yb && t.xa) c.c10 = __ldg(flow_check + (t.i00 + (uint32_t)w));
    if (t.yb && t.xb) c.c11 = __ldg(flow_check + (t.i00 + (uint32_t)w + 1u));
    return c;
}
// CG: the state was written by other CTAs of the SAME launch (k4_steps_persistent): read it from L2, never from L1
template <bool CG = false>
__device__ __forceinline__ StateTaps load_state_taps(const Taps &t, const uint32_t *prev, int w) {
    StateTaps p;
    p.p00 = p.p01 = p.p10 = p.p11 = 0;                              // out-of-frame taps: zeros padding, not a hole
    if (t.ya && t.xa) p.p00 = CG ? __ldcg(prev + t.i00) : prev[t.i00];
    if (t.ya && t.xb) p.p01 = CG ? __ldcg(prev + (t.i00 + 1u)) : prev[t.i00 + 1u];
    if (t.yb && t.xa) p.p10 = CG ? __ldcg(prev + (t.i00 + (uint32_t)w)) : prev[t.i00 + (uint32_t)w];
    if (t.yb && t.xb) p.p11 = CG ? __ldcg(prev + (t.i00 + (uint32_t)w + 1u)) : prev[t.i00 + (uint32_t)w + 1u];
    return p;
}
// fbConsistencyCheck for one pixel: bilinear (zeros padding, torch CPU FMA chain) warp of the check flow at the
// sample position, then |f + bw|^2 < 0.01 (|f|^2 + |bw|^2) + 0.5
__device__ __forceinline__ bool flow_consistent(const float2 f, const Taps &t, const FlowTaps &c) {
    // r = a*nw; r = fma(b, ne, r); r = fma(c, sw, r); r = fma(d, se, r)
    const float bwx = __fmaf_rn(c.c11.x, t.se, __fmaf_rn(c.c10.x, t.sw, __fmaf_rn(c.c01.x, t.ne, __fmul_rn(c.c00.x, t.nw))));
    const float bwy = __fmaf_rn(c.c11.y, t.se, __fmaf_rn(c.c10.y, t.sw, __fmaf_rn(c.c01.y, t.ne, __fmul_rn(c.c00.y, t.nw))));
    const float dx = __fadd_rn(f.x, bwx), dy = __fadd_rn(f.y, bwy);
    const float diff = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float mag = __fadd_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)),
                                __fadd_rn(__fmul_rn(bwx, bwx), __fmul_rn(bwy, bwy)));
    return diff < __fadd_rn(__fmul_rn(0.01f, mag), 0.5f);
}
// The state-dependent part for a pixel whose flow passed the check: bilinear warp of the previous hole mask > 0.1
// blocks the fill; otherwise the NEAREST previous pixel (one of the four taps) or the zero padding is copied.
__device__ __forceinline__ uint32_t fill_from_prev(const Taps &t, uint32_t cur, const StateTaps &p) {
    const float h00 = (p.p00 & ST_HOLE) ? 1.f : 0.f, h01 = (p.p01 & ST_HOLE) ? 1.f : 0.f;
    const float h10 = (p.p10 & ST_HOLE) ? 1.f : 0.f, h11 = (p.p11 & ST_HOLE) ? 1.f : 0.f;
    const float mpv = __fmaf_rn(h11, t.se, __fmaf_rn(h10, t.sw, __fmaf_rn(h01, t.ne, __fmul_rn(h00, t.nw))));
    // nearest source: rint (half-to-even) of the same un-normalised position
    const float xr = rintf(t.ix), yr = rintf(t.iy);
    const bool right = xr > t.x0f, down = yr > t.y0f;
    const bool inb = (right ? t.xb : t.xa) && (down ? t.yb : t.ya);
    const uint32_t src = down ? (right ? p.p11 : p.p10) : (right ? p.p01 : p.p00);
    const uint32_t filled = inb ? (src & ~ST_HOLE) : ST_ZERO;
    return mpv > 0.1f ? cur : filled;
}

// One hole pixel, everything in one go (forward pass; backward pass when nothing was precomputed).
template <bool CG = false>
__device__ __forceinline__ uint32_t propagate_pixel(int x, int y, int h, int w, uint32_t cur, const float2 f,
                                                    const float2 *__restrict__ flow_check, const uint32_t *prev) {
    const float2 pos = sample_pos(x, y, f, h, w);
    const Taps t = make_taps(pos.x, pos.y, h, w);
    const FlowTaps c = load_flow_taps(t, flow_check, w);           // all eight taps requested together
    const StateTaps p = load_state_taps<CG>(t, prev, w);
    return flow_consistent(f, t, c) ? fill_from_prev(t, cur, p) : cur;
}

constexpr uint32_t XY_INVALID = 1u << 31;     // backward-list entry whose flow failed the consistency check

// Hole lists.  One entry per hole pixel: position (x | y << 16) and the flow vector that will
// propagate INTO that pixel, so that a step's dependency chain is two memory round trips (entry,
// then the 8 taps).  Two sets per frame: `l1` drives the backward pass (flow = flows_f[frame]),
// `l2` the forward pass (flow = flows_b[frame-1]) and only holds the holes the backward pass left.
struct HoleLists {
    uint32_t *xy;        // [frames][npx]
    float2 *flow;        // [frames][npx]
    uint32_t *count;     // [frames]
};

// Block-level append.  Hole entries are first collected in a shared-memory queue (one shared
// atomic per warp and call), then the block reserves its range of the frame's list with ONE global
// atomic and writes it out coalesced - a per-warp global atomic on the single per-frame counter
// serialises in L2 and dominated the first version of this stage.
constexpr int K4_BLOCK = 256;
constexpr int K4_PACK_UNROLL = 4;             // 4-pixel groups per thread and round in k4_pack
constexpr int K4_QCAP = 2 * 4 * K4_BLOCK * K4_PACK_UNROLL;   // queue entries: two k4_pack push rounds
struct BlockQueue {
    uint32_t xy[K4_QCAP];
    uint32_t count, base;
};

// All threads of the block must call (contains barriers).  `flow_frame` is gathered for each entry.
// With `force == false` the queue is only written out when another push round might overflow it.
// `check_frame` != NULL (backward-pass lists with "k4_precheck"): the entry does not carry the flow but what the
// serial step needs of it - the un-normalised sample position - and the verdict of the forward / backward
// consistency check in bit 31 of the position word, both computed here, fully parallel over all frames.
__device__ __forceinline__ void queue_flush(BlockQueue &q, const HoleLists &l, long long of, long long npx, int h, int w,
                                            const float2 *__restrict__ flow_frame, const float2 *__restrict__ check_frame,
                                            bool force) {
    __syncthreads();
    const uint32_t n = q.count;
    __syncthreads();                                            // everyone has read the count before it can change
    if (!force && n + 4 * K4_BLOCK * K4_PACK_UNROLL <= K4_QCAP) return;   // block-uniform
    if (n) {
        if (threadIdx.x == 0) q.base = atomicAdd(l.count + of, n);
        __syncthreads();
        const long long dst = of * npx + q.base;
        // four entries per thread and trip: the flow gathers of a batch are in flight together
        for (uint32_t j0 = threadIdx.x; j0 < n; j0 += 4 * blockDim.x) {
            uint32_t xy[4];
            float2 f[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + u * blockDim.x;
                xy[u] = j < n ? q.xy[j] : 0u;
                f[u] = __ldg(flow_frame + (long long)(xy[u] >> 16) * w + (xy[u] & 0xffffu));
            }
            if (check_frame != nullptr) {                       // block-uniform
                // two entries at a time: their 8 check-flow taps are in flight together (four would spill)
#pragma unroll
                for (int u0 = 0; u0 < 4; u0 += 2) {
                    Taps tp[2];
                    FlowTaps ct[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float2 pos = sample_pos((int)(xy[u0 + u] & 0xffffu), (int)(xy[u0 + u] >> 16), f[u0 + u], h, w);
                        tp[u] = make_taps(pos.x, pos.y, h, w);
                        ct[u] = load_flow_taps(tp[u], check_frame, w);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (!flow_consistent(f[u0 + u], tp[u], ct[u])) xy[u0 + u] |= XY_INVALID;
                        f[u0 + u] = make_float2(tp[u].ix, tp[u].iy);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + u * blockDim.x;
                if (j < n) l.xy[dst + j] = xy[u], l.flow[dst + j] = f[u];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) q.count = 0;
    __syncthreads();
}

// ---- k4_pack: frames + masks -> packed state, and the per-frame lists of hole pixels -----------
// One launch for every frame of every window of the batch (blockIdx.y = output frame).  Holes of
// the last frame of a window are never touched by the backward pass, so they go straight to `l2`.
template <bool VEC, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
    k4_pack(const uint8_t *__restrict__ frames, const uint8_t *__restrict__ masks, uint32_t *__restrict__ state,
            uint32_t *__restrict__ pads, const float2 *__restrict__ flows_f, const float2 *__restrict__ flows_b,
            HoleLists l1, HoleLists l2, int h, int w, long long first_out_frame, int precheck,
            const __grid_constant__ SubBatch batch) {
    const long long of = first_out_frame + blockIdx.y;          // output frame handled by this CTA row
    int s = 0;
    while (s + 1 < batch.n && batch.sub[s + 1].out_frame <= of) ++s;
    const int idx = (int)(of - batch.sub[s].out_frame), len = batch.sub[s].len;
    const long long gframe = batch.sub[s].start + idx;
    const long long npx = (long long)h * w;
    const uint8_t *fr = frames + gframe * npx * 3;
    const uint8_t *mk = masks + gframe * npx;
    uint32_t *dst = frame_ptr(batch.sub[s], idx, state, pads, npx);
    const bool last = idx == len - 1;
    const bool listed = len > 1;                                // a single-frame window has no steps at all
    // flow that propagates INTO this frame: backward pass flows_f[gframe], forward pass flows_b[gframe-1]
    const float2 *pflow = last ? flows_b + (gframe - 1) * npx : flows_f + gframe * npx;
    // backward-pass entries: consistency of flows_f[gframe] against flows_b[gframe], decided here
    const float2 *pcheck = (precheck && !last && listed) ? flows_b + gframe * npx : nullptr;
    const HoleLists dl = {last ? l2.xy : l1.xy, last ? l2.flow : l1.flow, last ? l2.count : l1.count};   // selected member-wise: stays in registers
    __shared__ BlockQueue q;
    if (threadIdx.x == 0) q.count = 0;
    __syncthreads();
    const long long ngroups = (npx + 3) >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long iters = (ngroups + stride * K4_PACK_UNROLL - 1) / (stride * K4_PACK_UNROLL);
    const int lane = threadIdx.x & 31;
    // VEC (w % 4 == 0): a group never straddles a row end; its (row, first column) is advanced incrementally
    // from one group to the next (stride groups further) - one division per thread instead of one per group.
    const uint32_t w4 = VEC ? (uint32_t)w >> 2 : 1u;
    const uint32_t step_y = (uint32_t)(stride / w4), step_x = (uint32_t)(stride - (long long)step_y * w4);
    const uint32_t g_first = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t gy = g_first / w4, gx = g_first - gy * w4;
    for (long long itn = 0; itn < iters; ++itn) {
        // K4_PACK_UNROLL groups per thread: all loads first (memory-level parallelism), one flush check per round
        uint32_t c[K4_PACK_UNROLL][4];
        uint32_t m4[K4_PACK_UNROLL], fa[K4_PACK_UNROLL], fb2[K4_PACK_UNROLL], fd[K4_PACK_UNROLL];
        long long p0[K4_PACK_UNROLL];
        uint32_t xy0[K4_PACK_UNROLL];
        int n[K4_PACK_UNROLL];
#pragma unroll
        for (int u = 0; u < K4_PACK_UNROLL; ++u) {
            const long long g = (itn * K4_PACK_UNROLL + u) * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
            p0[u] = g * 4;
            if (VEC) {
                xy0[u] = (gx << 2) | (gy << 16);
                gy += step_y, gx += step_x;
                if (gx >= w4) gx -= w4, ++gy;
            }
            n[u] = g < ngroups ? (VEC ? 4 : (int)min(4LL, npx - p0[u])) : 0;
            if (VEC && n[u]) {
                m4[u] = __ldg(reinterpret_cast<const uint32_t *>(mk + p0[u]));
                const uint32_t *f3 = reinterpret_cast<const uint32_t *>(fr + p0[u] * 3);
                fa[u] = __ldg(f3), fb2[u] = __ldg(f3 + 1), fd[u] = __ldg(f3 + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < K4_PACK_UNROLL; ++u) {
            c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0;
            if (!n[u]) continue;
            if (VEC) {
                c[u][0] = fa[u] & 0x00ffffffu;
                c[u][1] = (fa[u] >> 24) | ((fb2[u] & 0x0000ffffu) << 8);
                c[u][2] = (fb2[u] >> 16) | ((fd[u] & 0x000000ffu) << 16);
                c[u][3] = fd[u] >> 8;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (byte_of(m4[u], i)) c[u][i] = ST_HOLE | ST_ZERO;
                *reinterpret_cast<uint4 *>(dst + p0[u]) = make_uint4(c[u][0], c[u][1], c[u][2], c[u][3]);
            } else {
                for (int i = 0; i < n[u]; ++i) {
                    const uint8_t *q8 = fr + (p0[u] + i) * 3;
                    c[u][i] = mk[p0[u] + i] ? (ST_HOLE | ST_ZERO) : (q8[0] | (q8[1] << 8) | ((uint32_t)q8[2] << 16));
                    dst[p0[u] + i] = c[u][i];
                }
            }
        }
        if (!listed) continue;                                 // block-uniform
        uint32_t holes[K4_PACK_UNROLL];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < K4_PACK_UNROLL; ++u) {
            holes[u] = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) holes[u] |= (uint32_t)(i < n[u] && (c[u][i] & ST_HOLE)) << i;
            cnt += __popc(holes[u]);
        }
        if (__ballot_sync(0xffffffffu, cnt != 0)) {            // warp-uniform
            // one scan and one shared-memory atomic per warp and round
            int pre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += v;
            }
            uint32_t base = 0;
            if (lane == 31) base = atomicAdd(&q.count, (uint32_t)pre);
            base = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(pre - cnt);
#pragma unroll
            for (int u = 0; u < K4_PACK_UNROLL; ++u) {
                if (holes[u]) {
                    if (VEC) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if ((holes[u] >> i) & 1u) q.xy[base++] = xy0[u] + i;
                    } else {
                        const uint32_t p32 = (uint32_t)p0[u];      // h*w < 2^32 (both <= 65535)
                        const uint32_t y0 = p32 / (uint32_t)w, x0 = p32 - y0 * (uint32_t)w;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if ((holes[u] >> i) & 1u) {
                                uint32_t x = x0 + i, y = y0;
                                while (x >= (uint32_t)w) x -= w, ++y;   // a group may straddle a row end when w % 4 != 0
                                q.xy[base++] = x | (y << 16);
                            }
                        }
                    }
                }
            }
        }
        queue_flush(q, dl, of, npx, h, w, pflow, pcheck, false);          // same trip count for every thread of the block
    }
    if (listed) queue_flush(q, dl, of, npx, h, w, pflow, pcheck, true);
}

// ---- k4_step_lean: one time step of one direction, in place, over the hole lists ----------------
// blockIdx.y = sub-video.  The state frames hold the input after k4_pack, the backward result after
// pass 1 and the forward result after pass 2:
//   PASS2 == false (backward, t = len-2 .. 0):  frame idx is updated from frame idx+1; holes that
//                                               stay holes are appended to the frame's forward list
//   PASS2 == true  (forward,  t = 1 .. len-1):  frame idx (backward result) is updated from frame
//                                               idx-1 (already the forward result)
// Frame len-1 / frame 0 are the first step of their pass and stay as they are.
//
// The chain of 2*(len-1) dependent launches is latency bound, so the per-item dependency chain is
// kept short and the forward pass only visits what the backward pass left.  With programmatic
// dependent launch the next step's CTAs are resident before this one retires.
// Every frame / list pointer of the step is resolved on the HOST and arrives in the kernel-parameter
// constant bank (one StepWin per window): the device code has no 64-bit frame-offset arithmetic left,
// which matters because the step is bound by instruction issue as much as by its DRAM gathers (ncu:
// ~500 warp instructions per 32 hole pixels, issue active 70 %, before this).
// The re-list queue holds four rounds, so the block-wide flush (2 barriers) runs every fourth round.
struct StepWin {
    const uint32_t *prev;        // state of the frame propagated FROM;  NULL = this window has no such step
    uint32_t *cur;               // state of the frame propagated INTO (updated in place)
    const float2 *flow_check;    // flow used by the forward/backward consistency check
    const float2 *next_flow;     // backward pass: forward-pass flow into this frame; NULL = do not re-list
    const uint32_t *lxy;         // list walked by this step
    const float2 *lflow;
    const uint32_t *count;
    uint32_t *oxy;               // backward pass: forward list of this frame (append)
    float2 *oflow;
    uint32_t *ocount;
};
struct StepArgs {
    StepWin win[K4_MAX_SUB];
};
constexpr int K4_LEAN_ROUNDS = 4;
struct LeanQueue {
    uint32_t xy[K4_LEAN_ROUNDS * K4_BLOCK];
    float2 flow[K4_LEAN_ROUNDS * K4_BLOCK];
    uint32_t count, base;
};
// All threads of the block must call.
__device__ __forceinline__ void lean_flush(LeanQueue &q, const StepWin &sw) {
    __syncthreads();
    const uint32_t n = q.count;
    if (n) {
        if (threadIdx.x == 0) q.base = atomicAdd(sw.ocount, n);
        __syncthreads();
        const uint32_t dst = q.base;
        for (uint32_t j = threadIdx.x; j < n; j += K4_BLOCK) {
            sw.oxy[dst + j] = q.xy[j];
            sw.oflow[dst + j] = q.flow[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) q.count = 0;
    }
    __syncthreads();
}

// NPT = hole pixels per thread and trip: their list entries and taps are independent, so NPT = 2 doubles
// the loads a thread keeps in flight at the price of registers (fewer resident CTAs).
// PRE (backward pass with "k4_precheck"): the list entry holds the un-normalised sample position and, in bit 31
// of the position word, the verdict of the consistency check (both from k4_pack); the step only gathers the four
// state taps.  Entries that failed the check go straight to the forward list.
template <bool PASS2, int MIN_CTAS, bool PRE, int NPT>
__global__ void __launch_bounds__(K4_BLOCK, MIN_CTAS)
    k4_step_lean(const __grid_constant__ StepArgs args, int h, int w, int speculate) {
    asm volatile("griddepcontrol.launch_dependents;");      // let the next step's CTAs get scheduled early
    const StepWin &sw = args.win[blockIdx.y];
    if (sw.prev == nullptr) return;                          // block-uniform
    __shared__ LeanQueue q;
    const bool relist = !PASS2 && sw.next_flow != nullptr;   // holes that stay holes go to the forward list
    if (!PASS2) {
        if (threadIdx.x == 0) q.count = 0;
        __syncthreads();
    }
    const uint32_t stride = gridDim.x * K4_BLOCK;            // entries per round; a trip is NPT rounds
    const uint32_t first = blockIdx.x * K4_BLOCK + (threadIdx.x & ~31u);      // warp-uniform loop bounds
    const uint32_t lane = threadIdx.x & 31u;
    // The backward pass reads what k4_pack wrote (complete before the first step was launched), so it may
    // load its first entries before the grid dependency resolves; the forward lists come from the preceding
    // launches.
    uint32_t xy[NPT], n = 0;
    float2 f[NPT];
#pragma unroll
    for (int u = 0; u < NPT; ++u) xy[u] = 0, f[u] = make_float2(0.f, 0.f);
    if (!PASS2) {
        n = *sw.count;
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const uint32_t i = first + lane + u * stride;
            if (i < n) xy[u] = sw.lxy[i], f[u] = sw.lflow[i];
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (PASS2) {
        n = *sw.count;
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const uint32_t i = first + lane + u * stride;
            if (i < n) xy[u] = sw.lxy[i], f[u] = sw.lflow[i];
        }
    }
    // backward pass: block-uniform trip count (the queue flush has barriers)
    const uint32_t n_loop = PASS2 ? n : min(n + (K4_BLOCK - 1), 0xffffff00u) / K4_BLOCK * K4_BLOCK;
    int trip = 0;
    for (uint32_t base = first; base < n_loop; base += NPT * stride, ++trip) {
        uint32_t xy_cur[NPT], nv[NPT], pix[NPT];
        float2 f_cur[NPT], nf[NPT];
        bool valid[NPT];
        // software pipeline: the entries of the NEXT trip are requested before this trip's taps, so a thread's
        // dependency chain is one round trip per item (plus one) instead of two
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const uint32_t i = base + lane + u * stride;
            valid[u] = i < n;
            xy_cur[u] = xy[u], f_cur[u] = f[u];
            const uint32_t inext = i + NPT * stride;
            if (inext < n) xy[u] = sw.lxy[inext], f[u] = sw.lflow[inext];
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const int x = (int)(xy_cur[u] & 0xffffu), y = (int)((xy_cur[u] >> 16) & 0x7fffu);
            pix[u] = (uint32_t)y * (uint32_t)w + (uint32_t)x;
            nv[u] = ST_HOLE | ST_ZERO;
            nf[u] = make_float2(0.f, 0.f);
            if (valid[u]) {
                // the forward-pass flow is fetched speculatively, in the same round trip as the taps
                if (relist && speculate) nf[u] = __ldg(sw.next_flow + pix[u]);
                if (PRE) {
                    if (!(xy_cur[u] & XY_INVALID))
                    {
                        const Taps tp = make_taps(f_cur[u].x, f_cur[u].y, h, w);
                        nv[u] = fill_from_prev(tp, ST_HOLE | ST_ZERO, load_state_taps(tp, sw.prev, w));
                    }
                } else {
                    nv[u] = propagate_pixel(x, y, h, w, ST_HOLE | ST_ZERO, f_cur[u], sw.flow_check, sw.prev);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            if (valid[u]) {
                if (nv[u] != (ST_HOLE | ST_ZERO)) sw.cur[pix[u]] = nv[u];
                else if (relist && !speculate) nf[u] = __ldg(sw.next_flow + pix[u]);
            }
            if (relist) {                   // still a hole: the forward pass gets another chance
                const bool take = valid[u] && nv[u] == (ST_HOLE | ST_ZERO);
                const uint32_t m = __ballot_sync(0xffffffffu, take);
                if (m) {
                    uint32_t at = 0;
                    if (lane == 0) at = atomicAdd(&q.count, (uint32_t)__popc(m));
                    at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
                    if (take) q.xy[at] = xy_cur[u] & ~XY_INVALID, q.flow[at] = nf[u];
                }
            }
        }
        constexpr int FLUSH_EVERY = K4_LEAN_ROUNDS / NPT;
        if (relist && (trip % FLUSH_EVERY) == FLUSH_EVERY - 1) lean_flush(q, sw);
    }
    if (relist) lean_flush(q, sw);
}


// ---- k4_steps_persistent: the whole serial scan (both passes, every time step of every window) in ONE launch ----
// The 2 * (window - 1) steps are dependent, and as separate launches each of them pays the turnaround of a dependent
// kernel (grid drain, completion, launch, ramp-up: a few microseconds even with programmatic dependent launch) for
// 5 - 10 us of latency-bound gathers.  Here a resident grid (cooperative launch: every CTA is on an SM) walks the
// steps itself and meets at a grid-wide barrier in between - one atomic per CTA on a counter in L2 and an acquire
// spin.  Everything another CTA wrote earlier in the launch (state frames, forward hole lists, their counters) is
// read from L2 (ld.global.cg), the flows (read-only for the whole launch) through the read-only path.  The step
// arithmetic is k4_step_lean's.  blockIdx.y = sub-video window, as there.
// MEASURED (B200, 240 frames 960x540): 1.59 ms against 1.36 ms for the chain of launches - 740 same-address arrivals
// and an L2 polling round trip per barrier cost more than the turnaround programmatic dependent launch leaves, and
// the launch chain already loads a step's first list entries before the previous step has retired.  Option
// k4_persist = 1, off by default; kept for A/B runs and tested for parity.
struct PersistArgs {
    SubBatch b;
    uint32_t *out, *pads;
    const float2 *ff, *fb;
    HoleLists l1, l2;
    long long npx;
    int h, w, blen, speculate;
    uint32_t *barrier;           // zeroed before the launch
};
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// All threads of all CTAs call; `target` = arrivals expected so far (monotonic counter, never reset).
__device__ __forceinline__ void grid_barrier(uint32_t *bar, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                     // this CTA's writes are visible before its arrival
        atomicAdd(bar, 1u);
        while ((int32_t)(ld_acquire_gpu(bar) - target) < 0) {
        }
    }
    __syncthreads();
}

template <bool PASS2>
__device__ __forceinline__ void persistent_step(const StepWin &sw, LeanQueue &q, int h, int w, int speculate) {
    const bool relist = !PASS2 && sw.next_flow != nullptr;   // holes that stay holes go to the forward list
    if (!PASS2) {
        if (threadIdx.x == 0) q.count = 0;
        __syncthreads();
    }
    const uint32_t stride = gridDim.x * K4_BLOCK;
    const uint32_t first = blockIdx.x * K4_BLOCK + (threadIdx.x & ~31u);      // warp-uniform loop bounds
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = __ldcg(sw.count);
    uint32_t xy = 0;
    float2 f = make_float2(0.f, 0.f);
    {
        const uint32_t i = first + lane;
        if (i < n) xy = __ldcg(sw.lxy + i), f = __ldcg(sw.lflow + i);
    }
    // backward pass: block-uniform trip count (the queue flush has barriers)
    const uint32_t n_loop = PASS2 ? n : min(n + (K4_BLOCK - 1), 0xffffff00u) / K4_BLOCK * K4_BLOCK;
    int trip = 0;
    for (uint32_t base = first; base < n_loop; base += stride, ++trip) {
        const uint32_t i = base + lane;
        const bool valid = i < n;
        const uint32_t xy_cur = xy;
        const float2 f_cur = f;
        const uint32_t inext = i + stride;                   // the next trip's entry, requested before this trip's taps
        if (inext < n) xy = __ldcg(sw.lxy + inext), f = __ldcg(sw.lflow + inext);
        const int x = (int)(xy_cur & 0xffffu), y = (int)((xy_cur >> 16) & 0x7fffu);
        const uint32_t pix = (uint32_t)y * (uint32_t)w + (uint32_t)x;
        uint32_t nv = ST_HOLE | ST_ZERO;
        float2 nf = make_float2(0.f, 0.f);
        if (valid) {
            if (relist && speculate) nf = __ldg(sw.next_flow + pix);
            nv = propagate_pixel<true>(x, y, h, w, ST_HOLE | ST_ZERO, f_cur, sw.flow_check, sw.prev);
            if (nv != (ST_HOLE | ST_ZERO)) sw.cur[pix] = nv;
            else if (relist && !speculate) nf = __ldg(sw.next_flow + pix);
        }
        if (relist) {                       // still a hole: the forward pass gets another chance
            const bool take = valid && nv == (ST_HOLE | ST_ZERO);
            const uint32_t m = __ballot_sync(0xffffffffu, take);
            if (m) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(&q.count, (uint32_t)__popc(m));
                at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
                if (take) q.xy[at] = xy_cur & ~XY_INVALID, q.flow[at] = nf;
            }
            if ((trip % K4_LEAN_ROUNDS) == K4_LEAN_ROUNDS - 1) lean_flush(q, sw);
        }
    }
    if (relist) lean_flush(q, sw);
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(K4_BLOCK, MIN_CTAS) k4_steps_persistent(const __grid_constant__ PersistArgs a) {
    __shared__ LeanQueue q;
    __shared__ StepWin s_sw;
    const SubDesc &sd = a.b.sub[blockIdx.y];
    const uint32_t n_ctas = gridDim.x * gridDim.y;
    uint32_t arrivals = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int step = 1; step < a.blen; ++step) {
            const bool active = step < sd.len;               // block-uniform: shorter windows only keep the barrier
            if (active) {
                if (threadIdx.x == 0) {
                    StepWin sw = StepWin();
                    const int idx = pass ? step : sd.len - 1 - step;
                    const long long gframe = sd.start + idx, of = sd.out_frame + idx;
                    const HoleLists &li = pass ? a.l2 : a.l1;
                    sw.cur = frame_ptr(sd, idx, a.out, a.pads, a.npx);
                    sw.prev = frame_ptr(sd, pass ? idx - 1 : idx + 1, a.out, a.pads, a.npx);
                    sw.flow_check = pass ? a.ff + (gframe - 1) * a.npx : a.fb + gframe * a.npx;
                    sw.next_flow = (!pass && idx >= 1) ? a.fb + (gframe - 1) * a.npx : nullptr;
                    sw.lxy = li.xy + of * a.npx, sw.lflow = li.flow + of * a.npx, sw.count = li.count + of;
                    sw.oxy = a.l2.xy + of * a.npx, sw.oflow = a.l2.flow + of * a.npx, sw.ocount = a.l2.count + of;
                    s_sw = sw;
                }
                __syncthreads();
                if (pass)
                    persistent_step<true>(s_sw, q, a.h, a.w, a.speculate);
                else
                    persistent_step<false>(s_sw, q, a.h, a.w, a.speculate);
            }
            arrivals += n_ctas;
            grid_barrier(a.barrier, arrivals);
        }
}

__global__ void __launch_bounds__(256)
    k4_unpack(const uint32_t *__restrict__ packed, long long n, uint8_t zero_level, uint8_t *__restrict__ rgb,
              uint8_t *__restrict__ hole) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t v = packed[i];
        if (rgb) {
            const bool z = v & ST_ZERO;
            rgb[3 * i] = z ? zero_level : (uint8_t)v;
            rgb[3 * i + 1] = z ? zero_level : (uint8_t)(v >> 8);
            rgb[3 * i + 2] = z ? zero_level : (uint8_t)(v >> 16);
        }
        if (hole) hole[i] = (v & ST_HOLE) ? 255 : 0;
    }
}

// Side streams of "k4_streams" (one set per device, created on first use; the mutex keeps the event record -> wait
// pairs of concurrent vv_propagate calls apart - sharing the streams between calls only serialises their side chains).
constexpr int K4_MAX_STREAMS = 4;
struct SideStreams {
    std::mutex mu;
    cudaStream_t s[K4_MAX_STREAMS - 1];
    cudaEvent_t fork, join[K4_MAX_STREAMS - 1];
    bool ready = false;
};
static SideStreams *side_streams() {
    static SideStreams pool[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStreams &p = pool[dev];
    std::lock_guard<std::mutex> lk(p.mu);
    if (!p.ready) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int i = 0; i < K4_MAX_STREAMS - 1; ++i) {
            if (cudaStreamCreateWithPriority(&p.s[i], cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        p.ready = true;
    }
    return &p;
}

}  // namespace vv

using namespace vv;

extern "C" size_t vv_propagate_workspace_bytes(int n_window_frames, int n_pad_frames, int h, int w) {
    if (n_window_frames <= 0 || h <= 0 || w <= 0) return 0;
    if (n_pad_frames < 0) n_pad_frames = 0;
    // two hole-list sets, worst case one entry per pixel: u32 position + float2 flow; counters per frame;
    // packed state of the pad frames
    const size_t n = (size_t)n_window_frames * h * w;
    return 2 * (align_up(n * 4, 256) + align_up(n * 8, 256)) + align_up((size_t)n_window_frames * 8, 256) +
           align_up((size_t)n_pad_frames * h * w * 4, 256) + 256;
}

extern "C" int vv_propagate(const uint8_t *frames, const uint8_t *masks, const float *flows_f, const float *flows_b,
                            int n_frames, int h, int w, const int *sub_start, const int *sub_len,
                            const int *sub_keep_start, const int *sub_keep_len, int n_sub, uint32_t *out,
                            void *workspace, size_t workspace_bytes, void *stream) {
    VV_CHECK_ARG(frames && masks && out && workspace && sub_start && sub_len, "vv_propagate: NULL pointer");
    VV_CHECK_ARG((sub_keep_start == nullptr) == (sub_keep_len == nullptr),
                 "vv_propagate: sub_keep_start and sub_keep_len must be given together");
    VV_CHECK_ARG(n_frames > 0 && h > 0 && w > 0 && n_sub > 0, "vv_propagate: bad shape");
    VV_CHECK_ARG(h <= 32767 && w <= 65535, "vv_propagate: frame too large (h <= 32767, w <= 65535)");
    VV_CHECK_ARG(n_frames == 1 || (flows_f && flows_b), "vv_propagate: flows required when there is more than one frame");
    long long total = 0, kept = 0;
    for (int s = 0; s < n_sub; ++s) {
        VV_CHECK_ARG(sub_len[s] > 0 && sub_start[s] >= 0 && sub_start[s] + sub_len[s] <= n_frames,
                     "vv_propagate: sub-video %d [%d,+%d) outside the clip of %d frames", s, sub_start[s], sub_len[s],
                     n_frames);
        if (sub_keep_start)
            VV_CHECK_ARG(sub_keep_start[s] >= 0 && sub_keep_len[s] >= 0 && sub_keep_start[s] + sub_keep_len[s] <= sub_len[s],
                         "vv_propagate: kept range [%d,+%d) outside sub-video %d of %d frames", sub_keep_start[s],
                         sub_keep_len[s], s, sub_len[s]);
        total += sub_len[s];
        kept += sub_keep_start ? sub_keep_len[s] : sub_len[s];
    }
    VV_CHECK_ARG(total < (1LL << 31), "vv_propagate: too many frames");
    VV_CHECK_ARG(workspace_bytes >= vv_propagate_workspace_bytes((int)total, (int)(total - kept), h, w),
                 "vv_propagate: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long npx = (long long)h * w;
    const bool vec = (w % 4 == 0) && ((uintptr_t)frames % 4 == 0) && ((uintptr_t)masks % 4 == 0) &&
                     ((uintptr_t)out % 16 == 0);
    const size_t l4 = align_up((size_t)total * npx * 4, 256), l8 = align_up((size_t)total * npx * 8, 256);
    uint8_t *wsp = (uint8_t *)workspace;
    HoleLists l1, l2;
    l1.xy = (uint32_t *)wsp, l1.flow = (float2 *)(wsp + l4);
    l2.xy = (uint32_t *)(wsp + l4 + l8), l2.flow = (float2 *)(wsp + 2 * l4 + l8);
    l1.count = (uint32_t *)(wsp + 2 * (l4 + l8));
    l2.count = l1.count + total;
    uint32_t *pads = (uint32_t *)(wsp + 2 * (l4 + l8) + align_up((size_t)total * 8, 256));
    // the grid barrier of k4_steps_persistent lives in the 256 spare bytes behind the pad frames
    uint32_t *barrier = (uint32_t *)((uint8_t *)pads + align_up((size_t)(total - kept) * npx * 4, 256));
    const float2 *ff = (const float2 *)flows_f, *fb = (const float2 *)flows_b;
    cudaError_t e = cudaMemsetAsync(l1.count, 0, (size_t)total * 8, st);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");

    long long out_frame = 0, keep_out = 0, pad_out = 0;
    for (int base = 0; base < n_sub; base += K4_MAX_SUB) {
        SubBatch b;
        b.n = min(K4_MAX_SUB, n_sub - base);
        int blen = 0;
        const long long first = out_frame;
        for (int s = 0; s < b.n; ++s) {
            SubDesc &sd = b.sub[s];
            sd.start = sub_start[base + s];
            sd.len = sub_len[base + s];
            sd.pad_s = sub_keep_start ? sub_keep_start[base + s] : 0;
            sd.keep = sub_keep_start ? sub_keep_len[base + s] : sd.len;
            sd.out_frame = out_frame, sd.keep_out = keep_out, sd.pad_out = pad_out;
            out_frame += sd.len, keep_out += sd.keep, pad_out += sd.len - sd.keep;
            blen = max(blen, sd.len);
        }
        const long long bframes = out_frame - first;
        const int pre = get_option(OPT_K4_TAPS) != 0;          // "k4_precheck"
        // pack: all frames of the batch at once (grid.y <= 65535 frames per launch)
        for (long long f0 = 0; f0 < bframes; f0 += 32768) {
            const int ny = (int)min(32768LL, bframes - f0);
            const int ctas = 148 * max(1, min(1024, get_option(OPT_K4_PACK_CTAS)));
            const int gx = max(1, min(ceil_div((npx + 3) / 4, 256 * K4_PACK_UNROLL), ceil_div(ctas, ny)));
            const int occ = get_option(OPT_K4_PACK_OCC);
#define VV_K4_PACK(V, O) \
    k4_pack<V, O><<<dim3(gx, ny), 256, 0, st>>>(frames, masks, out, pads, ff, fb, l1, l2, h, w, first + f0, pre, b)
            if (!vec)
                VV_K4_PACK(false, 4);
            else if (occ >= 6)
                VV_K4_PACK(true, 6);
            else if (occ == 5)
                VV_K4_PACK(true, 5);
            else
                VV_K4_PACK(true, 4);
#undef VV_K4_PACK
            VV_POST_LAUNCH("k4_pack");
        }
        // "k4_streams" = G > 1: the windows of the batch are dealt to G groups (window s -> group s % G) and every
        // group runs its own chain of dependent step launches on its own stream (group 0 on the caller's, the others
        // on internal side streams of the same priority, forked behind k4_pack and joined at the end).  A step is
        // partly latency bound (launch turnaround, then count -> list entry -> taps, three dependent round trips, for a
        // few trips per thread), so chains of independent windows fill each other's bubbles.  Measured and dropped:
        // side streams of higher priority (their pre-launched, waiting CTAs keep the other chain off the SMs: 1.40 ->
        // 1.9 ms) and starting chain g underneath the pack of group g + 1 (two pack launches: 1.40 -> 1.48 ms).
        const bool persist = get_option(OPT_K4_PERSIST) != 0 && blen > 1 && !pre;
        const int groups = (persist || blen <= 1) ? 1 : max(1, min(min(get_option(OPT_K4_STREAMS), K4_MAX_STREAMS), b.n));
        SideStreams *side = nullptr;
        std::unique_lock<std::mutex> side_lock;
        struct SideGuard {           // an error exit must not leave side-stream work running on buffers the caller frees
            SideStreams *side = nullptr;
            int n = 0;
            ~SideGuard() {
                for (int i = 0; side != nullptr && i < n; ++i) cudaStreamSynchronize(side->s[i]);
            }
        } guard;
        if (groups > 1) {
            side = side_streams();
            if (side == nullptr) return fail_cuda(cudaGetLastError(), "vv_propagate: side streams");
            side_lock = std::unique_lock<std::mutex>(side->mu);       // record -> wait pairs of one call stay together
            guard.side = side, guard.n = groups - 1;                  // error exits wait for what the side chains got
            if ((e = cudaEventRecord(side->fork, st)) != cudaSuccess) return fail_cuda(e, "cudaEventRecord(fork)");
            for (int g = 1; g < groups; ++g)
                if ((e = cudaStreamWaitEvent(side->s[g - 1], side->fork, 0)) != cudaSuccess)
                    return fail_cuda(e, "cudaStreamWaitEvent(fork)");
        }
        // The serial scans touch hole pixels only.  The hole counts live on the device: "k4_step_ctas" > 0 =
        // that many CTAs per SM in total (a resident grid that strides over the lists), 0 = about one thread
        // per hole at a 25 % hole fraction.
        dim3 grid(max(1, min(ceil_div(npx / 4, 256), ceil_div(148 * 16, b.n))), b.n);
        const int per_sm = get_option(OPT_K4_STEP_CTAS);
        if (per_sm > 0) grid.x = max(1, min(ceil_div(npx / 4, 256), (148 * per_sm) / b.n));
        if (persist) {
            // one cooperative launch for the whole scan; the grid must be resident (occupancy x SMs)
            static std::atomic<int> occ_cache[64];
            int dev_id = 0, sms = 148;
            cudaGetDevice(&dev_id);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
            dev_id = min(max(dev_id, 0), 63);
            int occp = occ_cache[dev_id].load();
            if (occp == 0) {
                cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, k4_steps_persistent<5>, K4_BLOCK, 0);
                if (oe != cudaSuccess || occp < 1) return fail_cuda(oe, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(k4_steps_persistent)");
                occ_cache[dev_id].store(occp);
            }
            const int want = per_sm > 0 ? min(per_sm, occp) : occp;
            dim3 pgrid(max(1, min(ceil_div(npx / 4, 256), (sms * want) / b.n)), b.n);
            PersistArgs pa;
            pa.b = b, pa.out = out, pa.pads = pads, pa.ff = ff, pa.fb = fb, pa.l1 = l1, pa.l2 = l2, pa.npx = npx;
            pa.h = h, pa.w = w, pa.blen = blen, pa.speculate = get_option(OPT_K4_SPECULATE), pa.barrier = barrier;
            if ((e = cudaMemsetAsync(barrier, 0, 16, st)) != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(barrier)");
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = pgrid;
            cfg.blockDim = dim3(K4_BLOCK);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t le = cudaLaunchKernelEx(&cfg, k4_steps_persistent<5>, pa);
            if (le != cudaSuccess) return fail_cuda(le, "cudaLaunchKernelEx(k4_steps_persistent)");
            VV_POST_LAUNCH("k4_steps_persistent");
            continue;
        }
        const int pdl = get_option(OPT_K4_PDL) != 0;
        const int lean = get_option(OPT_K4_LEAN);
        const int spec = get_option(OPT_K4_SPECULATE);
        const int npt = get_option(OPT_K4_NPT) >= 2 ? 2 : 1;
        for (int pass = 0; pass < 2; ++pass)
            for (int step = 1; step < blen; ++step)
                for (int g = 0; g < groups; ++g) {
                    StepArgs sa;
                    int ng = 0;                                        // windows of this group that take this step
                    for (int s = g; s < b.n; s += groups) {
                        if (step >= b.sub[s].len) continue;
                        StepWin &sw = sa.win[ng++];
                        const SubDesc &sd = b.sub[s];
                        const int idx = pass ? step : sd.len - 1 - step;
                        const long long gframe = sd.start + idx, of = sd.out_frame + idx;
                        const HoleLists &li = pass ? l2 : l1;
                        sw.cur = frame_ptr(sd, idx, out, pads, npx);
                        sw.prev = frame_ptr(sd, pass ? idx - 1 : idx + 1, out, pads, npx);
                        // check flow: backward pass flows_b[idx] (prop = flows_f[idx]), forward pass flows_f[idx-1]
                        sw.flow_check = pass ? ff + (gframe - 1) * npx : fb + gframe * npx;
                        sw.next_flow = (!pass && idx >= 1) ? fb + (gframe - 1) * npx : nullptr;
                        sw.lxy = li.xy + of * npx, sw.lflow = li.flow + of * npx, sw.count = li.count + of;
                        sw.oxy = l2.xy + of * npx, sw.oflow = l2.flow + of * npx, sw.ocount = l2.count + of;
                    }
                    if (ng == 0) continue;
                    for (int s = ng; s < K4_MAX_SUB; ++s) sa.win[s] = StepWin();
                    dim3 ggrid(grid.x, ng);
                    // several chains share the SMs: "k4_chain_ctas" CTAs per SM for all of them together
                    const int budget = groups > 1 ? max(1, get_option(OPT_K4_CHAIN_CTAS)) : per_sm;
                    if (budget > 0) ggrid.x = max(1, min(ceil_div(npx / 4, 256), (148 * budget) / (groups * ng)));
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = ggrid;
                    cfg.blockDim = dim3(K4_BLOCK);
                    cfg.stream = g == 0 ? st : side->s[g - 1];
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    attr[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = attr;
                    // the first step after k4_pack (or after the fork) is an ordinary launch: backward steps read the
                    // lists k4_pack wrote BEFORE their griddepcontrol.wait, and only a full stream dependency makes
                    // those visible
                    cfg.numAttrs = (pdl && !(pass == 0 && step == 1)) ? 1 : 0;
                    cudaError_t le;
                    // backward pass with precheck: PRE kernel; forward pass always carries flows
#define VV_K4_LEAN(O, N)                                                                              \
    (pass == 0 ? (pre ? cudaLaunchKernelEx(&cfg, k4_step_lean<false, O, true, N>, sa, h, w, spec)      \
                      : cudaLaunchKernelEx(&cfg, k4_step_lean<false, O, false, N>, sa, h, w, spec))    \
               : cudaLaunchKernelEx(&cfg, k4_step_lean<true, O, false, N>, sa, h, w, spec))
                    if (npt == 2)
                        le = lean == 4 ? VV_K4_LEAN(4, 2) : VV_K4_LEAN(3, 2);
                    else if (lean >= 8)
                        le = VV_K4_LEAN(8, 1);
                    else if (lean >= 6)
                        le = VV_K4_LEAN(6, 1);
                    else
                        le = VV_K4_LEAN(5, 1);
#undef VV_K4_LEAN
                    if (le != cudaSuccess) return fail_cuda(le, "cudaLaunchKernelEx(k4_step_lean)");
                    VV_POST_LAUNCH("k4_step_lean");
                }
        if (side != nullptr)                                           // join: the caller's stream waits for every chain
            for (int g = 1; g < groups; ++g) {
                if ((e = cudaEventRecord(side->join[g - 1], side->s[g - 1])) != cudaSuccess) return fail_cuda(e, "cudaEventRecord(join)");
                if ((e = cudaStreamWaitEvent(st, side->join[g - 1], 0)) != cudaSuccess) return fail_cuda(e, "cudaStreamWaitEvent(join)");
            }
        guard.side = nullptr;                                          // joined on the caller's stream: nothing to wait for
    }
    return VV_OK;
}

extern "C" int vv_propagate_unpack(const uint32_t *packed, size_t n_pixels, uint8_t zero_level, uint8_t *rgb,
                                   uint8_t *hole_mask, void *stream) {
    VV_CHECK_ARG(packed && n_pixels > 0 && (rgb || hole_mask), "vv_propagate_unpack: bad argument");
    const int grid = (int)min((long long)ceil_div((long long)n_pixels, 256), (long long)148 * 32);
    k4_unpack<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, (long long)n_pixels, zero_level, rgb, hole_mask);
    VV_POST_LAUNCH("k4_unpack");
    return VV_OK;
}

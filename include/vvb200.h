/*
 * vvb200.h - C ABI of libvvb200.so: the B200 (sm_100a) pixel pipeline that sits behind
 * VideoVanish's `diffuerase.run_infill_on_frames` (reference: diffuerase.py:20-114).
 *
 * Every entry point takes plain pointers and sizes; there are no torch / C++ types in
 * the signatures.  `*_dev` style arguments are DEVICE pointers into contiguous NHWC
 * ("[T,H,W,C]", C fastest) uint8 arrays unless stated otherwise; `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = the legacy default stream).  All functions
 * are asynchronous with respect to the host (they only enqueue work on `stream`) except
 * the `*_host` family, which owns its staging buffers and returns when the result is in
 * the caller's host memory.
 *
 * Return value: 0 on success, a negative VV_ERR_* code on failure; `vv_last_error()`
 * returns a thread-local, human-readable message for the last failure on this thread.
 * There is no CPU fallback anywhere: without a CUDA device every call fails with
 * VV_ERR_CUDA.
 *
 * The reference is pure Python; the binding a maintainer adds on the reference side is a
 * `ctypes.CDLL` stub.  Each declaration cites the reference lines
 * it replaces (paths relative to /root/reference).
 */
#ifndef VVB200_H_
#define VVB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VV_API __attribute__((visibility("default")))
#else
#define VV_API
#endif

#define VV_OK 0
#define VV_ERR_INVALID (-1)     /* bad argument (NULL pointer, non-positive size, ...) */
#define VV_ERR_CUDA (-2)        /* CUDA runtime error / no device */
#define VV_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels implement */

#define VV_INTER_NEAREST 0 /* cv2.INTER_NEAREST semantics (diffuerase.py:86, tools.py:42) */
#define VV_INTER_LINEAR 1  /* cv2.INTER_LINEAR u8 fixed-point semantics (diffuerase.py:73) */

/* Library / device introspection. */
VV_API int vv_version(void);
VV_API const char *vv_last_error(void);
VV_API int vv_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
/* Number of kernels this library has launched since load (or since the last reset);
 * bench.py reports it as `gpu_launches`. */
VV_API unsigned long long vv_launch_count(void);
VV_API void vv_reset_launch_count(void);
/* Kernel-variant switches used for A/B measurements (defaults are the tuned choices):
 *   "k1b_exact"  1 = fully unrolled 8-round dilation for the default radius, 0 = generic loop.
 *   "k1b_diag"   1 = a dilation pass of radius 9..16 starts with the diamond of radius 2K (K = 4..7) built from two
 *                diagonal segments by doubling (9 shift-OR steps for up to 14 cross rounds); 2 (default) = also the
 *                unrolled radius-8 pass (radius-6 diamond in 7 steps + two cross rounds); 0 = cross rounds only.
 *   "k3_nt"      16-pixel groups per thread and iteration in K3 (1 or 2).
 *   "k3_tma"     1 = K3 stages the original strip through shared memory with bulk async copies
 *                (TMA) when the frame is 16-byte aligned, 0 = register pass-through kernel.
 *   "k3_tma_rows" maximum rows per staged strip (2..16);  "k3_tma_threads" 256, 384 or 512.
 *   "k4_pdl"     1 = propagation steps use programmatic dependent launch (multi-launch mode).
 *   "k4_npt"     hole pixels per thread and trip of a propagation step: 1 (default) or 2 (twice the loads in
 *                flight per thread, fewer resident CTAs).
 *   "k3_bits"    1 (default) = K3 reads the dilated 1-bit mask plane K1 left in its workspace when the caller
 *                passes it (vv_upscale_feather_composite_bits); 0 = always rebuild the bit rows from the u8 mask.
 *   "k3_x2"      3 (default) = the k3_fastw kernel whenever W0 == 2w or 4w (W0 <= 4096): closed-form horizontal pass,
 *                closed-form or table-driven vertical pass, 32-pixel word classification tasks, interior words without
 *                blend, software-pipelined worker, packed-fp32 blend; 2 = k3_fast, its predecessor (16-pixel rolling
 *                tasks); 1 = the round-1 closed-form worker (needs H0 == 2h as well); 0 = the generic tap-table worker.
 *   "k3_big_from" smallest feather window radius (ceil(feather_px) - 1) that runs k3_bigfeather, the table-walking kernel of
 *                feather_px in (8, 32]; default 8, 3..7 send those radii there as well instead of to the round-1 generic kernel.
 *   "k3_chain"   1 = the next k3_fastw launch is chained to the K3 launch before it in the stream (programmatic stream
 *                serialisation; it may start while that one drains).  Only for back-to-back K3 calls over DIFFERENT
 *                frames of a clip; the Python front end sets it per call (chain_previous=True).  Default 0.
 *   "k4_pack_ctas" k4_pack launches about 148 x this many CTAs per call (more, shorter CTAs shrink the tail
 *                of the last wave; default 128);  "k4_pack_occ" 4 (default), 5 or 6 = CTAs per SM the kernel is
 *                compiled for.
 *   "k4_lean"    5 (default), 6 or 8 = CTAs per SM the propagation step kernel is compiled for.
 *   "k4_step_ctas" CTAs per SM of the step grid (default 5: a resident grid that strides over the hole lists
 *                with the next entry pre-loaded); 0 = about one thread per hole at a 25 % hole fraction.
 *   "k4_precheck" 1 = k4_pack decides the forward / backward flow-consistency check of the backward pass for every
 *                hole of every frame at once and stores the un-normalised sample position in the list entry, so that the
 *                serial backward steps only gather the four state taps; 0 (default, measured faster: in the steps the
 *                flow taps ride along with the state taps in one latency-bound round trip, in the bandwidth-bound pack
 *                they cost their full, badly coalesced bytes) = the steps do it all.
 *   "k4_speculate" 1 (default, measured faster) = the backward pass fetches the forward-pass flow of a hole
 *                together with its taps; 0 = only after the hole turned out to stay a hole.
 *   "k4_persist" 1 = the whole propagation scan as ONE persistent cooperative launch with grid-wide barriers instead of
 *                one launch per time step (measured slower on B200: default 0).
 *   "k4_streams" G (default 2, 1..4): the propagation windows of a call are dealt to G groups (window s -> group s % G)
 *                and every group runs its chain of dependent step launches on its own stream (group 0 on the caller's,
 *                the others on internal side streams forked behind k4_pack): independent chains fill each other's
 *                launch / round-trip bubbles.  The caller's stream waits for all of them before vv_propagate's work
 *                counts as done.  1 = one chain on the caller's stream.
 *   "k4_chain_ctas" CTAs per SM of all G > 1 chains together (default 8; "k4_step_ctas" applies to G == 1).
 *   "k5_halo_ctas" CTA budget of vv_halo_blend (default 148 x 8; chunking.produce_and_blend_boundaries lowers it to
 *                128 while the exchange runs on a side stream underneath K3). */
VV_API int vv_set_option(const char *name, int value);
VV_API int vv_get_option(const char *name, int *value);

/* ---------------------------------------------------------------------------------
 * K1  mask binarise + L1 dilation.                       Replaces diffuerase.py:28-31
 *   out[t,y,x] = 255 if some pixel q with any(mask[t,q,:] > 0) has |p-q|_1 <= iterations
 *   (scipy.ndimage.binary_dilation, cross structure, border 0); iterations < 1 means
 *   "until convergence": the whole frame is set if any pixel is set.
 *   mask: u8 [T,H,W,C], C in {1,3,4}.  out: u8 [T,H,W] in {0,255}.
 *   workspace: vv_binarize_dilate_workspace_bytes(T,H,W) bytes of device scratch (bit planes).
 *   lowres_out (optional, may be NULL): u8 [T,lh,lw] = INTER_NEAREST down-size of `out`
 *   written by the same pass (the model-side mask of SURVEY row A9).
 * --------------------------------------------------------------------------------- */
VV_API size_t vv_binarize_dilate_workspace_bytes(int T, int H, int W);
VV_API int vv_binarize_dilate(const uint8_t *mask, int T, int H, int W, int C, int iterations,
                       uint8_t *out, uint8_t *lowres_out, int lh, int lw,
                       void *workspace, size_t workspace_bytes, void *stream);
/* Same, and additionally (bits_out != NULL) the dilated mask as a 1-bit plane: u32 [T,H,ceil(W/32)], pixel x of a
 * row = bit (x & 31) of word (x >> 5), bits at x >= W zero.  The pass computes it anyway; handing it to
 * vv_upscale_feather_composite_bits saves K3 the re-read and re-packing of the u8 mask (stage fusion across
 * diffuerase.py:28-31 and :77-90). */
VV_API int vv_binarize_dilate_ex(const uint8_t *mask, int T, int H, int W, int C, int iterations,
                          uint8_t *out, uint8_t *lowres_out, int lh, int lw, uint32_t *bits_out,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------
 * K2  resize.            Replaces cv2.resize at diffuerase.py:73 / :86, tools.py:42 and the
 *   un-vendored down-size to inference resolution triggered by diffuerase.py:62-64.
 *   src: u8 [T,H,W,C] -> dst: u8 [T,h,w,C]; C in {1,3,4}; interp = VV_INTER_*.
 *   LINEAR is bit-exact with OpenCV's u8 path (11-bit coefficients; exact x2 down-scale
 *   == 2x2 box).  workspace: vv_resize_workspace_bytes(h,w) bytes (tap tables).
 * --------------------------------------------------------------------------------- */
VV_API size_t vv_resize_workspace_bytes(int h, int w);
VV_API int vv_resize(const uint8_t *src, int T, int H, int W, int C, uint8_t *dst, int h, int w,
              int interp, void *workspace, size_t workspace_bytes, void *stream);
/* (h, w) the model wrapper picks for `max_img_size`. */
VV_API int vv_inference_size(int H0, int W0, int max_img_size, int *h, int *w);

/* ---------------------------------------------------------------------------------
 * K3  resize-back + feather alpha + composite, fused.   Replaces diffuerase.py:70-112
 *   inp:  u8 [T,h,w,3]    inpainted frames at inference resolution
 *   orig: u8 [T,H0,W0,3]  original frames        (ignored when keep_unmasked == 0)
 *   mask: u8 [T,H0,W0]    dilated mask, > 0 = masked (ignored when keep_unmasked == 0)
 *   out:  u8 [T,H0,W0,3]
 *   alpha = clip(0.5 + (d_in - d_out) / (2*feather_px), 0, 1) with the 5x5 chamfer
 *   distance transforms of cv2.distanceTransform(DIST_L2, 5); feather_px <= 0 gives the
 *   hard composite; out = u8(rint(f32(alpha*up) + f32((1-alpha)*orig))), round-half-even.
 *   feather_px up to VV_MAX_FEATHER is supported.
 * --------------------------------------------------------------------------------- */
#define VV_MAX_FEATHER 32.0f
VV_API size_t vv_composite_workspace_bytes(int H0, int W0);
VV_API int vv_upscale_feather_composite(const uint8_t *inp, int T, int h, int w,
                                 const uint8_t *orig, const uint8_t *mask, int H0, int W0,
                                 float feather_px, int keep_unmasked, uint8_t *out,
                                 void *workspace, size_t workspace_bytes, void *stream);
/* Same with the mask ALSO given as the 1-bit plane of vv_binarize_dilate_ex (mask_bits may be NULL): kernels
 * that can use it read 1/8 byte per pixel instead of 1; `mask` is still required (other kernels use it) and must
 * describe the same set. */
VV_API int vv_upscale_feather_composite_bits(const uint8_t *inp, int T, int h, int w,
                                 const uint8_t *orig, const uint8_t *mask, const uint32_t *mask_bits,
                                 int H0, int W0, float feather_px, int keep_unmasked, uint8_t *out,
                                 void *workspace, size_t workspace_bytes, void *stream);

/* The chamfer table the feather stage works from (host code, needs no GPU): table[(oy + R) * (2R + 1) + (ox + R)] =
 * what cv2.distanceTransform(DIST_L2, 5) (diffuerase.py:95-96) yields at offset (oy, ox) FROM a single zero pixel - the
 * float32 result of its two raster passes, not symmetric from d ~ 12 on.  0 <= R <= 31; table: (2R + 1)^2 floats. */
VV_API int vv_chamfer_table(int R, float *table);

/* ---------------------------------------------------------------------------------
 * K4  ProPainter-style flow-guided propagation prior (image propagation, 'nearest').
 *   Replaces the un-vendored propainter.forward call at diffuerase.py:49-57
 *   [BidirectionalPropagation(learnable=False) + fbConsistencyCheck + flow_warp].
 *   The clip is given once; `n_sub` sub-video windows [sub_start[s], sub_start[s]+sub_len[s])
 *   (HOST int arrays; propainter/inference.py uses 50 frames + 10 pad frames each side) are
 *   independent scans and are advanced in lock step.
 *   frames:   u8  [N,h,w,3]        masks: u8 [N,h,w] (>0 = hole)
 *   flows_f:  f32 [N-1,h,w,2]      flow t -> t+1 (x,y in pixels);  flows_b: t+1 -> t
 *   sub_keep_start / sub_keep_len (HOST int arrays, both NULL = keep every frame of every window):
 *             frames [keep_start, keep_start+keep_len) of window s are kept, the others are the pad
 *             frames upstream computes and throws away; their state only lives in the workspace.
 *   out:      u32 [sum(keep_len),h,w], the kept frames of the windows concatenated in the given order
 *             (for the upstream window plan that is exactly [N,h,w]); each word is
 *             R | G<<8 | B<<16 | state<<24 of the forward pass; state bit0 = still a hole,
 *             bit1 = value is the float 0.0 of the masked frame (hole or zero-padded warp)
 *             rather than a u8 level.
 *   workspace: vv_propagate_workspace_bytes(sum(sub_len), sum(sub_len) - sum(keep_len), h, w) bytes.
 * --------------------------------------------------------------------------------- */
VV_API size_t vv_propagate_workspace_bytes(int n_window_frames, int n_pad_frames, int h, int w);
VV_API int vv_propagate(const uint8_t *frames, const uint8_t *masks, const float *flows_f,
                 const float *flows_b, int n_frames, int h, int w, const int *sub_start,
                 const int *sub_len, const int *sub_keep_start, const int *sub_keep_len, int n_sub,
                 uint32_t *out, void *workspace, size_t workspace_bytes, void *stream);
/* Unpack K4's output: rgb u8 [N,3] (zero-valued pixels get `zero_level`), hole mask u8 [N]
 * in {0,255}.  Either output may be NULL. */
VV_API int vv_propagate_unpack(const uint32_t *packed, size_t n_pixels, uint8_t zero_level,
                        uint8_t *rgb, uint8_t *hole_mask, void *stream);

/* ---------------------------------------------------------------------------------
 * K5  chunk-overlap feather blend (builder-defined spec; the reference
 *   only lists it as a TODO, README.md:76).  For overlap frame k in [k0, k0+O):
 *   w = f32(k+1)/f32(O_total+1); out = u8(rint(f32((1-w)*A) + f32(w*B))).
 *   A = earlier chunk's tail, B = later chunk's head, both u8 [O,H,W,C]; B may be a
 *   peer-GPU pointer (P2P / IPC mapped) - the kernel reads it in place over NVLink.
 * --------------------------------------------------------------------------------- */
VV_API int vv_chunk_blend(const uint8_t *A, const uint8_t *B, int O, size_t frame_bytes, int k0,
                   int O_total, uint8_t *out, void *stream);

/* Rank-boundary halo blend (one process per GPU, consecutive ranks share `overlap` frames), with a
 * device-side handshake instead of host barriers.  `out`: this rank's u8 frames [T, frame_bytes], blended in
 * place: overlap indices [0, overlap/2) of the boundary with the next rank (frames T-overlap+k) against
 * `next_head` = the next rank's frame 0 (peer memory, CUDA-IPC mapped), and [overlap/2, overlap) of the boundary
 * with the previous rank (frames k) against `prev_tail` = the previous rank's frame T_prev-overlap+overlap/2.
 * Either neighbour may be absent (NULL frames and NULL flags).  `*_flags`: 8 x u32 per rank in IPC-mapped,
 * zero-initialised device memory: [0] ready epoch, [1] consumed by the previous rank, [2] consumed by the next
 * rank, [3] block counter, [4] error (a peer did not show up within ~2 s).  `epoch` must increase by one on
 * every call, identically on all ranks.  The kernel polls the peers' ready flags before reading their frames and
 * only completes once the neighbours have read this rank's frames, so the stream may overwrite `out` next. */
VV_API int vv_halo_blend(uint8_t *out, int T, size_t frame_bytes, int overlap, const uint8_t *next_head,
                         const uint8_t *prev_tail, uint32_t *my_flags, uint32_t *next_flags, uint32_t *prev_flags,
                         uint32_t epoch, void *stream);

/* ---------------------------------------------------------------------------------
 * "Next" rows: the pixel glue immediately either side of the hot path.
 * N3  SAM2 mask colour painter.                        Replaces sam2_masker.py:151-175
 *   masks: u8 (non-zero = set) or f32 logits (> 0 = set) [T,K,mh,mw], object k in paint order
 *   (ascending object id: the LAST one wins where masks overlap, :159-173); NEAREST-resized to
 *   the canvas when (mh,mw) != (H0,W0) (:167).  colors: HOST u8 [K,3], written as given (the
 *   reference stores its (B,G,R) tuples).  out: u8 [T,H0,W0,3], black where no object is set.
 * N2  K4 packed state -> f32 [n_frames,3,h,w] in [-1,1] (`to_tensors()*2-1`, 0.0 for state bit 1)
 *   and the f32 hole mask [n_frames,h,w] (may be NULL): what the ProPainter network consumes.
 * --------------------------------------------------------------------------------- */
VV_API size_t vv_paint_masks_workspace_bytes(int H0, int W0);
VV_API int vv_paint_masks(const void *masks, int mask_is_f32, int T, int K, int mh, int mw,
                          const uint8_t *colors_host, uint8_t *out, int H0, int W0, void *workspace,
                          size_t workspace_bytes, void *stream);
VV_API int vv_propagate_to_float(const uint32_t *packed, int n_frames, int h, int w, float *rgb_chw,
                                 float *hole_mask, void *stream);
/* N4  DiffuEraser wrapper pixel steps either side of the diffusion call made at diffuerase.py:62-67
 *   [recalled-upstream diffueraser/diffueraser.py; parity unpinned by the reference, the OpenCV
 *   primitives are pinned against cv2: oracle/wrapper.py].
 *   vv_wrapper_mask: read_mask = (mask > 0) -> cv2.erode(3x3 rect, 1) -> cv2.dilate(3x3 rect,
 *     dilation_iter) -> {0,255}; mask, out: u8 [T,h,w] (the reference passes dilation_iter 0).
 *   vv_wrapper_compose: blended != 0: alpha = u8((1 - (1 - m/255.)(1 - cv2.GaussianBlur(m,(21,21),0)/255.))
 *     * 255) for the {0, non-zero} mask m, else alpha = m;  out = u8(img * a + frame * (1 - a)) with
 *     a = f32(alpha)/255 in fp32, truncated.  img, frames, out: u8 [T,h,w,3]; mask255: u8 [T,h,w].
 *     Blending needs h, w >= 11 (single REFLECT_101 bounce) and w <= ~4000 (shared-memory strip). */
VV_API int vv_wrapper_mask(const uint8_t *mask, int T, int h, int w, int dilation_iter, uint8_t *out, void *stream);
VV_API int vv_wrapper_compose(const uint8_t *img, const uint8_t *frames, const uint8_t *mask255, int T, int h,
                              int w, int blended, uint8_t *out, void *stream);

/* N2  neighbour-window merge of the ProPainter network's output (call site diffuerase.py:52-57;
 *   [recalled-upstream propainter/inference.py], oracle/propagation.py ref_neighbor_merge):
 *     img  = mask ? u8(((pred + 1) / 2) * 255) : ori          (float32, truncation)
 *     comp = first ? img : (comp + img) >> 1                   (== u8(f32(comp)*0.5 + f32(img)*0.5))
 *   pred_chw: f32 [L,3,h,w] in [-1,1]; mask: u8 [L,h,w] (> 0 = masked); ori, comp: u8 [L,h,w,3]; comp is
 *   updated in place.  Bit l of first_mask = frame l of the window has not been composed before.  L <= 64. */
VV_API int vv_neighbor_merge(const float *pred_chw, const uint8_t *mask, const uint8_t *ori, uint8_t *comp, int L,
                             int h, int w, unsigned long long first_mask, void *stream);
/* N4  masked frames of the DiffuEraser wrapper: out = mask > 0 ? 0 : frame  (frame * (1 - m)).
 *   frames, out: u8 [T,h,w,3]; mask: u8 [T,h,w]. */
VV_API int vv_apply_mask(const uint8_t *frames, const uint8_t *mask, int T, int h, int w, uint8_t *out, void *stream);
/* N1  channel swap BGR <-> RGB of packed 3-byte pixels (tools.py:21 cv2.cvtColor(..., COLOR_BGR2RGB), and
 *   the swap back before VideoWriter.write at tools.py:43); src == dst is allowed. */
VV_API int vv_swap_rb(const uint8_t *src, uint8_t *dst, size_t n_pixels, void *stream);

/* ---------------------------------------------------------------------------------
 * Host-buffer pipeline (the path the Python drop-in takes for lists of numpy frames, i.e. the
 * whole of diffuerase.py:26-31 and :69-114 with per-frame HOST pointers in and out).
 * A context owns device buffers, streams and (lazily) pinned staging rings for one original
 * geometry H0 x W0; batches of `frames_per_batch` frames rotate over `n_slots` streams so that
 * H2D, kernels and D2H of neighbouring batches overlap.  Page-locked host buffers are copied
 * directly, pageable ones through the staging ring.  Calls block until the results are in the
 * caller's host buffers.  One job at a time per context (internally locked).
 * --------------------------------------------------------------------------------- */
typedef struct vv_pipeline vv_pipeline;
VV_API int vv_pipeline_create(vv_pipeline **p, int device, int H0, int W0, int frames_per_batch,
                              int n_slots);
VV_API void vv_pipeline_destroy(vv_pipeline *p);
/* pre (diffuerase.py:28-31): T host pointers to u8 [H0,W0,C] masks -> host u8 [H0,W0] dilated
 * masks, and optionally (lowres_out != NULL) the NEAREST [lh,lw] masks.  Device copies of the
 * dilated masks stay resident in the context for the matching vv_pipeline_post call. */
VV_API int vv_pipeline_pre(vv_pipeline *p, const uint8_t *const *masks, int T, int C, int iterations,
                           uint8_t *const *dilated_out, uint8_t *const *lowres_out, int lh, int lw);
/* down-size frames for the model (row A9): host [H0,W0,3] -> host [h,w,3]. */
VV_API int vv_pipeline_downsize(vv_pipeline *p, const uint8_t *const *frames, int T, int h, int w,
                                uint8_t *const *small_out);
/* post (diffuerase.py:70-112): inpainted [h,w,3] + originals [H0,W0,3] -> out [H0,W0,3].
 * `dilated` may be NULL to reuse the masks kept by the preceding vv_pipeline_pre. */
VV_API int vv_pipeline_post(vv_pipeline *p, const uint8_t *const *inpainted, int h, int w,
                            const uint8_t *const *orig, const uint8_t *const *dilated, int T,
                            float feather_px, int keep_unmasked, uint8_t *const *out);

/* Frame rows per direction that the last vv_pipeline_post moved over PCIe, and frames x H0 (equal unless the call was
 * row bounded: page-locked orig / out buffers, resident masks, at most 75 % of the rows inside the masks' row ranges
 * + feather radius; the other rows of `out` are then host copies of `orig`, diffuerase.py:108-112 with alpha = 0).
 * Option "pipe_rows" = 0 switches the row-bounded mode off. */
VV_API int vv_pipeline_last_rows(vv_pipeline *p, long long *rows_moved, long long *rows_total);

/* Device-resident clips (the adapter around the two networks, videovanish_b200/wrappers.py): T per-frame host
 * buffers of `frame_bytes` each -> one contiguous device array, and back.  `stream` is the stream that consumes
 * (upload) / produced (download) the device data: the upload makes it wait on the copies, the download waits
 * for the work enqueued on it so far.  The upload is asynchronous for page-locked sources (they must stay
 * alive and unmodified until `stream` has passed the wait; pageable ones are staged before it returns); the
 * download returns when the data is in the caller's host buffers. */
VV_API int vv_pipeline_upload(vv_pipeline *p, const uint8_t *const *src, int T, size_t frame_bytes, uint8_t *dev_dst,
                              void *stream);
VV_API int vv_pipeline_download(vv_pipeline *p, const uint8_t *dev_src, int T, size_t frame_bytes,
                                uint8_t *const *dst, void *stream);

/* Row-bounded results.  The composite (diffuerase.py:70-112) only changes pixels within the feather radius of a mask
 * pixel, so outside the row range [lo, hi) of a frame the finished frame equals the input frame (`orig`, :108).
 * vv_mask_row_bounds: bounds[2t] = lo, bounds[2t+1] = hi (exclusive) of frame t from K1's dilated 1-bit plane
 *   (i32 device array of 2T; `margin` rows added either side: the feather radius; (0, 0) for an empty mask).
 * vv_pipeline_host_rows_begin: starts copying the rows OUTSIDE [lo, hi) from the caller's input frames `src` to the
 *   result frames `dst` on a background memcpy pool (host pointers; arrays of T); returns at once.
 * vv_pipeline_download_rows: the rows INSIDE [lo, hi) of the device frames -> `dst` (page-locked), waits for the
 *   background copy; like vv_pipeline_download it first waits for the work enqueued on `stream`. */
VV_API int vv_mask_row_bounds(const uint32_t *mask_bits, int T, int H, int Wp, int margin, int *bounds, void *stream);
VV_API int vv_pipeline_host_rows_begin(vv_pipeline *p, int T, int H, size_t row_bytes, uint8_t *const *dst,
                                       const uint8_t *const *src, const int *lo, const int *hi);
VV_API int vv_pipeline_download_rows(vv_pipeline *p, const uint8_t *dev_src, int T, int H, size_t row_bytes,
                                     uint8_t *const *dst, const int *lo, const int *hi, void *stream);

/* Peer-memory helpers for the multi-GPU halo blend (one process per GPU). */
/* handle of the allocation that contains dev_ptr + the offset of dev_ptr inside it */
VV_API int vv_ipc_get_handle(const void *dev_ptr, void *handle_out_64B, size_t *offset_out);
VV_API int vv_ipc_open_handle(const void *handle_64B, void **mapped_ptr);
VV_API int vv_ipc_close_handle(void *mapped_ptr);

#ifdef __cplusplus
}
#endif
#endif /* VVB200_H_ */

"""Device-resident adapters for the pixel stages either side of the two networks.

The reference hands Python lists of host frames to two un-vendored model wrappers
(/root/reference/diffuerase.py:49-57 ``propainter.forward`` and :62-67 ``video_inpainting_sd.forward``).
Inside those wrappers, around the networks proper, sit more pixel stages of the same class as the hot
path.  The classes below restate those sequences
[recalled-upstream ``propainter/inference.py`` / ``diffueraser/diffueraser.py``; PARITY UNPINNED, see
oracle/propagation.py and oracle/wrapper.py] on the kernels of ``libvvb200.so`` with the NETWORKS left as
pluggable callables, so that a whole ``run_infill_on_frames`` call can stay on the device between the
first upload and the last download:

    ProPainterPrior   K2 down-size + NEAREST mask -> [flow_fn: RAFT + flow completion] -> K4 propagation ->
                      N2 float hand-off -> per neighbour window [network_fn] -> N2 0.5/0.5 u8 merge
    DiffuEraserWrapper  K2 down-size + NEAREST mask -> N4 read_mask (erode / dilate) + masked frames ->
                      [network_fn: diffusion] -> N4 (blurred) compose

Both keep the upstream ``forward(...)`` signatures the reference calls (host lists in, host lists out) and
add ``forward_device(clip, ...)``, which ``diffuerase.run_infill_on_frames`` uses when every model in play
offers it.  There is no CPU fallback.
"""
import torch

from . import hostpipe, ops


def neighbor_plan(video_length, neighbor_length=10, ref_stride=10, subvideo_length=50):
    """Sliding windows of the upstream feature-propagation loop: [(neighbor_ids, ref_ids)] for
    ``f in range(0, video_length, neighbor_length // 2)`` (diffuerase.py:54 passes 10 / 10 / 50)."""
    stride = neighbor_length // 2
    ref_num = subvideo_length // ref_stride if video_length > subvideo_length else -1
    plan = []
    for f in range(0, video_length, stride):
        ids = list(range(max(0, f - stride), min(video_length, f + stride + 1)))
        refs = []
        if ref_num == -1:
            refs = [i for i in range(0, video_length, ref_stride) if i not in ids]
        else:
            lo = max(0, f - ref_stride * (ref_num // 2))
            hi = min(video_length, f + ref_stride * (ref_num // 2))
            for i in range(lo, hi, ref_stride):
                if i not in ids:
                    if len(refs) > ref_num:
                        break
                    refs.append(i)
        plan.append((ids, refs))
    return plan


class DeviceClip:
    """One clip resident in HBM: original frames, dilated masks and the inference-resolution copies the
    stages share (each computed once)."""

    def __init__(self, frames, dilated, lowres=None):
        self.frames = frames                  # u8 [T,H0,W0,3]
        self.dilated = dilated                # u8 [T,H0,W0] in {0,255}
        self.mask_bits = None                 # optional 1-bit plane of `dilated` left by K1 (consumed by K3)
        self._small, self._low = {}, {}
        if lowres is not None:
            self._low[tuple(lowres.shape[1:3])] = lowres

    @property
    def size(self):
        return tuple(self.frames.shape[1:3])

    def small(self, h, w):
        """Frames at (h, w): row A9, cv2 INTER_LINEAR semantics (K2)."""
        if (h, w) == self.size:
            return self.frames
        if (h, w) not in self._small:
            self._small[(h, w)] = ops.resize(self.frames, h, w)
        return self._small[(h, w)]

    def low_mask(self, h, w):
        """Dilated masks at (h, w): INTER_NEAREST (fused into K1 when the size was known there)."""
        if (h, w) == self.size:
            return self.dilated
        if (h, w) not in self._low:
            self._low[(h, w)] = ops.resize(self.dilated, h, w, ops.INTER_NEAREST)
        return self._low[(h, w)]


def identity_flow_fn(small, low):
    raise RuntimeError("ProPainterPrior needs a flow_fn(frames_u8[T,h,w,3], masks_u8[T,h,w]) -> (flows_f, flows_b)")


class ProPainterPrior:
    """Pixel stages of the upstream ``Propainter.forward`` (call site diffuerase.py:52-57).

    ``flow_fn(frames, masks) -> (flows_f, flows_b)``: f32 [T-1,h,w,2] device tensors (RAFT + the flow-completion
    network upstream).  ``network_fn(updated f32 [T,3,h,w], updated_masks f32 [T,h,w], masks u8 [T,h,w],
    neighbor_ids, ref_ids) -> f32 [len(neighbor_ids),3,h,w]`` in [-1,1] (the ProPainter transformer); ``None``
    returns the propagated pixels themselves (holes left at mid-grey)."""

    def __init__(self, flow_fn=identity_flow_fn, network_fn=None, max_img_size=960):
        self.flow_fn, self.network_fn, self.max_img_size = flow_fn, network_fn, max_img_size
        self._pipe = None

    def process_size(self, h0, w0):
        return ops.inference_size(h0, w0, self.max_img_size)

    def forward_device(self, clip, ref_stride=10, neighbor_length=10, subvideo_length=50, mask_dilation=0,
                       progress=None):
        """-> u8 [T,h,w,3] prior frames at the processing size, on the device."""
        h, w = self.process_size(*clip.size)
        small, low = clip.small(h, w), clip.low_mask(h, w)
        if mask_dilation > 0:                                   # upstream dilates only when asked to (:55 passes 0)
            low = ops.binarize_dilate(low, mask_dilation)
        t = small.shape[0]
        flows_f, flows_b = self.flow_fn(small, low) if t > 1 else (None, None)
        packed = ops.propagate(small, low, flows_f, flows_b, subvideo_length=subvideo_length)     # K4
        if self.network_fn is None:
            return ops.propagate_unpack(packed, zero_level=127)[0]
        updated, updated_masks = ops.propagate_to_float(packed)                                  # N2 hand-off
        comp = torch.empty_like(small)
        seen = [False] * t
        for ids, refs in neighbor_plan(t, neighbor_length, ref_stride, subvideo_length):
            a, b = ids[0], ids[-1] + 1
            pred = self.network_fn(updated, updated_masks, low, ids, refs)
            ops.neighbor_merge(pred, low[a:b], small[a:b], comp[a:b], [not seen[i] for i in ids])  # N2 merge
            for i in ids:
                seen[i] = True
        return comp

    def forward(self, frames, masks, ref_stride=10, neighbor_length=10, subvideo_length=50, mask_dilation=0,
                progress=None):
        """Upstream signature: lists of host frames / single-channel masks in, list of H0xW0x3 host frames out
        (the prior is brought back to the original size with INTER_LINEAR, like upstream's cv2.resize)."""
        h0, w0 = frames[0].shape[:2]
        pipe = self._get_pipe(h0, w0)
        clip = DeviceClip(pipe.upload(frames, (h0, w0, 3)), pipe.upload(masks, (h0, w0), is_mask=True))
        comp = self.forward_device(clip, ref_stride, neighbor_length, subvideo_length, mask_dilation, progress)
        if tuple(comp.shape[1:3]) != (h0, w0):
            comp = ops.resize(comp, h0, w0)
        return pipe.download(comp)

    def _get_pipe(self, h0, w0):
        if self._pipe is None or self._pipe.geometry != (h0, w0):
            self._pipe = hostpipe.HostPipeline(h0, w0)
        return self._pipe


class DiffuEraserWrapper:
    """Pixel stages of the upstream ``DiffuEraser.forward`` (call site diffuerase.py:62-67).

    ``network_fn(masked_frames u8 [T,h,w,3], masks u8 [T,h,w], priors u8 [T,h,w,3]) -> u8 [T,h,w,3]``: the
    diffusion pipeline proper."""

    def __init__(self, network_fn, blended=True):
        self.network_fn, self.blended = network_fn, blended
        self._pipe = None

    def forward_device(self, clip, priors, max_img_size=960, mask_dilation_iter=0, guidance_scale=None, progress=None):
        """-> u8 [T,h,w,3] composed frames at inference resolution, on the device."""
        h, w = ops.inference_size(*clip.size, max_img_size)
        small, low = clip.small(h, w), clip.low_mask(h, w)
        m = ops.wrapper_mask(low, mask_dilation_iter)                    # N4 read_mask: erode 3x3, dilate 3x3 x iter
        masked = ops.apply_mask(small, m)                                # N4 frame * (1 - m)
        if tuple(priors.shape[1:3]) != (h, w):
            priors = ops.resize(priors, h, w)                            # read_priori's resize
        images = self.network_fn(masked, m, priors)
        return ops.wrapper_compose(images, small, m, blended=self.blended)   # N4 compose

    def forward(self, frames, masks, priors, max_img_size=960, mask_dilation_iter=0, guidance_scale=None,
                progress=None):
        """Upstream signature: host lists in, list of hxwx3 host frames out."""
        h0, w0 = frames[0].shape[:2]
        if self._pipe is None or self._pipe.geometry != (h0, w0):
            self._pipe = hostpipe.HostPipeline(h0, w0)
        pipe = self._pipe
        clip = DeviceClip(pipe.upload(frames, (h0, w0, 3)), pipe.upload(masks, (h0, w0), is_mask=True))
        pri = pipe.upload(priors, tuple(priors[0].shape))
        return pipe.download(self.forward_device(clip, pri, max_img_size, mask_dilation_iter, guidance_scale, progress))


class LazyHostFrames:
    """The dilated masks of a device-resident call, as the list the reference hands to its models
    (diffuerase.py:53, :63): downloaded on first use, never if nobody looks."""

    def __init__(self, pipe, tensor):
        self._pipe, self._tensor, self._host = pipe, tensor, None

    def _materialise(self):
        if self._host is None:
            self._host = self._pipe.download(self._tensor)
        return self._host

    def __len__(self):
        return self._tensor.shape[0]

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())

"""Frame I/O with the call surface of the reference's ``tools`` module
(/root/reference/tools.py:4 ``load_video_frames_from_path``, :30 ``write_video_frames_to_path``):
same arguments, same return values, same container / codec, same asserts.

Decoding and encoding stay with OpenCV (there is no codec library in this image; NVDEC / NVENC would slot in
here).  What this module changes (SURVEY "next" row N1):

* host route (default, the reference's contract): decoded frames land in page-locked blocks, ``_BLOCK_FRAMES`` at a
  time, so ``run_infill_on_frames`` can DMA them to the GPU without a staging copy - up to ``PINNED_BUDGET`` bytes
  per call; beyond it the blocks are ordinary host memory (page-locking tens of GB starves the host);
* device route (``device="cuda"``): the clip goes straight to HBM.  A block of BGR frames is decoded into a
  page-locked ring slot, uploaded asynchronously and channel-swapped ON THE DEVICE (``vv_swap_rb``) while OpenCV
  decodes the next block; the result is a ``DeviceFrames`` sequence that ``diffuerase.run_infill_on_frames`` takes
  in place of the list of host arrays (no second upload).  The writer accepts it too: swap on the device, download
  of block b+1 under the encode of block b.
"""
import collections.abc

import cv2
import numpy as np

_BLOCK_FRAMES = 32
PINNED_BUDGET = 8 << 30          # page-locked bytes handed out per load call (same bound as hostpipe.PINNED_RESULT_LIMIT)


class DeviceFrames(collections.abc.Sequence):
    """A clip resident on the GPU, u8 [T,H,W,3] RGB, that still behaves like the reference's list of HxWx3 uint8
    arrays: indexing / iterating downloads the whole clip once (page-locked when the budget allows)."""

    def __init__(self, tensor):
        self.tensor = tensor
        self._host = None

    def _materialise(self):
        if self._host is None:
            from . import hostpipe
            t = self.tensor.shape[0]
            self._host = hostpipe.pinned_frames(t, tuple(self.tensor.shape[1:]))
            import torch
            whole = torch.from_numpy(self._host[0].base)      # the block the per-frame views share
            whole.copy_(self.tensor)                          # one D2H copy
            torch.cuda.current_stream().synchronize()
        return self._host

    def __len__(self):
        return int(self.tensor.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return DeviceFrames(self.tensor[i])
        return self._materialise()[i]

    @property
    def shape(self):
        return tuple(self.tensor.shape)


class _FramePool:
    """Hands out HxWx3 uint8 slots carved from page-locked blocks (ordinary memory beyond the budget)."""

    def __init__(self, budget=None):
        self._block = None
        self._used = 0
        self._left = PINNED_BUDGET if budget is None else budget

    def _allocate(self, shape):
        nbytes = _BLOCK_FRAMES * int(np.prod(shape))
        if nbytes <= self._left:
            try:
                import torch
                if torch.cuda.is_available():
                    block = torch.empty((_BLOCK_FRAMES,) + tuple(shape), dtype=torch.uint8, pin_memory=True).numpy()
                    self._left -= nbytes
                    return block
            except Exception:
                pass
        return np.empty((_BLOCK_FRAMES,) + tuple(shape), np.uint8)

    def slot(self, shape):
        exhausted = self._block is None or self._used == _BLOCK_FRAMES or self._block.shape[1:] != tuple(shape)
        if exhausted:
            self._block, self._used = self._allocate(shape), 0
        view = self._block[self._used]
        self._used += 1
        return view


def _decoded_frames(capture):
    """Yield (index, BGR frame) until the stream ends."""
    index = 0
    while True:
        ok, bgr = capture.read()
        if not ok:
            return
        yield index, bgr
        index += 1


def _selected(capture, start_frame, max_frames):
    n = 0
    for index, bgr in _decoded_frames(capture):
        if index < start_frame:
            continue
        yield bgr
        n += 1
        if 0 < max_frames <= n:
            return


def _load_to_device(capture, start_frame, max_frames):
    """Decode -> pinned ring slot (BGR) -> async H2D -> swap to RGB on the device; decoding block b+1 overlaps the
    upload of block b (two ring slots, one copy stream)."""
    import torch

    from . import ops
    if not torch.cuda.is_available():
        raise RuntimeError("load_video_frames_from_path(device='cuda') needs a CUDA device")
    copy_stream = torch.cuda.Stream()
    ring, ring_free, blocks = None, None, []
    fill, slot = 0, 0

    def flush(n):
        nonlocal slot
        with torch.cuda.stream(copy_stream):
            dev = torch.empty((n,) + tuple(ring[slot].shape[1:]), dtype=torch.uint8, device="cuda")
            dev.copy_(ring[slot][:n], non_blocking=True)
            ops.swap_rb(dev, out=dev)                                  # tools.py:21, on the device, in place
            ring_free[slot].record(copy_stream)
        blocks.append(dev)
        slot ^= 1
        ring_free[slot].synchronize()                                  # the slot decoded into next is no longer being read

    for bgr in _selected(capture, start_frame, max_frames):
        if ring is None:
            ring = [torch.empty((_BLOCK_FRAMES,) + bgr.shape, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            ring_free = [torch.cuda.Event(), torch.cuda.Event()]
        ring[slot][fill].numpy()[...] = bgr
        fill += 1
        if fill == _BLOCK_FRAMES:
            flush(fill)
            fill = 0
    if fill:
        flush(fill)
    assert len(blocks) > 0, "No frames read"
    copy_stream.synchronize()
    clip = blocks[0] if len(blocks) == 1 else torch.cat(blocks)
    return DeviceFrames(clip)


def load_video_frames_from_path(video_path, start_frame=0, max_frames=-1, device=None):
    """Decode ``video_path`` and return ``(frames, fps)``: RGB uint8 HxWx3 arrays from
    ``start_frame`` on, at most ``max_frames`` of them when that is positive (tools.py:4-28).
    ``device="cuda"`` returns a ``DeviceFrames`` clip instead of the list of host arrays."""
    capture = cv2.VideoCapture(video_path)
    assert capture.isOpened(), f"Failed to open video: {video_path}"
    fps = capture.get(cv2.CAP_PROP_FPS)
    try:
        if device is not None:
            return _load_to_device(capture, start_frame, max_frames), fps
        pool, frames = _FramePool(), []
        for bgr in _selected(capture, start_frame, max_frames):
            rgb = pool.slot(bgr.shape)
            cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=rgb)          # channel swap of tools.py:21, in place
            frames.append(rgb)
    finally:
        capture.release()
    assert len(frames) > 0, "No frames read"
    return frames, fps


def _fit_nearest(rgb, H0, W0):
    """A frame of another size is brought to (H0, W0) with NEAREST, like the reference's writer (:41-42) - on
    the GPU (K2 nearest kernel, bit-exact against cv2.INTER_NEAREST); like every pixel stage of this package it
    has no CPU fallback."""
    import torch

    from . import ops
    if not torch.cuda.is_available():
        raise RuntimeError("write_video_frames_to_path: resizing a frame needs a CUDA device (no CPU fallback)")
    src = torch.from_numpy(np.ascontiguousarray(rgb)[None]).cuda()
    return ops.resize(src, H0, W0, ops.INTER_NEAREST)[0].cpu().numpy()


def _device_bgr_blocks(clip, H0, W0):
    """Yield host BGR blocks of a DeviceFrames clip: NEAREST fix-up and RGB -> BGR on the device, block b+1 is
    downloaded (copy stream, page-locked ring) while the caller encodes block b."""
    import torch

    from . import ops
    t = clip.tensor
    if tuple(t.shape[1:3]) != (H0, W0):
        t = ops.resize(t, H0, W0, ops.INTER_NEAREST)
    copy_stream = torch.cuda.Stream()
    copy_stream.wait_stream(torch.cuda.current_stream())
    ring = [torch.empty((_BLOCK_FRAMES, H0, W0, 3), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def start(b, slot):
        lo = b * _BLOCK_FRAMES
        n = min(_BLOCK_FRAMES, t.shape[0] - lo)
        with torch.cuda.stream(copy_stream):
            bgr = ops.swap_rb(t[lo:lo + n])                             # tools.py:43, on the device
            ring[slot][:n].copy_(bgr, non_blocking=True)
            bgr.record_stream(copy_stream)
            done[slot].record(copy_stream)
        return n

    n_blocks = (t.shape[0] + _BLOCK_FRAMES - 1) // _BLOCK_FRAMES
    pending = start(0, 0)
    for b in range(n_blocks):
        slot = b & 1
        n = pending
        if b + 1 < n_blocks:
            pending = start(b + 1, slot ^ 1)
        done[slot].synchronize()
        yield ring[slot][:n].numpy()


def write_video_frames_to_path(out_video, mask_frames, fps, H0, W0):
    """Encode RGB frames as lossless FFV1 (tools.py:30-45); frames of another size are brought to
    (W0, H0) with NEAREST first, exactly like the reference's writer (:41-42).  ``mask_frames`` may be a
    ``DeviceFrames`` clip: the fix-up and the channel swap then run on the device."""
    sink = cv2.VideoWriter(out_video, cv2.VideoWriter_fourcc(*"FFV1"), fps, (W0, H0))
    assert sink.isOpened(), "Failed to open VideoWriter (FFV1/MKV). Try MJPG or mp4v if needed."
    count = 0
    if isinstance(mask_frames, DeviceFrames):
        for block in _device_bgr_blocks(mask_frames, H0, W0):
            for bgr in block:
                sink.write(bgr)
                count += 1
    else:
        for rgb in mask_frames:
            if rgb.shape[:2] != (H0, W0):
                rgb = _fit_nearest(rgb, H0, W0)         # NEAREST commutes with the channel swap below
            bgr = cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)
            sink.write(bgr)
            count += 1
    sink.release()
    print(f"[ok] wrote {count} frames to {out_video}")

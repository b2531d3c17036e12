"""Frame I/O with the call surface of the reference's ``tools`` module
(/root/reference/tools.py:4 ``load_video_frames_from_path``, :30 ``write_video_frames_to_path``):
same arguments, same return values, same container / codec, same asserts.

Decoding and encoding stay with OpenCV (device-side codecs are SURVEY "next" row N1).  What this
module changes is where decoded frames land: RGB frames are written straight into page-locked
blocks, ``_BLOCK_FRAMES`` at a time, so ``run_infill_on_frames`` can DMA them to the GPU without a
staging copy.  Without a CUDA device the blocks are ordinary host memory (this is I/O, not compute).
"""
import cv2
import numpy as np

_BLOCK_FRAMES = 32


class _FramePool:
    """Hands out HxWx3 uint8 slots carved from page-locked blocks."""

    def __init__(self):
        self._block = None
        self._used = 0

    @staticmethod
    def _allocate(shape):
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty((_BLOCK_FRAMES,) + tuple(shape), dtype=torch.uint8, pin_memory=True).numpy()
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


def load_video_frames_from_path(video_path, start_frame=0, max_frames=-1):
    """Decode ``video_path`` and return ``(frames, fps)``: RGB uint8 HxWx3 arrays from
    ``start_frame`` on, at most ``max_frames`` of them when that is positive (tools.py:4-28)."""
    capture = cv2.VideoCapture(video_path)
    assert capture.isOpened(), f"Failed to open video: {video_path}"
    fps = capture.get(cv2.CAP_PROP_FPS)
    pool, frames = _FramePool(), []
    try:
        for index, bgr in _decoded_frames(capture):
            if index < start_frame:
                continue
            rgb = pool.slot(bgr.shape)
            cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=rgb)          # channel swap of tools.py:21, in place
            frames.append(rgb)
            if 0 < max_frames <= len(frames):
                break
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


def write_video_frames_to_path(out_video, mask_frames, fps, H0, W0):
    """Encode RGB frames as lossless FFV1 (tools.py:30-45); frames of another size are brought to
    (W0, H0) with NEAREST first, exactly like the reference's writer (:41-42)."""
    sink = cv2.VideoWriter(out_video, cv2.VideoWriter_fourcc(*"FFV1"), fps, (W0, H0))
    assert sink.isOpened(), "Failed to open VideoWriter (FFV1/MKV). Try MJPG or mp4v if needed."
    count = 0
    for rgb in mask_frames:
        if rgb.shape[:2] != (H0, W0):
            rgb = _fit_nearest(rgb, H0, W0)         # NEAREST commutes with the channel swap below
        bgr = cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)
        sink.write(bgr)
        count += 1
    sink.release()
    print(f"[ok] wrote {count} frames to {out_video}")

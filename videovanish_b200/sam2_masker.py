"""Drop-in pieces of the reference's ``sam2_masker`` module that touch pixels
(/root/reference/sam2_masker.py:27-37, :151-175): the per-object colour table and the colour painter
that turns SAM2's per-object masks into the mask video the hot path consumes.  The SAM2 network itself
is out of scope; ``paint_mask_frames`` takes what ``predictor.propagate_in_video`` produced.
"""
import cv2
import numpy as np
import torch

from . import ops


def color_for_obj(obj_id):
    """Deterministic bright colour per object id, as the reference computes it (:27-37)."""
    h = int((obj_id * 37) % 180)
    hsv = np.uint8([[[h, 200, 255]]])
    bgr = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)[0, 0]
    return tuple(int(x) for x in bgr)


def paint_mask_frames(video_segments, n_frames, H0, W0, device=None):
    """Same result as the loop at sam2_masker.py:151-175: ``video_segments`` is
    ``{frame_idx: {obj_id: mask}}`` with boolean / uint8 numpy masks, float logits, or CUDA tensors of a
    common shape; returns a list of ``n_frames`` uint8 HxWx3 arrays (black background, higher object
    ids painted over lower ones, masks NEAREST-resized to (H0, W0) when smaller)."""
    if not torch.cuda.is_available():
        raise RuntimeError("videovanish_b200: CUDA device required (there is no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ids = sorted({int(o) for seg in video_segments.values() for o in seg})
    if not ids:
        return [np.zeros((H0, W0, 3), np.uint8) for _ in range(n_frames)]
    first = next(m for seg in video_segments.values() for m in seg.values() if m is not None)
    shape = tuple(np.squeeze(np.asarray(first.cpu() if isinstance(first, torch.Tensor) else first)).shape)
    is_float = (first.dtype.is_floating_point if isinstance(first, torch.Tensor) else np.asarray(first).dtype.kind == "f")
    stack = torch.zeros((n_frames, len(ids)) + shape, dtype=torch.float32 if is_float else torch.uint8, device=dev)
    for idx, seg in video_segments.items():
        if not (0 <= idx < n_frames):
            continue
        for obj, m in seg.items():
            if m is None:
                continue
            t = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.squeeze(np.asarray(m))))
            t = t.reshape(shape).to(dev)
            stack[idx, ids.index(int(obj))] = t.to(stack.dtype) if is_float else (t != 0).to(torch.uint8)
    out = ops.paint_masks(stack, [color_for_obj(o) for o in ids], out_size=(H0, W0))
    host = out.cpu().numpy()
    return [host[i] for i in range(n_frames)]

"""Drop-in for the reference's ``tools`` module (/root/reference/tools.py:4-45): same two
functions, same arguments and return values.  Decoding / encoding stays with OpenCV (frame I/O is
SURVEY "next" row N1); what changes is where the frames land: decoded RGB frames are written
straight into page-locked blocks, so ``run_infill_on_frames`` can DMA them without a staging copy.
"""
import cv2
import numpy as np

_BLOCK_FRAMES = 32


def _new_block(shape):
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty((_BLOCK_FRAMES,) + tuple(shape), dtype=torch.uint8, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty((_BLOCK_FRAMES,) + tuple(shape), np.uint8)      # plain host memory: I/O only, no compute


def load_video_frames_from_path(video_path, start_frame=0, max_frames=-1):
    """tools.py:4-28.  Returns (list of RGB uint8 HxWx3 arrays, fps)."""
    cap = cv2.VideoCapture(video_path)
    assert cap.isOpened(), f"Failed to open video: {video_path}"
    fps = cap.get(cv2.CAP_PROP_FPS)
    frames = []
    block, used = None, 0
    idx = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        if idx >= start_frame:
            if block is None or used == _BLOCK_FRAMES or block.shape[1:] != frame.shape:
                block, used = _new_block(frame.shape), 0
            cv2.cvtColor(frame, cv2.COLOR_BGR2RGB, dst=block[used])              # tools.py:21
            frames.append(block[used])
            used += 1
            if max_frames > 0 and len(frames) >= max_frames:
                break
        idx += 1
    cap.release()
    assert len(frames) > 0, "No frames read"
    return frames, fps


def write_video_frames_to_path(out_video, mask_frames, fps, H0, W0):
    """tools.py:30-45 (FFV1 / MKV, RGB->BGR, NEAREST fix-up of off-size frames)."""
    writer = cv2.VideoWriter(out_video, cv2.VideoWriter_fourcc(*"FFV1"), fps, (W0, H0))
    assert writer.isOpened(), "Failed to open VideoWriter (FFV1/MKV). Try MJPG or mp4v if needed."
    for f in mask_frames:
        f = cv2.cvtColor(f, cv2.COLOR_RGB2BGR)
        if f.shape[0] != H0 or f.shape[1] != W0:
            f = cv2.resize(f, (W0, H0), interpolation=cv2.INTER_NEAREST)         # tools.py:41-42
        writer.write(f)
    writer.release()
    print(f"[ok] wrote {len(mask_frames)} frames to {out_video}")

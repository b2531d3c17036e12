"""CPU oracle for SURVEY "next" row N3: the SAM2 mask colour painter.  TEST INFRASTRUCTURE ONLY.

``ref_paint`` restates /root/reference/sam2_masker.py:151-175 with the same numpy / cv2 calls;
``color_for_obj`` restates :27-37.  ``reference_paint`` runs the UNMODIFIED reference function
``run_sam2_on_frames`` (build container only) with a stub predictor that replays given logits, which is
how ``tests/golden/paint.npz`` was produced (pinned).
"""
import os
import sys
import types

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

from . import reference_harness as rh


def color_for_obj(obj_id):
    """sam2_masker.py:27-37."""
    h = int((obj_id * 37) % 180)
    hsv = np.uint8([[[h, 200, 255]]])
    bgr = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)[0, 0]
    return tuple(int(x) for x in bgr)


def ref_paint(video_segments, n_frames, h0, w0):
    """sam2_masker.py:151-175.  video_segments: {frame_idx: {obj_id: bool mask}}."""
    mask_frames = []
    for idx in range(n_frames):
        masks_dict = video_segments.get(idx, {})
        out = np.zeros((h0, w0, 3), dtype=np.uint8)                                     # :155
        for obj_id in sorted(masks_dict.keys()):                                        # :159
            m = masks_dict[obj_id]
            if m is None or m.size == 0:
                continue
            m = np.asarray(m)
            if m.ndim > 2:
                m = m.squeeze()
            if m.shape != (h0, w0):
                m = cv2.resize(m.astype(np.uint8), (w0, h0), interpolation=cv2.INTER_NEAREST).astype(bool)   # :167
            else:
                m = m.astype(bool)
            out[m] = color_for_obj(int(obj_id))                                         # :171-173
        mask_frames.append(out)
    return mask_frames


def reference_paint(frames, logits_by_frame, obj_ids):
    """Run the unmodified ``sam2_masker.run_sam2_on_frames`` with a stub SAM2 predictor that yields
    ``logits_by_frame[idx]`` (float32 [K,1,mh,mw]) for ``obj_ids``."""
    import importlib.util
    import torch
    if not rh.available():
        raise RuntimeError("reference tree not present")

    class _Predictor:
        def init_state(self, video_path=None):
            return {}

        def add_new_points_or_box(self, **kw):
            return None

        def propagate_in_video(self, state):
            for idx in sorted(logits_by_frame):
                yield idx, list(obj_ids), torch.from_numpy(logits_by_frame[idx])

    m_sam = types.ModuleType("sam2")
    m_bs = types.ModuleType("sam2.build_sam")
    m_bs.build_sam2_video_predictor = lambda *a, **k: _Predictor()
    m_sam.build_sam = m_bs
    sys.modules["sam2"], sys.modules["sam2.build_sam"] = m_sam, m_bs
    saved = list(sys.path)
    sys.path.insert(0, rh.REFERENCE_DIR)
    try:
        spec = importlib.util.spec_from_file_location("_vv_reference_sam2_masker",
                                                      os.path.join(rh.REFERENCE_DIR, "sam2_masker.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
        sys.modules.pop("tools", None)
        sys.modules.pop("sam2", None)
        sys.modules.pop("sam2.build_sam", None)
    return mod.run_sam2_on_frames(frames, {"keyframes": []}, device=torch.device("cpu")), mod

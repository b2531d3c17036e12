"""CPU-side tests: the C ABI library loads and exports every symbol include/vvb200.h declares,
argument validation and error reporting work without a GPU, and the host logic (plans, sharding,
pointer marshalling, drop-in module surface) matches the oracle / the reference."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from oracle import chunk_blend as ocb
from oracle import prepost as op
from oracle import propagation as opp
from oracle import reference_harness as rh
from videovanish_b200 import _lib, chunking, diffuerase, hostpipe, ops, tools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vvb200.h")).read()
    return sorted(set(re.findall(r"VV_API[^;(]*?\b(vv_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(_lib.lib, n), "libvvb200.so does not export %s" % n
    assert sorted(_lib.EXPORTED) == names, "ctypes signatures and header out of sync"
    assert _lib.lib.vv_version() >= 100


def test_argument_validation_and_error_text():
    lib = _lib.lib
    rc = lib.vv_binarize_dilate(None, 1, 8, 8, 3, 1, None, None, 0, 0, None, 0, None)
    assert rc == -1 and "NULL" in _lib.last_error()
    h, w = ctypes.c_int(), ctypes.c_int()
    assert lib.vv_inference_size(0, 10, 960, ctypes.byref(h), ctypes.byref(w)) == -1
    with pytest.raises(RuntimeError, match="vv_inference_size"):
        ops.inference_size(-5, 10)
    assert lib.vv_chunk_blend(None, None, 1, 16, 0, 1, None, None) == -1
    assert lib.vv_resize_workspace_bytes(0, 5) == 0


@pytest.mark.parametrize("hw,s", [((1080, 1920), 960), ((360, 640), 960), ((360, 640), 320), ((2160, 3840), 960),
                                  ((1920, 1080), 960), ((720, 1280), 64), ((1000, 1000), 999)])
def test_inference_size_matches_oracle(hw, s):
    assert ops.inference_size(hw[0], hw[1], s) == op.inference_size(hw[0], hw[1], s)


def test_plans_match_oracle():
    for n in (1, 40, 50, 51, 120, 300, 5000):
        assert ops.subvideo_plan(n) == opp.subvideo_plan(n)
        assert ops.subvideo_plan(n, 7, 3) == opp.subvideo_plan(n, 7, 3)
    for n, c, o in [(600, 80, 16), (300, 80, 16), (80, 80, 16), (81, 80, 16), (10, 4, 1)]:
        assert chunking.chunk_plan(n, c, o) == ocb.chunk_plan(n, c, o)
    plan = chunking.chunk_plan(600)
    assert [s for s, _ in plan] == list(range(0, 577, 64)) and len(plan) == 10
    for world in (1, 2, 4, 8):
        sh = chunking.shard_chunks(plan, world)
        assert sum(sh, []) == list(range(10)) and max(map(len, sh)) - min(map(len, sh)) <= 1
    with pytest.raises(ValueError):
        chunking.chunk_plan(100, 16, 16)


def test_stitch_chunks_on_cpu_with_oracle_blend():
    import torch
    rng = np.random.default_rng(0)
    plan = chunking.chunk_plan(50, 20, 6)
    outs = [rng.integers(0, 256, (e - s, 6, 8, 3), dtype=np.uint8) for s, e in plan]

    def blend(tail, head, k0, total, out):
        out.copy_(torch.from_numpy(ocb.blend_overlap(tail.numpy(), head.numpy(), k0, total)))

    got = chunking.stitch_chunks([torch.from_numpy(o) for o in outs], plan, blend_fn=blend).numpy()
    assert np.array_equal(got, ocb.stitch_chunks(outs, plan, 6))


def test_pointer_marshalling():
    frames = [np.zeros((4, 6, 3), np.uint8), np.zeros((8, 6, 3), np.uint8)[::2]]      # second one is a strided view
    arr, keep = hostpipe._ptr_array(frames, (4, 6, 3))
    assert arr[0] == frames[0].ctypes.data and keep[1].flags.c_contiguous and arr[1] == keep[1].ctypes.data
    with pytest.raises(ValueError):
        hostpipe._ptr_array([np.zeros((4, 6, 3), np.uint8), np.zeros((5, 6, 3), np.uint8)], (4, 6, 3))


def test_dropin_surface_matches_reference():
    sig = inspect.signature(diffuerase.run_infill_on_frames)
    assert list(sig.parameters) == ["frames_rgb", "mask_frames", "mask_dilation_iter", "ckpt", "propainer_frames",
                                    "max_img_size", "keep_unmasked_original", "feather_px", "prog"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
    assert d == dict(mask_dilation_iter=8, ckpt="2-Step", propainer_frames=None, max_img_size=960,
                     keep_unmasked_original=True, feather_px=3, prog=None)
    if rh.available():
        ref = rh.load_reference()
        assert str(inspect.signature(ref.run_infill_on_frames)) == str(sig)
    # the reference's three parameters first, same defaults; `device=None` is an optional extension (N1)
    load = inspect.signature(tools.load_video_frames_from_path).parameters
    assert list(load)[:3] == ["video_path", "start_frame", "max_frames"]
    assert [load[k].default for k in ("start_frame", "max_frames")] == [0, -1]
    assert all(p.default is None for k, p in load.items() if k not in ("video_path", "start_frame", "max_frames"))
    assert list(inspect.signature(tools.write_video_frames_to_path).parameters) == ["out_video", "mask_frames", "fps", "H0", "W0"]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    z = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        diffuerase.run_infill_on_frames([z], [z])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.binarize_dilate(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))


def test_tools_roundtrip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    frames = [rng.integers(0, 256, (48, 64, 3), dtype=np.uint8) for _ in range(5)]
    path = str(tmp_path / "clip.mkv")
    try:
        tools.write_video_frames_to_path(path, frames, 25.0, 48, 64)
    except AssertionError:
        pytest.skip("FFV1 writer unavailable in this OpenCV build")
    got, fps = tools.load_video_frames_from_path(path, start_frame=1, max_frames=3)
    assert len(got) == 3 and abs(fps - 25.0) < 1e-6
    assert all(np.array_equal(g, f) for g, f in zip(got, frames[1:4]))          # FFV1 is lossless


def test_dropin_imports_models_lazily_like_the_reference(monkeypatch):
    """Without injected models the drop-in imports the un-vendored packages exactly where the reference
    does (diffuerase.py:8-9); their absence surfaces as ImportError, not as a silent fallback."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("the pre stage needs a GPU before the model import is reached")
    monkeypatch.setattr(diffuerase, "video_inpainting_sd", None)
    monkeypatch.setattr(diffuerase, "last_ckpt", None)
    z = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(ImportError):
        diffuerase.run_infill_on_frames([z], [z])

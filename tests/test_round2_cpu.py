"""CPU tests added in round 2: the N2 / full-path oracles, the window planners mirrored on the product side,
and the C-ABI surface (symbols only: no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.ndimage

from oracle import full_path as ofp
from oracle import propagation as opp
from oracle import wrapper as ow
from videovanish_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_neighbor_plan_matches_between_oracle_and_product():
    pytest.importorskip("torch")
    from videovanish_b200 import wrappers
    for n in (1, 4, 10, 11, 23, 50, 51, 120):
        for nl, rs, sv in ((10, 10, 50), (6, 4, 20)):
            assert wrappers.neighbor_plan(n, nl, rs, sv) == opp.neighbor_plan(n, nl, rs, sv)
    plan = opp.neighbor_plan(23)
    assert plan[0][0] == list(range(0, 6)) and plan[1][0] == list(range(0, 11)) and plan[-1][0] == list(range(15, 23))
    covered = sorted({i for ids, _ in plan for i in ids})
    assert covered == list(range(23)), "every frame is composed at least once"


def test_neighbor_merge_is_integer_average_and_select():
    """The compose loop of propainter/inference.py in closed form: masked pixels take u8(((p+1)/2)*255)
    (truncation), the others the original; a frame seen twice is (a + b) >> 1."""
    rng = np.random.default_rng(5)
    t, h, w = 12, 5, 7
    plan = opp.neighbor_plan(t, 4, 10, 50)
    preds = [rng.uniform(-1, 1, (len(ids), 3, h, w)).astype(np.float32) for ids, _ in plan]
    m = (rng.random((t, h, w)) < 0.5).astype(np.uint8)
    ori = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    got = opp.ref_neighbor_merge(preds, plan, m, ori)
    comp = [None] * t
    for (ids, _), p in zip(plan, preds):
        for i, idx in enumerate(ids):
            v = ((p[i] + np.float32(1)) * np.float32(0.5) * np.float32(255)).astype(np.float32)
            img = np.where(m[idx][..., None] > 0, np.transpose(v, (1, 2, 0)).astype(np.int32) & 255, ori[idx]).astype(np.uint8)
            comp[idx] = img if comp[idx] is None else ((comp[idx].astype(np.int32) + img) >> 1).astype(np.uint8)
    assert all(np.array_equal(a, b) for a, b in zip(got, comp))


def test_masked_frame_and_full_path_smoke():
    """oracle.full_path.run composes the stage oracles; with an empty mask the clip comes back unchanged,
    with a mask only the dilated region (+ feather band) may change."""
    t, h0, w0 = 7, 72, 128
    fr, mk = synth.frames(t, h0, w0, seed=1), synth.masks(t, h0, w0, seed=2, salt=0.0005)
    flow_fn = lambda small, low: synth.flows(len(small), small.shape[1], small.shape[2], seed=3)
    st = {}
    out = ofp.run(list(fr), list(mk), flow_fn, mask_dilation_iter=3, max_img_size=64, stages=st)
    assert len(out) == t and out[0].shape == (h0, w0, 3) and out[0].dtype == np.uint8
    dil = np.stack(st["dil"])
    untouched = np.stack([ow.model_wrapper_mask(d, 0) for d in dil]) == 0          # not even eroded-in
    far = np.stack([scipy.ndimage.binary_dilation(d > 0, iterations=4) for d in dil])        # mask + feather band
    assert np.array_equal(np.stack(out)[~far], fr[~far]), "pixels away from the mask keep the original bytes"
    assert np.array_equal(st["masked"][0][st["wrapper_mask"][0] > 0], np.zeros_like(st["masked"][0][st["wrapper_mask"][0] > 0]))
    empty = ofp.run(list(fr), [np.zeros_like(m) for m in mk], flow_fn, max_img_size=64)
    assert np.array_equal(np.stack(empty), fr)
    assert untouched.any()


def test_header_symbols_are_exported_and_bound():
    """Every VV_API declaration of include/vvb200.h is exported by the built library and bound in _lib."""
    hdr = open(os.path.join(ROOT, "include", "vvb200.h")).read()
    names = set(re.findall(r"VV_API\s+[\w\s\*]+?\b(vv_\w+)\s*\(", hdr))
    assert {"vv_neighbor_merge", "vv_apply_mask", "vv_swap_rb", "vv_binarize_dilate_ex", "vv_upscale_feather_composite_bits",
            "vv_pipeline_upload", "vv_pipeline_download"} <= names
    path = os.path.join(ROOT, "videovanish_b200", "csrc", "libvvb200.so")
    if not os.path.isfile(path):
        pytest.skip("library not built")
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n), n
    pytest.importorskip("torch")
    from videovanish_b200 import _lib
    assert names == set(_lib.EXPORTED), names ^ set(_lib.EXPORTED)


def test_synth_config_table_matches_the_inference_size_rule():
    """synth.CONFIGS names the five BASELINE configurations; their inference sizes follow the wrapper's rule
    (round 1 carried a wrong 536x952 for 4K), c2 being the one BASELINE names at 960x540 instead of 960x536."""
    from oracle import prepost as op
    for name, (t, h0, w0, hw) in synth.CONFIGS.items():
        rule = op.inference_size(h0, w0, 320 if name == "c1_360p" else 960)
        if name == "c3_720p_flow":
            rule = (h0, w0)                       # propagation runs at the clip's own resolution in config 3
        if name == "c2_1080p":
            assert hw == (540, 960) and rule == (536, 960)
        else:
            assert hw == rule, name


def test_every_tuning_option_is_documented_and_readable():
    """vv_set_option / vv_get_option work without a GPU; every option the library knows (capi.cu) is described in
    include/vvb200.h, and the documented defaults are the ones the library starts with."""
    import ctypes
    import os
    import re
    from videovanish_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    capi = open(os.path.join(root, "videovanish_b200", "csrc", "capi.cu")).read()
    names = re.search(r"g_option_names\[OPT_COUNT\] = \{([^}]*)\}", capi).group(1)
    names = re.findall(r'"([a-z0-9_]+)"', names)
    assert len(names) >= 20 and len(set(names)) == len(names)
    header = open(os.path.join(root, "include", "vvb200.h")).read()
    for n in names:
        assert '"%s"' % n in header, "option %s is not documented in include/vvb200.h" % n
        v = ctypes.c_int(-12345)
        assert _lib.lib.vv_get_option(n.encode(), ctypes.byref(v)) == 0 and v.value != -12345
    assert _lib.get_option("k4_streams") == 2 and _lib.get_option("k4_chain_ctas") == 8 and _lib.get_option("k1b_diag") == 2
    v = ctypes.c_int(0)
    assert _lib.lib.vv_get_option(b"no_such_option", ctypes.byref(v)) != 0


def test_host_chamfer_table_equals_cv2_and_the_oracle():
    """The table the feather kernels work from is built on the HOST (k3_composite.cu build_chamfer_table): without a GPU
    it can be compared with the oracle's construction and with cv2.distanceTransform around one zero pixel."""
    import ctypes
    import cv2
    from oracle import prepost as op
    from videovanish_b200 import _lib
    for r in (2, 7, 15, 31):
        n = 2 * r + 1
        buf = (ctypes.c_float * (n * n))()
        assert _lib.lib.vv_chamfer_table(r, buf) == 0
        tab = np.frombuffer(buf, np.float32).reshape(n, n)
        m = np.full((n, n), 255, np.uint8)
        m[r, r] = 0
        assert np.array_equal(tab, cv2.distanceTransform(m, cv2.DIST_L2, 5))
        assert np.array_equal(tab[::-1, ::-1], op.chamfer_cost_table(r))
    assert _lib.lib.vv_chamfer_table(32, buf) != 0


@pytest.mark.parametrize("k,extra", [(3, 2), (4, 1), (5, 1), (5, 2), (6, 1), (7, 1), (7, 2)])
def test_diagonal_segments_plus_cross_rounds_make_the_l1_ball(k, extra):
    """The identity K1's diagonal blocks rest on (k1_mask.cu diamond_block): dilating by the two diagonal segments
    {i (1,1)}, {j (1,-1)}, |i|, |j| <= K, and then by `extra` >= 1 cross rounds equals 2K + extra cross rounds, i.e. the L1
    ball of that radius - on random masks with set pixels on the frame edges (the frame is NOT padded for the reference
    side; intermediate pixels outside the frame must not matter)."""
    from oracle import prepost as op
    rng = np.random.default_rng(10 * k + extra)
    h, w = 41, 57
    m = rng.random((h, w)) < 0.004
    m[0, 0] = m[h - 1, w - 1] = m[0, w - 1] = m[h // 2, 0] = True
    want = op.model_dilate_l1(m.astype(np.uint8), 2 * k + extra) > 0
    pad = 2 * k + extra + 1                                   # the kernel's halo lanes / rows play this role
    big = np.zeros((h + 2 * pad, w + 2 * pad), bool)
    big[pad:pad + h, pad:pad + w] = m

    def shift(a, dy, dx):
        return np.roll(np.roll(a, dy, axis=0), dx, axis=1)    # the padding keeps the wrap-around out of the frame

    acc = big.copy()
    for dy_sign in (1, -1):                                   # S1 = (1, 1) direction, S2 = (1, -1)
        seg = acc.copy()
        for i in range(1, k + 1):
            seg |= shift(acc, i * dy_sign, i) | shift(acc, -i * dy_sign, -i)
        acc = seg
    for _ in range(extra):
        acc = acc | shift(acc, 1, 0) | shift(acc, -1, 0) | shift(acc, 0, 1) | shift(acc, 0, -1)
    assert np.array_equal(acc[pad:pad + h, pad:pad + w], want)

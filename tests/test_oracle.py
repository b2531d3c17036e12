"""Pin the oracle: closed-form models == the reference's library calls == the
UNMODIFIED reference run (golden vectors; live when /root/reference exists).
CPU only.  Encodes SURVEY.md KATs T1-T9."""
import glob
import os

import numpy as np
import pytest

from oracle import prepost as op
from oracle import reference_harness as rh
from videovanish_b200 import synth
from tests.golden.make_golden import inputs as golden_inputs

cv2 = pytest.importorskip("cv2")
scipy_ndimage = pytest.importorskip("scipy.ndimage")

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if os.path.basename(p) not in ("paint.npz", "other_mask_size.npz"))


def load_case(path):
    z = np.load(path)
    t, h0, w0, h, w, n, seed = [int(v) for v in z["args"]]
    fr, mk, inp = golden_inputs(t, h0, w0, h, w, seed)
    if bool(z["empty_frame1"]):
        mk[1] = 0
    return z, fr, mk, inp, n, float(z["feather"]), bool(z["keep"])


def test_golden_present():
    assert len(GOLDEN) >= 10


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    """ref_* (library restatement) and model_* (closed form) both reproduce the
    reference's own outputs bit for bit."""
    z, fr, mk, inp, n, f, keep = load_case(path)
    dil_ref = op.ref_binarize_dilate(list(mk), n)
    dil_model = op.model_binarize_dilate(list(mk), n)
    assert np.array_equal(np.stack(dil_ref), z["dilated"])
    assert np.array_equal(np.stack(dil_model), z["dilated"])
    for i in range(len(fr)):
        o_ref = op.ref_post_frame(inp[i], fr[i], dil_ref[i], keep, f)
        o_model = op.model_post_frame(inp[i], fr[i], dil_model[i], keep, f)
        assert np.array_equal(o_ref, z["out"][i])
        assert np.array_equal(o_model, z["out"][i]), "the closed-form model is bit-exact for every feather_px <= 32"
    assert bool(z["literal_rest_raw"])          # reference bug :114 - frames 1.. returned raw
    assert np.array_equal(z["literal_frame0"], z["out"][0])


def other_mask_size_case():
    """Masks of ANOTHER size than the frames (diffuerase.py:85-86), golden from the unmodified reference."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "other_mask_size.npz"))
    t, h0, w0, h, w, n, seed = [int(v) for v in z["args"]]
    hm, wm = [int(v) for v in z["mask_hw"]]
    fr, _, inp = golden_inputs(t, h0, w0, h, w, seed)
    mk = synth.masks(t, hm, wm, seed=seed + 1, salt=0.002)
    return z, fr, mk, inp, n, float(z["feather"])


def test_oracle_matches_golden_with_masks_of_another_size():
    """The reference dilates such masks at their own size and fits them with INTER_NEAREST in the post loop; both
    oracle flavours reproduce its outputs bit for bit."""
    z, fr, mk, inp, n, f = other_mask_size_case()
    assert mk.shape[1:3] != fr.shape[1:3]
    dil_ref, dil_model = op.ref_binarize_dilate(list(mk), n), op.model_binarize_dilate(list(mk), n)
    assert np.array_equal(np.stack(dil_ref), z["dilated"]) and np.array_equal(np.stack(dil_model), z["dilated"])
    for i in range(len(fr)):
        assert np.array_equal(op.ref_post_frame(inp[i], fr[i], dil_ref[i], True, f), z["out"][i])
        assert np.array_equal(op.model_post_frame(inp[i], fr[i], dil_model[i], True, f), z["out"][i])


@pytest.mark.reference
@pytest.mark.skipif(not rh.available(), reason="/root/reference not present")
def test_oracle_matches_live_reference():
    fr = synth.frames(3, 90, 150, seed=7)
    mk = synth.masks(3, 90, 150, seed=8, salt=0.003)
    inp = synth.noise_frames(3, 40, 72, seed=9)
    outs, dil = rh.ref_post_all_frames(list(fr), list(mk), list(inp), mask_dilation_iter=5, feather_px=3)
    assert np.array_equal(np.stack(dil), np.stack(op.ref_binarize_dilate(list(mk), 5)))
    got = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp],
                                      mask_dilation_iter=5, propainer_frames=list(fr))
    assert np.array_equal(np.stack(got), np.stack(outs))
    lit, _, kw = rh.ref_run_literal(list(fr), list(mk), list(inp), mask_dilation_iter=5)
    bug = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp],
                                      mask_dilation_iter=5, propainer_frames=list(fr), bug_compat=True)
    assert all(np.array_equal(a, b) for a, b in zip(lit, bug))
    assert kw == {"max_img_size": 960, "mask_dilation_iter": 0, "guidance_scale": None, "progress": None}


# ---------------------------------------------------------------- KATs T1-T3: dilation
@pytest.mark.parametrize("n", [1, 3, 8, 25])
def test_T1_dilation_is_l1_ball(n):
    rng = np.random.default_rng(n)
    m = rng.random((97, 131)) < 0.003
    ref = scipy_ndimage.binary_dilation(m, iterations=n)
    ys, xs = np.nonzero(m)
    yy, xx = np.mgrid[0:97, 0:131]
    dist = np.min(np.abs(yy[..., None] - ys) + np.abs(xx[..., None] - xs), axis=2)
    assert np.array_equal(ref, dist <= n)
    assert np.array_equal(op.model_dilate_l1(m, n), ref.astype(np.uint8) * 255)


def test_T2_iterations_zero_fills_frame():
    m = np.zeros((20, 30), bool)
    assert not scipy_ndimage.binary_dilation(m, iterations=0).any()
    assert op.model_dilate_l1(m, 0).max() == 0
    m[7, 11] = True
    assert scipy_ndimage.binary_dilation(m, iterations=0).all()
    assert op.model_dilate_l1(m, 0).min() == 255
    assert op.model_dilate_l1(m, -3).min() == 255


def test_T3_binarize_any_channel():
    rng = np.random.default_rng(3)
    m = (rng.integers(0, 256, (40, 50, 3)) * (rng.random((40, 50, 3)) < 0.05)).astype(np.uint8)
    assert np.array_equal(np.any(m > 0, axis=2).astype(np.uint8), op.model_binarize(m))


# ---------------------------------------------------------------- KATs T4/T5/T9: resize
RESIZE_CASES = [(540, 960, 1080, 1920), (536, 960, 1080, 1920), (176, 320, 360, 640), (1080, 1920, 540, 960),
                (1080, 1920, 536, 960), (97, 131, 200, 333), (200, 333, 97, 131), (7, 5, 31, 47), (1, 1, 8, 8),
                (2, 3, 9, 9), (100, 100, 25, 25), (2160, 3840, 536, 960), (536, 960, 2160, 3840), (64, 64, 64, 200), (300, 400, 150, 100)]


@pytest.mark.parametrize("sh,sw,dh,dw", RESIZE_CASES)
def test_T5_linear_model_bit_exact(sh, sw, dh, dw):
    rng = np.random.default_rng(sh * 7 + dw)
    for c in (1, 3):
        src = rng.integers(0, 256, (sh, sw, c), dtype=np.uint8)
        ref = cv2.resize(src, (dw, dh)).reshape(dh, dw, c)
        assert np.array_equal(op.model_resize_linear(src, dh, dw), ref)


def test_T4_half_scale_is_box():
    rng = np.random.default_rng(4)
    src = rng.integers(0, 256, (180, 320, 3), dtype=np.uint8).astype(np.int32)
    box = ((src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    assert np.array_equal(box, cv2.resize(src.astype(np.uint8), (160, 90)))
    assert np.array_equal(box, op.model_resize_linear(src.astype(np.uint8), 90, 160))


@pytest.mark.parametrize("sh,sw,dh,dw", RESIZE_CASES)
def test_T9_nearest_model_bit_exact(sh, sw, dh, dw):
    rng = np.random.default_rng(sh + dw)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(op.model_resize_nearest(src, dh, dw), ref)


def test_inference_size():
    assert op.inference_size(1080, 1920, 960) == (536, 960)
    assert op.inference_size(360, 640, 960) == (360, 640)
    assert op.inference_size(360, 640, 320) == (176, 320)
    assert op.inference_size(2160, 3840, 960) == (536, 960)
    assert op.inference_size(1920, 1080, 960) == (960, 536)


# ---------------------------------------------------------------- KATs T6-T8: feather + composite
@pytest.mark.parametrize("f", [1, 2, 3, 2.5, 4, 5, 8, 9, 12.5, 16, 21, 27.5, 32])
def test_T6_feather_alpha_model(f):
    """The windowed first-hit model with the table of the two raster passes equals cv2.distanceTransform's alpha bit for
    bit for every feather_px <= 32 (dense, sparse, one-object and one-hole masks: distances up to the window radius)."""
    rng = np.random.default_rng(int(f * 10))
    for dens in (0.0005, 0.01, 0.3, 0.7, 0.99, 0.9995):
        m = (rng.random((83, 117)) < dens).astype(np.uint8) * 255
        if 0.001 < dens < 0.999:
            m[:9, :13] = 255
            m[-6:, -20:] = 0
            m[30:50, 40:80] = 255
        assert np.array_equal(op.ref_feather_alpha(m, f), op.model_feather_alpha(m, f))
    for m in (np.zeros((9, 9), np.uint8), np.full((9, 9), 255, np.uint8)):
        assert np.array_equal(op.ref_feather_alpha(m, f), op.model_feather_alpha(m, f))


def test_T6_chamfer_table_is_the_two_pass_table_not_a_metric():
    """From d ~ 12 on the table of one zero pixel is not symmetric (float32 addition is not associative and the raster
    order fixes the order in which a path adds up its steps); cv2 itself shows the same asymmetry."""
    r = 31
    tab = op.chamfer_cost_table(r)
    m = np.full((2 * r + 1, 2 * r + 1), 255, np.uint8)
    m[r, r] = 0
    cv = cv2.distanceTransform(m, cv2.DIST_L2, 5)
    assert np.array_equal(tab[::-1, ::-1], cv)                 # tab[dy + r, dx + r]: the zero pixel sits at p + (dy, dx)
    assert not np.array_equal(tab, tab[::-1, ::-1]) and not np.array_equal(cv, cv.T)
    small = op.chamfer_cost_table(7)
    assert np.array_equal(small, small[::-1, ::-1]) and np.array_equal(small, small.T)      # below 8 it still is


def test_T7_alpha_levels_at_F3():
    rng = np.random.default_rng(5)
    m = (rng.random((200, 200)) < 0.4).astype(np.uint8) * 255
    m[50:120, 60:150] = 255
    m[130:190, 10:100] = 0
    lv = np.unique(op.ref_feather_alpha(m, 3))
    expect = np.array([0, .0333, .1339, .1667, .2667, .3333, .6667, .7333, .8333, .8662, .9667, 1], np.float32)
    assert len(lv) <= 12 and all(np.min(np.abs(expect - v)) < 1e-3 for v in lv)
    assert not np.any(lv == np.float32(0.5))


def test_T8_composite_model_bit_exact():
    rng = np.random.default_rng(6)
    a = rng.choice(np.unique(op.model_feather_alpha(
        (rng.random((64, 64)) < 0.4).astype(np.uint8) * 255, 3)), (120, 160)).astype(np.float32)
    out = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    orig = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    assert np.array_equal(op.ref_composite(a, out, orig), op.model_composite(a, out, orig))
    a2 = rng.random((120, 160)).astype(np.float32)
    assert np.array_equal(op.ref_composite(a2, out, orig), op.model_composite(a2, out, orig))


def test_empty_and_full_masks_post():
    fr = synth.frames(1, 48, 64, seed=1)[0]
    inp = synth.noise_frames(1, 24, 32, seed=2)[0]
    up = cv2.resize(inp, (64, 48))
    assert np.array_equal(op.ref_post_frame(inp, fr, np.zeros((48, 64), np.uint8)), fr)
    assert np.array_equal(op.ref_post_frame(inp, fr, np.full((48, 64), 255, np.uint8)), up)
    assert np.array_equal(op.model_post_frame(inp, fr, np.zeros((48, 64), np.uint8)), fr)
    assert np.array_equal(op.model_post_frame(inp, fr, np.full((48, 64), 255, np.uint8)), up)


# ---------------------------------------------------------------- next row N3: SAM2 colour painter
def _paint_case():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "paint.npz"))
    ids = [int(v) for v in z["obj_ids"]]
    segs = {i: {o: z["logits"][i][k] > 0 for k, o in enumerate(ids)} for i in range(len(z["logits"]))}
    segs_same = {i: {o: z["logits_same"][i][k] > 0 for k, o in enumerate(ids)} for i in range(len(z["logits_same"]))}
    return z, ids, segs, segs_same


def test_painter_oracle_matches_golden():
    from oracle import painter
    z, ids, segs, segs_same = _paint_case()
    h0, w0 = z["out"].shape[1:3]
    assert np.array_equal(np.stack(painter.ref_paint(segs, len(segs), h0, w0)), z["out"])
    assert np.array_equal(np.stack(painter.ref_paint(segs_same, len(segs_same), h0, w0)), z["out_same"])
    # highest object id wins where masks overlap; background stays black
    both = segs[0][ids[0]] & segs[0][ids[1]]
    up = cv2.resize(both.squeeze().astype(np.uint8), (w0, h0), interpolation=cv2.INTER_NEAREST).astype(bool)
    assert up.any() and np.all(z["out"][0][up & ~cv2.resize(segs[0][ids[2]].squeeze().astype(np.uint8), (w0, h0),
                                                          interpolation=cv2.INTER_NEAREST).astype(bool)] ==
                               np.array(painter.color_for_obj(ids[1])))


# ------------------------------------------------------------------ N4 (DiffuEraser wrapper glue): cv2 pins
@pytest.mark.parametrize("h,w", [(40, 64), (33, 47), (54, 96), (11, 11)])
@pytest.mark.parametrize("n_dilate", [0, 1, 4])
def test_N4_wrapper_models_match_cv2(h, w, n_dilate):
    """The closed forms the K7 kernels implement equal the cv2 / numpy restatement of the upstream wrapper:
    3x3 erode + dilate with cv2's default borders, the bit-exact u8 GaussianBlur((21, 21), 0) and the
    float64 -> u8 alpha truncation."""
    import cv2
    from oracle import wrapper as ow
    rng = np.random.default_rng(h * w + n_dilate)
    for dens in (0.01, 0.2, 0.6):
        m = (rng.random((h, w)) < dens).astype(np.uint8) * rng.integers(1, 256, (h, w)).astype(np.uint8)
        m = cv2.dilate(m, np.ones((3, 3), np.uint8), iterations=2)
        a = ow.ref_wrapper_mask(m, n_dilate)
        assert np.array_equal(a, ow.model_wrapper_mask(m, n_dilate))
        assert set(np.unique(a)) <= {0, 255}
        assert np.array_equal(ow.ref_soft_alpha(a), ow.model_soft_alpha(a))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        fr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for blended in (True, False):
            assert np.array_equal(ow.ref_wrapper_compose(img, fr, a, blended), ow.model_wrapper_compose(img, fr, a, blended))


def test_N4_gaussian_taps_are_cv2s():
    """GAUSS21_Q8 is OpenCV's bit-exact Q0.8 kernel for ksize 21, sigma 0: impulse responses of the 1-D blur."""
    import cv2
    from oracle import wrapper as ow
    assert int(ow.GAUSS21_Q8.sum()) == 256
    for amp in (255, 128, 37):
        img = np.zeros((1, 64), np.uint8)
        img[0, 32] = amp
        got = cv2.GaussianBlur(img, (21, 1), 0)[0, 22:43].astype(np.int64)
        assert np.array_equal(got, (amp * ow.GAUSS21_Q8 + 128) >> 8)
    lut = ow.alpha_lut()
    assert lut[0] == 0 and lut[255] == 255 and int((lut != np.arange(256)).sum()) == 42

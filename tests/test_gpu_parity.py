"""GPU parity tests (run with ``-m gpu`` on the B200 box): every call goes through the C ABI
(ctypes -> libvvb200.so) and is compared with the CPU oracle on the same seeded inputs and with
the golden vectors the unmodified reference produced (tests/golden).

Tolerances (BASELINE.json north_star): masks bit-exact; u8 frames within +-1 LSB - and in fact
asserted bit-exact wherever the closed-form model is (feather_px <= 3, every resize)."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import chunk_blend as ocb
from oracle import prepost as op
from oracle import propagation as opp
from videovanish_b200 import synth
from tests.golden.make_golden import inputs as golden_inputs

pytestmark = pytest.mark.gpu

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if os.path.basename(p) not in ("paint.npz", "other_mask_size.npz"))


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videovanish_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_pre_and_post(ops, path):
    z = np.load(path)
    t, h0, w0, h, w, n, seed = [int(v) for v in z["args"]]
    f, keep = float(z["feather"]), bool(z["keep"])
    fr, mk, inp = golden_inputs(t, h0, w0, h, w, seed)
    if bool(z["empty_frame1"]):
        mk[1] = 0
    dil = ops.binarize_dilate(dev(mk), n)
    assert np.array_equal(host(dil), z["dilated"]), "dilated masks must be bit-exact"
    out = ops.upscale_feather_composite(dev(inp), dev(fr), dil, feather_px=f, keep_unmasked_original=keep)
    assert np.array_equal(host(out), z["out"]), "composited frames must be bit-exact"


# ------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n", [1, 2, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 24, 25, 26, 31, 32, 33, 40, 70])
def test_k1_dilation_radii(ops, n):
    mk = synth.masks(2, 150, 208, seed=n, salt=0.001)
    ref = np.stack(op.model_binarize_dilate(list(mk), n))
    assert np.array_equal(host(ops.binarize_dilate(dev(mk), n)), ref)


@pytest.mark.parametrize("h,w,c", [(97, 131, 3), (64, 64, 1), (33, 1000, 4), (1, 17, 3), (40, 2048 + 16, 3), (5, 3, 3)])
def test_k1_shapes_and_channels(ops, h, w, c):
    rng = np.random.default_rng(h * w)
    mk = (rng.integers(0, 256, (3, h, w, c)) * (rng.random((3, h, w, c)) < 0.004)).astype(np.uint8)
    ref = np.stack([op.model_dilate_l1(op.model_binarize(m), 6) for m in mk])
    assert np.array_equal(host(ops.binarize_dilate(dev(mk), 6)), ref)


@pytest.mark.parametrize("diag", [1, 0, 2], ids=["diagonal-blocks", "cross-rounds", "diagonal-blocks-8"])
@pytest.mark.parametrize("h,w", [(97, 131), (40, 2048 + 16), (33, 1000), (64, 64), (9, 31)])
def test_k1_large_radii_at_frame_edges(ops, h, w, diag):
    """Radii 9..16 per pass run as a diamond of radius 2K built from two diagonal segments by doubling, plus cross
    rounds (k1b_diag): set pixels in the corners, along the edges and at word / tile boundaries, every radius that
    picks another block, with and without the blocks."""
    from videovanish_b200 import _lib
    rng = np.random.default_rng(h + w)
    mk = np.zeros((3, h, w, 1), np.uint8)
    for y, x in ((0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, 0), (h // 2, w - 1), (0, w // 2), (h - 1, w // 2)):
        mk[0, y, x] = 255
    for x in (31, 32, 63, 64, 959, 960, 961, 991, 992):          # word and 30-word tile boundaries
        if x < w:
            mk[1, rng.integers(0, h), x] = 7
    mk[2] = (rng.random((h, w, 1)) < 0.002) * 255
    try:
        _lib.set_option("k1b_diag", diag)
        for n in (8, 9, 10, 11, 12, 13, 14, 15, 16, 25, 29):
            ref = np.stack([op.model_dilate_l1(op.model_binarize(m), n) for m in mk])
            assert np.array_equal(host(ops.binarize_dilate(dev(mk), n)), ref), n
    finally:
        _lib.set_option("k1b_diag", 2)


def test_k1_iterations_zero_fills(ops):
    mk = synth.masks(3, 72, 96, seed=1, salt=0.0005)
    mk[1] = 0
    for it in (0, -2):
        got = host(ops.binarize_dilate(dev(mk), it))
        assert got[0].min() == 255 and got[2].min() == 255 and got[1].max() == 0


def test_k1_scipy_cross_check(ops):
    mk = synth.masks(2, 120, 176, seed=3, salt=0.002)
    ref = np.stack(op.ref_binarize_dilate(list(mk), 8))
    assert np.array_equal(host(ops.binarize_dilate(dev(mk), 8)), ref)


def test_k1_fused_lowres_mask(ops):
    mk = synth.masks(2, 180, 320, seed=4, salt=0.002)
    full, low = ops.binarize_dilate(dev(mk), 8, lowres_size=(88, 160))
    ref_full = np.stack(op.model_binarize_dilate(list(mk), 8))
    assert np.array_equal(host(full), ref_full)
    assert np.array_equal(host(low), np.stack([op.ref_resize_nearest(m, 88, 160) for m in ref_full]))
    full0, low0 = ops.binarize_dilate(dev(mk), 0, lowres_size=(88, 160))
    assert host(low0).min() == 255


# ------------------------------------------------------------------------------- K2
RESIZE = [(540, 960, 1080, 1920), (176, 320, 360, 640), (1080, 1920, 540, 960), (1080, 1920, 536, 960),
          (97, 131, 200, 333), (200, 333, 97, 131), (7, 5, 31, 47), (1, 1, 8, 8), (360, 640, 176, 320),
          (64, 64, 64, 200), (300, 400, 150, 100), (50, 70, 50, 70),
          (270, 480, 67, 120), (128, 256, 32, 64), (135, 3840, 34, 960)]        # W == 4w: the 4K -> 960-wide path


@pytest.mark.parametrize("sh,sw,dh,dw", RESIZE)
def test_k2_linear_bit_exact_vs_cv2(ops, sh, sw, dh, dw):
    rng = np.random.default_rng(sh + 3 * dw)
    for c in (3, 1, 4):
        src = rng.integers(0, 256, (2, sh, sw, c), dtype=np.uint8)
        ref = np.stack([op.ref_resize_linear(s, dh, dw).reshape(dh, dw, c) for s in src])
        assert np.array_equal(host(ops.resize(dev(src), dh, dw, ops.INTER_LINEAR)), ref)


@pytest.mark.parametrize("sh,sw,dh,dw", RESIZE)
def test_k2_nearest_bit_exact_vs_cv2(ops, sh, sw, dh, dw):
    rng = np.random.default_rng(sh + dw)
    src = rng.integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
    ref = np.stack([op.ref_resize_nearest(s, dh, dw) for s in src])
    assert np.array_equal(host(ops.resize(dev(src), dh, dw, ops.INTER_NEAREST)), ref)


@pytest.mark.parametrize("sh,sw,dh,dw", RESIZE + [(64, 512, 16, 128), (33, 64, 17, 32)])
def test_k2_nearest_masks_bit_exact_vs_cv2(ops, sh, sw, dh, dw):
    """Single-channel NEAREST (the dilated masks): the exact x2 / x4 column paths and the gather kernel."""
    rng = np.random.default_rng(sh + 7 * dw)
    src = ((rng.random((3, sh, sw)) < 0.3) * rng.integers(1, 256, (3, sh, sw))).astype(np.uint8)
    ref = np.stack([op.ref_resize_nearest(s, dh, dw) for s in src])
    assert np.array_equal(host(ops.resize(dev(src), dh, dw, ops.INTER_NEAREST)), ref)


# ------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("f", [3, 1, 2, 2.5, 0, -1, 0.5, 3.5, 4, 5, 6.5, 8])
def test_k3_feather_values(ops, f):
    """Every feather width is bit-exact against the cv2 / numpy reference stage, including the generic-radius
    path (feather_px > 3: chamfer windows up to 15x15)."""
    fr = synth.frames(2, 120, 176, seed=11)
    inp = synth.noise_frames(2, 56, 88, seed=12)
    dil = np.stack(op.model_binarize_dilate(list(synth.masks(2, 120, 176, seed=13, salt=0.003)), 3))
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, f) for i in range(2)])
    got = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("variant", [dict(k3_tma=1), dict(k3_tma=0, k3_nt=1), dict(k3_tma=0, k3_nt=2)],
                         ids=["tma", "regs-nt1", "regs-nt2"])
def test_k3_kernel_variants(ops, variant):
    """Every K3 variant (TMA-staged strip / register pass-through) gives the same bytes."""
    from videovanish_b200 import _lib
    fr = synth.frames(3, 200, 320, seed=41)
    inp = synth.noise_frames(3, 96, 160, seed=42)
    dil = np.stack(op.model_binarize_dilate(list(synth.masks(3, 200, 320, seed=43, salt=0.003)), 4))
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, 3) for i in range(3)])
    ref5 = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, 5) for i in range(3)])
    try:
        for k, v in variant.items():
            _lib.set_option(k, v)
        got = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=3))
        got5 = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=5))
    finally:
        _lib.set_option("k3_tma", 1)
        _lib.set_option("k3_nt", 2)
    assert np.array_equal(got, ref)
    assert np.array_equal(got5, ref5)


@pytest.mark.parametrize("h,w", [(60, 88), (8, 8), (33, 40), (540, 960)])
@pytest.mark.parametrize("f", [3, 2.5, 1.5, 0.5])
def test_k3_exact_x2_worker(ops, h, w, f):
    """W0 == 2w and H0 == 2h takes the closed-form x2 worker: same bytes as the oracle and as the generic
    tap-table worker, including the clamped border columns / rows (masks that touch all four borders)."""
    from videovanish_b200 import _lib
    h0, w0, t = 2 * h, 2 * w, 3
    fr = synth.frames(t, h0, w0, seed=61)
    inp = synth.noise_frames(t, h, w, seed=62)
    dil = np.stack(op.model_binarize_dilate(list(synth.masks(t, h0, w0, seed=63, salt=0.004)), 2))
    dil[1] = 255                                             # everything inside: every quad, all borders
    dil[2, :3] = dil[2, -2:] = 255
    dil[2, :, :5] = dil[2, :, -3:] = 255
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, f) for i in range(t)])
    assert _lib.get_option("k3_x2") == 3                     # default: the k3_fastw kernel
    fast = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
    try:
        _lib.set_option("k3_x2", 1)                          # round-1 closed-form x2 worker
        got = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
        _lib.set_option("k3_x2", 0)
        generic = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
    finally:
        _lib.set_option("k3_x2", 3)
    assert np.array_equal(fast, ref)
    assert np.array_equal(got, ref)
    assert np.array_equal(generic, ref)


@pytest.mark.parametrize("h0,w0,h,w", [(97, 131, 40, 56), (360, 640, 176, 320), (72, 128, 72, 128), (50, 1040, 24, 520),
                                        (35, 16, 70, 32), (1, 16, 1, 8), (20, 4096, 10, 2048)])
def test_k3_shapes(ops, h0, w0, h, w):
    fr = synth.frames(2, h0, w0, seed=h0)
    inp = synth.noise_frames(2, h, w, seed=w0)
    rng = np.random.default_rng(h0 * w0)
    dil = ((rng.random((2, h0, w0)) < 0.3) * 255).astype(np.uint8)
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, 3) for i in range(2)])
    got = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=3))
    assert np.array_equal(got, ref)


def test_k3_empty_full_masks_and_no_keep(ops):
    fr = synth.frames(2, 64, 96, seed=1)
    inp = synth.noise_frames(2, 32, 48, seed=2)
    up = np.stack([op.ref_resize_linear(i, 64, 96) for i in inp])
    z = np.zeros((2, 64, 96), np.uint8)
    assert np.array_equal(host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(z))), fr)
    assert np.array_equal(host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(z + 255))), up)
    assert np.array_equal(host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(z), keep_unmasked_original=False)), up)


@pytest.mark.parametrize("f", [8.5, 9, 12.5, 16, 24, 31.5, 32])
@pytest.mark.parametrize("h0,w0,h,w,bits", [(97, 131, 40, 56, False), (120, 176, 60, 88, True), (72, 128, 72, 128, False),
                                             (64, 4096 + 52, 32, 2048 + 26, False)])
def test_k3_big_feather(ops, h0, w0, h, w, bits, f):
    """feather_px in (8, 32]: k3_bigfeather walks the cost-sorted table of the two raster passes (not symmetric from
    d ~ 12 on) - bit-exact against the reference's own cv2.distanceTransform calls, on ragged sizes, with masks touching
    all borders, from the u8 mask and from K1's bit plane."""
    t = 4
    fr = synth.frames(t, h0, w0, seed=int(f * 2))
    inp = synth.noise_frames(t, h, w, seed=int(f * 2) + 1)
    mk = synth.masks(t, h0, w0, seed=int(f * 2) + 2, salt=0.0004)
    mk[1] = 0
    mk[1, h0 // 3:h0 // 3 + 9, w0 // 2:w0 // 2 + 40] = 255            # one object, far from everything else
    mk[2] = 255
    mk[2, h0 // 2, w0 // 3] = 0                                       # one hole in a full mask
    mk[3, :2] = mk[3, -1:] = 255
    mk[3, :, :3] = mk[3, :, -2:] = 255
    d_dil, _, d_bits = ops.binarize_dilate(dev(mk), 2, return_bits=True)
    dil = host(d_dil)
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, f) for i in range(t)])
    got = host(ops.upscale_feather_composite(dev(inp), dev(fr), d_dil, feather_px=f, mask_bits=d_bits if bits else None))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("f", [3.5, 5, 6.5, 8])
def test_k3_table_walking_kernel_on_small_radii(ops, f):
    """k3_big_from = 3 sends the radii 3..7 to k3_bigfeather as well: same bytes as the generic kernel and the oracle."""
    from videovanish_b200 import _lib
    t, h0, w0, h, w = 3, 97, 131, 40, 56
    fr, inp = synth.frames(t, h0, w0, seed=3), synth.noise_frames(t, h, w, seed=4)
    dil = np.stack(op.model_binarize_dilate(list(synth.masks(t, h0, w0, seed=5, salt=0.002)), 2))
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, f) for i in range(t)])
    generic = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
    try:
        _lib.set_option("k3_big_from", 3)
        walked = host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), feather_px=f))
    finally:
        _lib.set_option("k3_big_from", 8)
    assert np.array_equal(generic, ref) and np.array_equal(walked, ref)


def test_k3_unsupported_feather_raises(ops):
    fr = synth.frames(1, 32, 32, seed=1)
    with pytest.raises(RuntimeError, match="feather_px"):
        ops.upscale_feather_composite(dev(fr), dev(fr), dev(fr[..., 0]), feather_px=50)


# ------------------------------------------------------------------------------- K4
def prop_clip(t, h, w, seed, shift=0.0):
    fr = synth.frames(t, h, w, seed=seed)
    m = (synth.masks(t, h, w, seed=seed + 1, salt=0.002).max(axis=3) > 0).astype(np.uint8) * 255
    ff, fb = synth.flows(t, h, w, seed=seed + 2)
    ff[..., 0] += shift
    fb[..., 0] -= shift
    return fr, m, ff, fb


@pytest.mark.parametrize("t,h,w,shift", [(6, 48, 64, 0.0), (5, 40, 56, 9.0), (12, 72, 96, 0.0), (2, 33, 47, 0.0),
                                         (1, 16, 16, 0.0), (7, 35, 51, 0.0)])
def test_k4_matches_model_and_torch(ops, t, h, w, shift):
    fr, m, ff, fb = prop_clip(t, h, w, seed=t * 100 + h, shift=shift)
    got = host(ops.propagate(dev(fr), dev(m), dev(ff), dev(fb))).view(np.uint32)
    assert np.array_equal(got, opp.model_propagate(fr, m, ff, fb)), "packed state must equal the explicit model"
    ref_frames, ref_masks = opp.img_propagation_torch(fr, m, ff, fb)
    gf, gm = opp.decode_state(got)
    assert (gm != ref_masks).mean() <= 1e-5 and (gf != ref_frames).mean() <= 1e-5


def test_k4_zero_padding_fill(ops):
    """Holes on the left border with a small outward flow are filled from the zero padding."""
    t, h, w = 3, 24, 32
    fr = synth.frames(t, h, w, seed=5)
    m = np.zeros((t, h, w), np.uint8)
    m[1, 4:20, 0:2] = 255        # holes in the middle frame only, so the neighbour frames are known there
    ff = np.zeros((t - 1, h, w, 2), np.float32)
    fb = np.zeros((t - 1, h, w, 2), np.float32)
    ff[..., 0] = -0.61
    fb[..., 0] = 0.58
    want = opp.model_propagate(fr, m, ff, fb)
    assert ((want >> 24) == 2).sum() > 0
    got = host(ops.propagate(dev(fr), dev(m), dev(ff), dev(fb))).view(np.uint32)
    assert np.array_equal(got, want)
    rf, rm = opp.img_propagation_torch(fr, m, ff, fb)
    gf, gm = opp.decode_state(got)
    assert np.array_equal(gf, rf) and np.array_equal(gm, rm)


@pytest.mark.parametrize("persist", [0, 1], ids=["step-launches", "persistent"])
@pytest.mark.parametrize("variant", [dict(k4_pdl=1), dict(k4_pdl=0), dict(k4_lean=8, k4_step_ctas=0),
                                     dict(k4_lean=6, k4_precheck=1, k4_step_ctas=1), dict(k4_precheck=1), dict(k4_precheck=1, k4_npt=2, k4_lean=4), dict(k4_npt=2, k4_lean=3, k4_step_ctas=3),
                                     dict(k4_npt=2, k4_lean=4, k4_step_ctas=4, k4_speculate=0),
                                     dict(k4_pack_ctas=1, k4_pack_occ=4), dict(k4_pack_ctas=1024, k4_pack_occ=6),
                                     dict(k4_streams=1), dict(k4_streams=3, k4_chain_ctas=6), dict(k4_streams=4, k4_pdl=0),
                                     dict(k4_streams=2, k4_precheck=1, k4_chain_ctas=1)],
                         ids=["lean-pdl", "lean", "lean8-wide-grid", "lean6-precheck", "precheck", "precheck-npt2", "two-per-trip-3", "two-per-trip-4",
                              "pack-few-ctas", "pack-many-ctas", "one-chain", "three-chains", "four-chains-no-pdl", "two-chains-precheck"])
def test_k4_kernel_variants(ops, variant, persist):
    """Every step-kernel / pack-kernel variant gives the same state (and the pad frames kept in scratch
    never leak into the [N,h,w] result), with the scan as one launch per step (the default) or as one persistent
    cooperative launch with grid barriers (k4_persist = 1: measured slower, kept as an option; the step-kernel options
    then only choose its grid)."""
    from videovanish_b200 import _lib
    fr, m, ff, fb = prop_clip(26, 96, 160, seed=91)
    want = opp.model_propagate_clip(fr, m, ff, fb, subvideo_length=8, pad_len=3)
    assert _lib.get_option("k4_persist") == 0
    streams_default = _lib.get_option("k4_streams")
    try:
        _lib.set_option("k4_persist", persist)
        for k, v in variant.items():
            _lib.set_option(k, v)
        got = host(ops.propagate(dev(fr), dev(m), dev(ff), dev(fb), subvideo_length=8, pad_len=3)).view(np.uint32)
    finally:
        for k, v in dict(k4_pdl=1, k4_npt=1, k4_lean=5, k4_precheck=0, k4_step_ctas=5, k4_pack_ctas=128, k4_pack_occ=4,
                         k4_speculate=1, k4_persist=0, k4_streams=streams_default, k4_chain_ctas=8).items():
            _lib.set_option(k, v)
    assert np.array_equal(got, want)


def test_k4_keep_pads_returns_raw_windows(ops):
    """keep_pads=True: the windows, pad frames included, concatenated; the kept frames of each window are
    the same bytes the default (pads discarded, written in place) call returns."""
    from videovanish_b200 import ops as vops
    fr, m, ff, fb = prop_clip(23, 32, 48, seed=78)
    args = (dev(fr), dev(m), dev(ff), dev(fb))
    raw = host(ops.propagate(*args, subvideo_length=6, pad_len=2, keep_pads=True)).view(np.uint32)
    got = host(ops.propagate(*args, subvideo_length=6, pad_len=2)).view(np.uint32)
    plan = vops.subvideo_plan(23, 6, 2)
    assert raw.shape[0] == sum(e - s for s, e, _, _ in plan) and got.shape[0] == 23
    off, kept = 0, []
    for s, e, ps, pe in plan:
        kept.append(raw[off + ps: off + (e - s) - pe])
        off += e - s
    assert np.array_equal(np.concatenate(kept), got)
    buf = torch.empty((23, 32, 48), dtype=torch.int32, device="cuda")
    assert ops.propagate(*args, subvideo_length=6, pad_len=2, out=buf) is buf
    assert np.array_equal(host(buf).view(np.uint32), got)


def test_k4_subvideo_windows(ops):
    fr, m, ff, fb = prop_clip(23, 32, 48, seed=77)
    got = host(ops.propagate(dev(fr), dev(m), dev(ff), dev(fb), subvideo_length=6, pad_len=2)).view(np.uint32)
    assert np.array_equal(got, opp.model_propagate_clip(fr, m, ff, fb, subvideo_length=6, pad_len=2))
    rgb, hole = ops.propagate_unpack(dev(got.view(np.int32)), zero_level=127)
    assert np.array_equal(host(hole) > 0, (got >> 24) & 1 == 1)


# ------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("o,shape", [(16, (36, 64, 3)), (3, (17, 13, 3)), (1, (8, 8, 1))])
def test_k5_chunk_blend(ops, o, shape):
    rng = np.random.default_rng(o)
    a = rng.integers(0, 256, (o,) + shape, dtype=np.uint8)
    b = rng.integers(0, 256, (o,) + shape, dtype=np.uint8)
    assert np.array_equal(host(ops.chunk_blend(dev(a), dev(b))), ocb.blend_overlap(a, b))
    if o > 2:       # a slice of a longer overlap (what a rank holding half the overlap computes)
        got = host(ops.chunk_blend(dev(a[1:]), dev(b[1:]), k0=1, overlap_total=o))
        assert np.array_equal(got, ocb.blend_overlap(a, b)[1:])


# ------------------------------------------------------------------------------- drop-in + pipeline
class _StubDiffuEraser:
    def __init__(self, inpainted):
        self.inpainted, self.seen = inpainted, None

    def forward(self, frames, masks, priors, **kw):
        self.seen = dict(masks=[m.copy() for m in masks], kw=kw, n=len(frames))
        return [f.copy() for f in self.inpainted]


@pytest.mark.parametrize("pinned_inputs", [False, True])
def test_dropin_run_infill_on_frames(ops, pinned_inputs):
    from videovanish_b200 import diffuerase as vvd, hostpipe
    t, h0, w0 = 11, 180, 320
    h, w = ops.inference_size(h0, w0, 160)
    fr, mk, inp = synth.frames(t, h0, w0, seed=21), synth.masks(t, h0, w0, seed=22), synth.noise_frames(t, h, w, seed=23)
    frames, masks = list(fr), list(mk)
    if pinned_inputs:
        frames = hostpipe.pinned_frames(t, (h0, w0, 3))
        masks = hostpipe.pinned_frames(t, (h0, w0, 3))
        for i in range(t):
            frames[i][...] = fr[i]
            masks[i][...] = mk[i]
    stub = _StubDiffuEraser(list(inp))
    vvd.set_models(diffueraser=stub)
    calls = []
    out = vvd.run_infill_on_frames(frames, masks, mask_dilation_iter=5, propainer_frames=frames, max_img_size=160,
                                   prog=lambda p, s: calls.append((p, s)))
    ref = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp],
                                      mask_dilation_iter=5, propainer_frames=list(fr), max_img_size=160)
    assert isinstance(out, list) and len(out) == t
    assert all(o.dtype == np.uint8 and o.flags.c_contiguous and o.shape == (h0, w0, 3) for o in out)
    assert np.array_equal(np.stack(out), np.stack(ref))
    assert np.array_equal(np.stack(stub.seen["masks"]), np.stack(op.ref_binarize_dilate(list(mk), 5)))
    kw = dict(stub.seen["kw"])
    assert callable(kw.pop("progress"))
    assert kw == {"max_img_size": 160, "mask_dilation_iter": 0, "guidance_scale": None}
    assert [c[0] for c in calls] == [5, 10, 50, 90]
    assert np.array_equal(np.stack(frames), fr), "inputs must not be mutated"
    vvd.BUG_COMPAT = True
    try:
        lit = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=5, propainer_frames=list(fr), max_img_size=160)
    finally:
        vvd.BUG_COMPAT = False
    assert np.array_equal(lit[0], ref[0]) and all(np.array_equal(a, b) for a, b in zip(lit[1:], inp[1:]))


def test_pipeline_batches_and_downsize(ops):
    from videovanish_b200 import hostpipe
    t, h0, w0, h, w = 21, 90, 160, 40, 80
    fr, mk, inp = synth.frames(t, h0, w0, seed=31), synth.masks(t, h0, w0, seed=32), synth.noise_frames(t, h, w, seed=33)
    pipe = hostpipe.HostPipeline(h0, w0, frames_per_batch=4, n_slots=2)
    dil, low = pipe.pre(list(mk), 3, lowres_size=(h, w))
    ref_dil = op.ref_binarize_dilate(list(mk), 3)
    assert np.array_equal(np.stack(dil), np.stack(ref_dil))
    assert np.array_equal(np.stack(low), np.stack([op.ref_resize_nearest(m, h, w) for m in ref_dil]))
    small = pipe.downsize(list(fr), h, w)
    assert np.array_equal(np.stack(small), np.stack([op.ref_resize_linear(f, h, w) for f in fr]))
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], ref_dil[i], True, 3) for i in range(t)])
    assert np.array_equal(np.stack(pipe.post(list(inp), list(fr))), ref)                 # resident masks
    assert np.array_equal(np.stack(pipe.post(list(inp), list(fr), dilated=ref_dil)), ref)  # supplied masks
    pipe.close()


# ------------------------------------------------------------------------------- full-size properties
def test_full_size_1080p_properties(ops):
    """BASELINE config 2 shape (a 12-frame slice): oracle comparison on 2 frames, plus
    size-independent properties on all of them."""
    t, h0, w0, h, w = 12, 1080, 1920, 540, 960
    fr, mk, inp = synth.frames(t, h0, w0, seed=2), synth.masks(t, h0, w0, seed=3), synth.noise_frames(t, h, w, seed=4)
    dmk, dfr, dinp = dev(mk), dev(fr), dev(inp)
    dil = ops.binarize_dilate(dmk, 8)
    hd = host(dil)
    for i in (0, t - 1):
        assert np.array_equal(hd[i], op.ref_binarize_dilate([mk[i]], 8)[0])
    # idempotence of binarisation, monotonicity and composition of L1 balls
    assert np.array_equal(host(ops.binarize_dilate(dil, 0 + 1)), host(ops.binarize_dilate(dmk, 9)))
    assert np.array_equal(host(ops.binarize_dilate(ops.binarize_dilate(dmk, 3), 5)), hd)
    assert np.all(hd >= (mk.max(axis=3) > 0) * 255)
    out = ops.upscale_feather_composite(dinp, dfr, dil, 3)
    ho = host(out)
    for i in (0, t - 1):
        assert np.array_equal(ho[i], op.ref_post_frame(inp[i], fr[i], hd[i], True, 3))
    # outside the feathered mask the original must come back untouched; deep inside, the resized frame
    # (window radius 2 at feather 3: every pixel with alpha > 0 is within L1 distance 4 of the mask)
    far = host(ops.binarize_dilate(dil, 4)) == 0
    assert np.array_equal(ho[far], fr[far])
    small = ops.resize(dfr, h, w)
    for i in (0, t - 1):
        assert np.array_equal(host(small[i]), op.ref_resize_linear(fr[i], h, w))
    # linearity check of the box filter: resize(255 - x) == 255 - resize(x) up to rounding
    inv = host(ops.resize(255 - dfr, h, w)).astype(int)
    assert np.abs((255 - inv) - host(small).astype(int)).max() <= 1


# ------------------------------------------------------------------------------- remaining BASELINE configs
def test_config3_720p_propagation_slice(ops):
    """BASELINE config 3 shape (720p, bidirectional flow): an 8-frame slice against the explicit model,
    plus size-independent properties (holes only shrink; known pixels never change)."""
    t, h, w = 8, 720, 1280
    fr, m, ff, fb = prop_clip(t, h, w, seed=33)
    got = host(ops.propagate(dev(fr), dev(m), dev(ff), dev(fb))).view(np.uint32)
    assert np.array_equal(got, opp.model_propagate(fr, m, ff, fb))
    hole_after = (got >> 24) & 1
    assert np.all(hole_after <= (m > 0)) and hole_after.sum() < (m > 0).sum()
    known = m == 0
    assert np.array_equal(got[known], opp.pack_state(fr, m)[known])


def test_config4_chunked_stitch(ops):
    """BASELINE config 4 scheme (chunk 80 / overlap 16, here 20 / 6 on a small clip): per-chunk outputs
    stitched with K5 == the oracle's stitch."""
    from videovanish_b200 import chunking
    rng = np.random.default_rng(4)
    plan = chunking.chunk_plan(50, 20, 6)
    outs = [rng.integers(0, 256, (e - s, 36, 64, 3), dtype=np.uint8) for s, e in plan]
    got = host(chunking.stitch_chunks([dev(o) for o in outs], plan))
    assert np.array_equal(got, ocb.stitch_chunks(outs, plan, 6))


def test_config5_long_clip_streams_through_pipeline(ops):
    """BASELINE config 5 idea (a clip much longer than one batch): 70 frames through the host pipeline
    in batches of 8 over 3 slots, pageable inputs (staging ring) - same bytes as the oracle."""
    from videovanish_b200 import hostpipe
    t, h0, w0, h, w = 70, 72, 128, 32, 64
    fr, mk, inp = synth.frames(t, h0, w0, seed=51), synth.masks(t, h0, w0, seed=52), synth.noise_frames(t, h, w, seed=53)
    pipe = hostpipe.HostPipeline(h0, w0)
    dil = pipe.pre([m.copy() for m in mk], 4)
    ref_dil = op.ref_binarize_dilate(list(mk), 4)
    assert np.array_equal(np.stack(dil), np.stack(ref_dil))
    out = pipe.post([x.copy() for x in inp], [f.copy() for f in fr])
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], ref_dil[i], True, 3) for i in range(t)])
    assert np.array_equal(np.stack(out), ref)
    pipe.close()


def test_empty_and_degenerate_inputs(ops):
    """Ragged / degenerate shapes the reference would accept: single pixel rows, 1-frame clips,
    all-masked and unmasked frames mixed in one batch."""
    fr = synth.frames(3, 2, 16, seed=1)
    inp = synth.noise_frames(3, 1, 8, seed=2)
    dil = np.zeros((3, 2, 16), np.uint8)
    dil[1] = 255
    dil[2, 0, 3] = 255
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, 3) for i in range(3)])
    assert np.array_equal(host(ops.upscale_feather_composite(dev(inp), dev(fr), dev(dil), 3)), ref)
    one = synth.masks(1, 1, 1, seed=3, salt=0)
    one[...] = 7
    assert host(ops.binarize_dilate(dev(one), 8)).tolist() == [[[255]]]
    with pytest.raises(RuntimeError):
        ops.binarize_dilate(dev(np.zeros((0, 4, 4, 3), np.uint8)), 1)


# ------------------------------------------------------------------------------- next rows N3 / N2
def test_n3_painter_matches_reference_golden(ops):
    from oracle import painter
    from videovanish_b200 import sam2_masker as vsm
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "paint.npz"))
    ids = [int(v) for v in z["obj_ids"]]
    h0, w0 = z["out"].shape[1:3]
    assert [vsm.color_for_obj(o) for o in ids] == [painter.color_for_obj(o) for o in ids]
    # float logits straight from the model (T,K,1,mh,mw), resized by the kernel
    got = host(ops.paint_masks(dev(z["logits"][:, :, 0]), [painter.color_for_obj(o) for o in ids], out_size=(h0, w0)))
    assert np.array_equal(got, z["out"])
    got_same = host(ops.paint_masks(dev((z["logits_same"][:, :, 0] > 0).astype(np.uint8)),
                                    [painter.color_for_obj(o) for o in ids]))
    assert np.array_equal(got_same, z["out_same"])
    # the dict-of-dicts front-end, with a missing object on one frame and odd sizes
    rng = np.random.default_rng(9)
    segs = {i: {o: rng.random((33, 47)) < 0.2 for o in (3, 1, 12)} for i in range(3)}
    del segs[1][12]
    want = painter.ref_paint(segs, 4, 67, 95)
    have = vsm.paint_mask_frames(segs, 4, 67, 95)
    assert all(np.array_equal(a, b) for a, b in zip(want, have))


@pytest.mark.parametrize("k", [3, 10])
@pytest.mark.parametrize("mh,mw,h0,w0", [(40, 64, 40, 64), (36, 48, 72, 96), (25, 32, 61, 64), (40, 64, 80, 96)])
def test_n3_painter_row_fast_paths(ops, k, mh, mw, h0, w0):
    """Same-width and exact-x2 canvases take the row-vectorised kernel (the last shape the generic one):
    bytes equal the reference painter's, for float logits and for u8 masks."""
    from oracle import painter
    rng = np.random.default_rng(mh * 7 + k)
    t = 3
    logits = (rng.standard_normal((t, k, mh, mw)) - 0.8).astype(np.float32)
    logits[1, k - 1] = 1.0                                   # the last object covers a whole frame
    ids = list(range(1, k + 1))
    colors = [painter.color_for_obj(o) for o in ids]
    segs = {i: {o: logits[i, j] > 0 for j, o in enumerate(ids)} for i in range(t)}
    want = np.stack(painter.ref_paint(segs, t, h0, w0))
    assert np.array_equal(host(ops.paint_masks(dev(logits), colors, out_size=(h0, w0))), want)
    as_u8 = (logits > 0).astype(np.uint8) * 255
    assert np.array_equal(host(ops.paint_masks(dev(as_u8), colors, out_size=(h0, w0))), want)


@pytest.mark.parametrize("h,w", [(540, 960), (54, 96), (33, 47), (40, 64), (11, 11), (70, 1100)])
@pytest.mark.parametrize("n_dilate", [0, 4])
def test_n4_wrapper_mask_and_compose(ops, h, w, n_dilate):
    """N4: read_mask (erode + dilate) and the blurred compose equal the cv2 / numpy restatement byte for byte,
    including masks that touch the borders, empty and full frames."""
    from oracle import wrapper as ow
    t = 4
    rng = np.random.default_rng(h + 3 * w + n_dilate)
    mk = np.stack(op.model_binarize_dilate(list(synth.masks(t, h, w, seed=h + n_dilate, salt=0.004)), 2))
    mk[1] = 0
    mk[2] = 255
    mk[3, :3] = 200                                  # non-{0,255} values, touching all four borders
    mk[3, -2:] = 7
    mk[3, :, :4] = 255
    mk[3, :, -1:] = 1
    got_mask = host(ops.wrapper_mask(dev(mk), n_dilate))
    want_mask = np.stack([ow.ref_wrapper_mask(m, n_dilate) for m in mk])
    assert np.array_equal(got_mask, want_mask)
    img = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    fr = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    for blended in (True, False):
        got = host(ops.wrapper_compose(dev(img), dev(fr), dev(want_mask), blended))
        want = np.stack([ow.ref_wrapper_compose(img[i], fr[i], want_mask[i], blended) for i in range(t)])
        assert np.array_equal(got, want), (blended, int(np.abs(got.astype(int) - want.astype(int)).max()))


def test_n4_argument_checks(ops):
    small = torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="not supported"):
        ops.wrapper_compose(small, small, small[..., 0].contiguous(), True)
    assert ops.wrapper_compose(small, small, small[..., 0].contiguous(), False).shape == small.shape


def test_n2_state_to_float(ops):
    fr, m, ff, fb = prop_clip(5, 36, 52, seed=61)
    packed = ops.propagate(dev(fr), dev(m), dev(ff), dev(fb))
    rgb, hole = ops.propagate_to_float(packed)
    want_rgb, want_hole = opp.decode_state(host(packed).view(np.uint32))
    assert np.array_equal(host(rgb), want_rgb) and np.array_equal(host(hole), want_hole)
    ref_rgb, ref_hole = opp.img_propagation_torch(fr, m, ff, fb)
    assert np.array_equal(host(rgb), ref_rgb) and np.array_equal(host(hole), ref_hole)


def test_pageable_results_when_pinned_budget_exceeded(ops):
    """Results above the pinned budget are returned in ordinary host memory and still match."""
    from videovanish_b200 import hostpipe
    t, h0, w0, h, w = 9, 64, 96, 32, 48
    fr, mk, inp = synth.frames(t, h0, w0, seed=71), synth.masks(t, h0, w0, seed=72), synth.noise_frames(t, h, w, seed=73)
    old = hostpipe.PINNED_RESULT_LIMIT
    hostpipe.PINNED_RESULT_LIMIT = 0
    try:
        pipe = hostpipe.HostPipeline(h0, w0, frames_per_batch=4, n_slots=2)
        dil = pipe.pre(list(mk), 2)
        out = pipe.post(list(inp), list(fr))
        pipe.close()
    finally:
        hostpipe.PINNED_RESULT_LIMIT = old
    ref_dil = op.ref_binarize_dilate(list(mk), 2)
    assert np.array_equal(np.stack(dil), np.stack(ref_dil))
    assert np.array_equal(np.stack(out), np.stack([op.ref_post_frame(inp[i], fr[i], ref_dil[i], True, 3) for i in range(t)]))


def test_multigpu_halo_blend_on_real_gpus(ops):
    """Rank-boundary halo blend over NCCL and over CUDA-IPC peer reads (tests/mgpu_check.py), when the
    box has at least two GPUs."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "MGPU CHECK PASS" in r.stdout, (r.stdout + r.stderr)[-3000:]


@pytest.mark.parametrize("keep,feather", [(False, 3), (True, 0), (True, 2.5)])
def test_dropin_option_combinations(ops, keep, feather):
    from videovanish_b200 import diffuerase as vvd
    t, h0, w0, h, w = 5, 96, 160, 40, 72
    fr, mk, inp = synth.frames(t, h0, w0, seed=81), synth.masks(t, h0, w0, seed=82), synth.noise_frames(t, h, w, seed=83)
    vvd.set_models(diffueraser=_StubDiffuEraser(list(inp)))
    out = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=3, propainer_frames=list(fr),
                                   keep_unmasked_original=keep, feather_px=feather)
    ref = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp], mask_dilation_iter=3,
                                      propainer_frames=list(fr), keep_unmasked_original=keep, feather_px=feather)
    assert np.array_equal(np.stack(out), np.stack(ref))


@pytest.mark.parametrize("feather", [3, 0, 6.5])
def test_dropin_host_lists_row_bounded_post(ops, feather):
    """Host-list route with page-locked frames and one-object masks: `post` uploads and downloads only the rows the
    resident dilated masks (+ feather radius) reach and copies the rest of each frame from the original on the host -
    same bytes as the reference and as the whole-frame transfers (option pipe_rows = 0)."""
    from videovanish_b200 import _lib, diffuerase as vvd, hostpipe
    t, h0, w0 = 9, 180, 320
    h, w = ops.inference_size(h0, w0, 160)
    fr, mk, inp = synth.frames(t, h0, w0, seed=121), synth.masks(t, h0, w0, seed=122, salt=0.0), synth.noise_frames(t, h, w, seed=123)
    mk[4] = 0                                                     # a frame without a mask: copied entirely on the host
    frames = hostpipe.pinned_frames(t, (h0, w0, 3))
    for i in range(t):
        frames[i][...] = fr[i]
    ref = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp], mask_dilation_iter=4,
                                      propainer_frames=list(fr), max_img_size=160, feather_px=feather)
    vvd.set_models(diffueraser=_StubDiffuEraser(list(inp)))
    assert _lib.get_option("pipe_rows") == 1
    out = vvd.run_infill_on_frames(frames, list(mk), mask_dilation_iter=4, propainer_frames=frames, max_img_size=160, feather_px=feather)
    moved, total = vvd._pipeline.last_rows()
    assert total == t * h0 and 0 < moved < 0.75 * total
    try:
        _lib.set_option("pipe_rows", 0)
        whole = vvd.run_infill_on_frames(frames, list(mk), mask_dilation_iter=4, propainer_frames=frames, max_img_size=160,
                                         feather_px=feather)
        assert vvd._pipeline.last_rows() == (total, total)
    finally:
        _lib.set_option("pipe_rows", 1)
    assert np.array_equal(np.stack(out), np.stack(ref))
    assert np.array_equal(np.stack(whole), np.stack(ref))
    assert np.array_equal(np.stack(frames), fr), "inputs must not be mutated"


def test_dropin_same_size_model_output(ops):
    """Model output already at the original size: the reference skips cv2.resize (:72) but still composites."""
    from videovanish_b200 import diffuerase as vvd
    t, h0, w0 = 4, 72, 128
    fr, mk, inp = synth.frames(t, h0, w0, seed=91), synth.masks(t, h0, w0, seed=92), synth.noise_frames(t, h0, w0, seed=93)
    vvd.set_models(diffueraser=_StubDiffuEraser(list(inp)))
    out = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=2, propainer_frames=list(fr))
    ref = op.ref_run_infill_on_frames(list(fr), list(mk), lambda *a, **k: [x.copy() for x in inp], mask_dilation_iter=2,
                                      propainer_frames=list(fr))
    assert np.array_equal(np.stack(out), np.stack(ref))
    raw = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=2, propainer_frames=list(fr),
                                   keep_unmasked_original=False)
    assert np.array_equal(np.stack(raw), inp)


def test_chunked_driver_blends_overlaps(ops):
    """run_infill_on_frames_chunked == per-chunk oracle outputs stitched by the oracle's blend (row A11)."""
    from videovanish_b200 import chunking
    from videovanish_b200 import diffuerase as vvd
    t, h0, w0, h, w, chunk, ov = 23, 72, 128, 32, 64, 10, 4
    fr, mk = synth.frames(t, h0, w0, seed=101), synth.masks(t, h0, w0, seed=102)
    plan = chunking.chunk_plan(t, chunk, ov)

    class _PerChunkModel:                       # a model whose output depends on the chunk it is called with
        def __init__(self):
            self.calls = 0

        def forward(self, frames, masks, priors, **kw):
            self.calls += 1
            return list(synth.noise_frames(len(frames), h, w, seed=200 + self.calls))

    model = _PerChunkModel()
    vvd.set_models(diffueraser=model)
    got = vvd.run_infill_on_frames_chunked(list(fr), list(mk), chunk=chunk, overlap=ov, mask_dilation_iter=2,
                                           propainer_frames=list(fr))
    assert model.calls == len(plan) and len(got) == t
    per_chunk = []
    for ci, (s, e) in enumerate(plan):
        inp = list(synth.noise_frames(e - s, h, w, seed=201 + ci))
        per_chunk.append(np.stack(op.ref_run_infill_on_frames(list(fr[s:e]), list(mk[s:e]), lambda *a, **k: inp,
                                                              mask_dilation_iter=2, propainer_frames=list(fr[s:e]))))
    assert np.array_equal(np.stack(got), ocb.stitch_chunks(per_chunk, plan, ov))


def test_tools_writer_fits_frames_on_the_gpu(ops, tmp_path):
    """tools.write_video_frames_to_path brings frames of another size to (H0, W0) with NEAREST (tools.py:41-42):
    same bytes as the reference's cv2.resize, read back from the lossless file."""
    cv2 = pytest.importorskip("cv2")
    from videovanish_b200 import tools
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (30, 44, 3), dtype=np.uint8), rng.integers(0, 256, (48, 64, 3), dtype=np.uint8),
              rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)]
    path = str(tmp_path / "fit.mkv")
    try:
        tools.write_video_frames_to_path(path, frames, 25.0, 48, 64)
    except AssertionError:
        pytest.skip("FFV1 writer unavailable in this OpenCV build")
    got, _ = tools.load_video_frames_from_path(path)
    want = [f if f.shape[:2] == (48, 64) else cv2.resize(f, (64, 48), interpolation=cv2.INTER_NEAREST) for f in frames]
    assert len(got) == 3 and all(np.array_equal(g, w) for g, w in zip(got, want))

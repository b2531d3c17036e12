"""GPU parity tests added in round 2 (run with ``-m gpu``): the k3_fast kernel and its 1-bit mask input, the
production 960x536 and 4K geometries, the N1 / N2 / N4 glue kernels, the device-resident route of the drop-in
through the wrapper adapters, and the pads-free K4 output.  Every call goes through the C ABI."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import full_path as ofp
from oracle import prepost as op
from oracle import propagation as opp
from oracle import wrapper as ow
from videovanish_b200 import synth

pytestmark = pytest.mark.gpu


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


def border_masks(t, h0, w0, seed, n=2):
    dil = np.stack(op.model_binarize_dilate(list(synth.masks(t, h0, w0, seed=seed, salt=0.004)), n))
    if t > 1:
        dil[1] = 255                                          # everything inside: every quad, all borders
    if t > 2:
        dil[2, :3] = dil[2, -2:] = 255
        dil[2, :, :5] = dil[2, :, -3:] = 255
    return dil


# ------------------------------------------------------------------------------- K1 bit plane
@pytest.mark.parametrize("h,w,n", [(64, 96, 3), (37, 50, 8), (120, 176, 0), (33, 64, 40)])
def test_k1_returns_its_bit_plane(ops, h, w, n):
    mk = synth.masks(3, h, w, seed=h + n, salt=0.002)
    mk[1] = 0
    out, low, bits = ops.binarize_dilate(dev(mk), n, return_bits=True)
    assert low is None
    ref = np.stack(op.ref_binarize_dilate(list(mk), n))
    assert np.array_equal(host(out), ref)
    wp = (w + 31) // 32
    padded = np.zeros((3, h, wp * 32), np.uint8)
    padded[:, :, :w] = ref > 0
    want = np.packbits(padded.reshape(3, h, wp, 32), axis=-1, bitorder="little").view(np.uint32).reshape(3, h, wp)
    assert np.array_equal(host(bits).view(np.uint32), want)
    out2, low2, bits2 = ops.binarize_dilate(dev(mk), n, lowres_size=(h // 2, w // 2), return_bits=True)
    assert np.array_equal(host(out2), ref) and np.array_equal(host(bits2).view(np.uint32), want)
    assert np.array_equal(host(low2), np.stack([op.ref_resize_nearest(m, h // 2, w // 2) for m in ref]))


# ------------------------------------------------------------------------------- K3 composited in parts
@pytest.mark.parametrize("h0,w0,h,w", [(120, 176, 60, 88), (120, 176, 56, 88)])
def test_k3_in_chained_parts_equals_one_call(ops, h0, w0, h, w):
    """A clip composited in three back-to-back K3 calls (boundary frames first, the later calls chained to the one
    before: programmatic stream serialisation, no drain bubble) gives the bytes of one call over the clip."""
    t = 12
    fr, inp = synth.frames(t, h0, w0, seed=71), synth.noise_frames(t, h, w, seed=72)
    dil, _, bits = ops.binarize_dilate(dev(synth.masks(t, h0, w0, seed=73, salt=0.004)), 3, return_bits=True)
    d_inp, d_fr = dev(inp), dev(fr)
    whole = ops.upscale_feather_composite(d_inp, d_fr, dil, 3, mask_bits=bits)
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], host(dil)[i], True, 3) for i in range(t)])
    assert np.array_equal(host(whole), ref)
    for rep in range(3):
        parts = torch.zeros_like(whole)
        for n, (lo, hi) in enumerate([(0, 3), (9, 12), (3, 9)]):
            ops.upscale_feather_composite(d_inp[lo:hi], d_fr[lo:hi], dil[lo:hi], 3, out=parts[lo:hi], mask_bits=bits[lo:hi],
                                          chain_previous=n > 0)
        assert torch.equal(parts, whole)


# ------------------------------------------------------------------------------- k3_fast
@pytest.mark.parametrize("h0,w0,h,w", [(120, 176, 60, 88), (120, 176, 56, 88), (16, 16, 8, 8), (66, 80, 33, 40),
                                       (72, 128, 32, 64), (100, 96, 48, 48), (30, 48, 16, 24),
                                       (72, 128, 20, 32), (40, 64, 10, 16), (135, 240, 33, 60), (64, 16, 16, 4)])
@pytest.mark.parametrize("f", [3, 2.5, 1.5, 0.5])
def test_k3_fast_kernel(ops, h0, w0, h, w, f):
    """W0 == 2w or 4w takes k3_fast (closed-form horizontal pass; vertical closed form when H0 == 2h, table driven
    otherwise - the 1080p <- 960x536 and 4K <- 960x536 production cases): same bytes as the oracle, as the older
    kernels, with and without the 1-bit mask plane, for every rows-per-task setting."""
    from videovanish_b200 import _lib
    t = 3
    fr = synth.frames(t, h0, w0, seed=61)
    inp = synth.noise_frames(t, h, w, seed=62)
    dil = border_masks(t, h0, w0, 63)
    ref = np.stack([op.ref_post_frame(inp[i], fr[i], dil[i], True, f) for i in range(t)])
    assert _lib.get_option("k3_x2") == 3                     # default: k3_fastw (word tasks)
    d_inp, d_fr, d_dil = dev(inp), dev(fr), dev(dil)
    wp = (w0 + 31) // 32
    padded = np.zeros((t, h0, wp * 32), np.uint8)
    padded[:, :, :w0] = dil > 0
    bits = dev(np.packbits(padded.reshape(t, h0, wp, 32), axis=-1, bitorder="little").view(np.int32).reshape(t, h0, wp))
    got = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f))
    got_bits = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
    variants = {}
    try:
        for thr, rows in ((256, 16), (256, 6), (512, 5), (384, 16), (384, 3)):
            _lib.set_option("k3_tma_threads", thr)
            _lib.set_option("k3_tma_rows", rows)
            variants["w-threads%d-rows%d" % (thr, rows)] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        _lib.set_option("k3_tma_threads", 512)
        _lib.set_option("k3_tma_rows", 16)
        _lib.set_option("k3_x2", 2)                          # k3_fast: 16-pixel rolling tasks
        variants["rolling"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f))
        variants["rolling-bits"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        for rpt in (3, 4, 8, 16):
            _lib.set_option("k3_nt", rpt)
            variants["rpt%d" % rpt] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        _lib.set_option("k3_nt", 2)
        for thr, rows in ((256, 16), (256, 6), (512, 5)):
            _lib.set_option("k3_tma_threads", thr)
            _lib.set_option("k3_tma_rows", rows)
            variants["threads%d-rows%d" % (thr, rows)] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        _lib.set_option("k3_tma_threads", 512)
        _lib.set_option("k3_tma_rows", 16)
        _lib.set_option("k3_bits", 0)
        variants["rolling-bits-ignored"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        _lib.set_option("k3_x2", 3)
        variants["bits-ignored"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f, mask_bits=bits))
        _lib.set_option("k3_bits", 1)
        _lib.set_option("k3_x2", 1)
        variants["old-x2"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f))
        _lib.set_option("k3_x2", 0)
        variants["generic"] = host(ops.upscale_feather_composite(d_inp, d_fr, d_dil, feather_px=f))
    finally:
        _lib.set_option("k3_nt", 2)
        _lib.set_option("k3_bits", 1)
        _lib.set_option("k3_x2", 3)
        _lib.set_option("k3_tma_threads", 512)
        _lib.set_option("k3_tma_rows", 16)
    assert np.array_equal(got, ref)
    assert np.array_equal(got_bits, ref)
    for name, v in variants.items():
        assert np.array_equal(v, ref), name


def test_k3_production_geometry_960x536(ops):
    """1080p originals, model output at 960x536 (what the wrapper's multiple-of-8 rule really produces,
    SURVEY row A9): k3_fast with table-driven vertical taps, bit-exact on whole frames."""
    t, h0, w0 = 2, 1080, 1920
    h, w = ops.inference_size(h0, w0, 960)
    assert (h, w) == (536, 960) == op.inference_size(h0, w0, 960)
    fr, inp = synth.frames(t, h0, w0, seed=71), synth.noise_frames(t, h, w, seed=72)
    mk = synth.masks(t, h0, w0, seed=73)
    dil, low, bits = ops.binarize_dilate(dev(mk), 8, lowres_size=(h, w), return_bits=True)
    ref_dil = op.ref_binarize_dilate(list(mk), 8)
    assert np.array_equal(host(dil), np.stack(ref_dil))
    assert np.array_equal(host(low), np.stack([op.ref_resize_nearest(m, h, w) for m in ref_dil]))
    out = host(ops.upscale_feather_composite(dev(inp), dev(fr), dil, 3, mask_bits=bits))
    for i in range(t):
        assert np.array_equal(out[i], op.ref_post_frame(inp[i], fr[i], ref_dil[i], True, 3))
    small = host(ops.resize(dev(fr), h, w))
    assert np.array_equal(small[0], op.ref_resize_linear(fr[0], h, w))


def test_config4_one_full_4k_frame(ops):
    """BASELINE config 4 geometry: one 2160x3840 frame through K1 (+ low-res mask), K2, K3 and K5 against the
    oracle (inference size 536x960)."""
    from oracle import chunk_blend as ocb
    h0, w0 = 2160, 3840
    h, w = ops.inference_size(h0, w0, 960)
    assert (h, w) == (536, 960)
    fr, inp = synth.frames(2, h0, w0, seed=81), synth.noise_frames(1, h, w, seed=82)
    mk = synth.masks(1, h0, w0, seed=83, salt=0.0002)
    dil, low, bits = ops.binarize_dilate(dev(mk), 8, lowres_size=(h, w), return_bits=True)
    ref_dil = op.ref_binarize_dilate(list(mk), 8)[0]
    assert np.array_equal(host(dil)[0], ref_dil)
    assert np.array_equal(host(low)[0], op.ref_resize_nearest(ref_dil, h, w))
    assert np.array_equal(host(ops.resize(dev(fr[:1]), h, w))[0], op.ref_resize_linear(fr[0], h, w))
    out = ops.upscale_feather_composite(dev(inp), dev(fr[:1]), dil, 3, mask_bits=bits)
    ref_out = op.ref_post_frame(inp[0], fr[0], ref_dil, True, 3)
    assert np.array_equal(host(out)[0], ref_out)
    bl = host(ops.chunk_blend(out, dev(fr[1:]), k0=3, overlap_total=16))
    assert np.array_equal(bl, ocb.blend_overlap(np.stack([ref_out] * 16), np.stack([fr[1]] * 16))[3:4])


# ------------------------------------------------------------------------------- N1 / N2 / N4 glue
@pytest.mark.parametrize("shape", [(3, 36, 64, 3), (1, 7, 5, 3), (2, 33, 47, 3)])
def test_n1_channel_swap(ops, shape):
    import cv2
    a = np.random.default_rng(shape[1]).integers(0, 256, shape, dtype=np.uint8)
    ref = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2RGB) for f in a])
    assert np.array_equal(host(ops.swap_rb(dev(a))), ref)
    d = dev(a)
    assert ops.swap_rb(d, out=d) is d and np.array_equal(host(d), ref)        # in place
    off = dev(np.concatenate([np.zeros(3, np.uint8), a.reshape(-1)]))[3:].view(shape)   # unaligned base
    assert np.array_equal(host(ops.swap_rb(off)), ref)


@pytest.mark.parametrize("t,h,w,nl", [(23, 36, 64, 10), (12, 17, 13, 4), (5, 8, 8, 10), (1, 16, 24, 10)])
def test_n2_neighbor_merge(ops, t, h, w, nl):
    rng = np.random.default_rng(t * h)
    plan = opp.neighbor_plan(t, nl, 10, 50)
    preds = [rng.uniform(-1, 1, (len(ids), 3, h, w)).astype(np.float32) for ids, _ in plan]
    preds[0][0, :, 0, :3] = [[-1.0, 1.0, 0.0]] * 3                                   # exact end points and the mid level
    m = (rng.random((t, h, w)) < 0.5).astype(np.uint8)
    ori = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    want = np.stack(opp.ref_neighbor_merge(preds, plan, m, ori))
    comp = torch.zeros((t, h, w, 3), dtype=torch.uint8, device="cuda")
    d_m, d_ori = dev(m * 255), dev(ori)
    seen = [False] * t
    for (ids, _), p in zip(plan, preds):
        a, b = ids[0], ids[-1] + 1
        ops.neighbor_merge(dev(p), d_m[a:b], d_ori[a:b], comp[a:b], [not seen[i] for i in ids])
        for i in ids:
            seen[i] = True
    assert np.array_equal(host(comp), want)


def test_n4_masked_frames(ops):
    rng = np.random.default_rng(9)
    for (t, h, w) in ((3, 36, 64), (2, 17, 13), (1, 5, 3)):
        fr = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
        m = ((rng.random((t, h, w)) < 0.4) * rng.integers(1, 256, (t, h, w))).astype(np.uint8)
        want = np.stack([ow.ref_masked_frame(fr[i], m[i]) for i in range(t)])
        assert np.array_equal(host(ops.apply_mask(dev(fr), dev(m))), want)


# ------------------------------------------------------------------------------- device-resident drop-in
def _flow_fn_np(seed):
    return lambda small, low: synth.flows(len(small), small.shape[1], small.shape[2], seed=seed)


def _install_adapters(seed):
    from videovanish_b200 import diffuerase as vvd, wrappers
    fn = _flow_fn_np(seed)

    def flow_fn(small, low):
        ff, fb = fn(small.cpu().numpy(), low.cpu().numpy())
        return dev(ff), dev(fb)

    prior = wrappers.ProPainterPrior(flow_fn=flow_fn, network_fn=lambda upd, um, low, ids, refs: upd[ids[0]:ids[-1] + 1],
                                     max_img_size=vvd_max[0])
    eraser = wrappers.DiffuEraserWrapper(network_fn=lambda masked, m, priors: priors)
    vvd.set_models(diffueraser=eraser, propainter_model=prior)
    return vvd


vvd_max = [160]


@pytest.mark.parametrize("t,h0,w0,size,dilate,feather,pinned", [(23, 180, 320, 160, 5, 3, False), (7, 72, 128, 64, 3, 3, True),
                                                                (61, 90, 160, 80, 2, 2.5, False), (1, 64, 96, 48, 8, 3, False)])
def test_device_resident_dropin_matches_full_path_oracle(ops, t, h0, w0, size, dilate, feather, pinned):
    """run_infill_on_frames with the wrapper adapters installed: frames and masks cross PCIe once, K1 -> K2 -> K4
    -> N2 -> N4 -> K3 run in HBM, and the result equals the composition of the stage oracles (networks stubbed
    identically on both sides)."""
    from videovanish_b200 import hostpipe
    vvd_max[0] = size
    vvd = _install_adapters(seed=t)
    fr, mk = synth.frames(t, h0, w0, seed=t + 1), synth.masks(t, h0, w0, seed=t + 2, salt=0.0008)
    frames, masks = list(fr), list(mk)
    if pinned:
        frames, masks = hostpipe.pinned_frames(t, (h0, w0, 3)), hostpipe.pinned_frames(t, (h0, w0, 3))
        for i in range(t):
            frames[i][...] = fr[i]
            masks[i][...] = mk[i]
    calls = []
    try:
        out = vvd.run_infill_on_frames(frames, masks, mask_dilation_iter=dilate, max_img_size=size, feather_px=feather,
                                       prog=lambda p, s: calls.append(p))
    finally:
        vvd.propainter = None
    want = ofp.run(list(fr), list(mk), _flow_fn_np(t), mask_dilation_iter=dilate, max_img_size=size, feather_px=feather)
    assert isinstance(out, list) and len(out) == t
    assert all(o.dtype == np.uint8 and o.flags.c_contiguous and o.shape == (h0, w0, 3) for o in out)
    assert np.array_equal(np.stack(out), np.stack(want))
    assert calls == [5, 10, 20, 50, 90]
    assert np.array_equal(np.stack(frames), fr) and np.array_equal(np.stack(masks), mk), "inputs must not be mutated"


def test_dropin_masks_of_another_size(ops):
    """Masks that do not have the frames' size (diffuerase.py:85-86): K1 runs at the masks' own size - that is what the
    models are handed, like in the reference - and the post stage fits them with INTER_NEAREST (K2).  Host-list route
    against the golden vector of the unmodified reference, device-resident route against the full-path oracle."""
    from tests.test_oracle import other_mask_size_case
    from videovanish_b200 import diffuerase as vvd

    class Stub:
        def forward(self, frames, masks, priors, **kw):
            self.masks = [m.copy() for m in masks]
            return [x.copy() for x in self.inpainted]

    z, fr, mk, inp, n, f = other_mask_size_case()
    stub = Stub()
    stub.inpainted = list(inp)
    vvd.set_models(diffueraser=stub)
    out = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=n, propainer_frames=list(fr), feather_px=f)
    assert np.array_equal(np.stack(stub.masks), z["dilated"]), "the models get the masks dilated at their own size"
    assert np.array_equal(np.stack(out), z["out"])
    # device-resident route: same geometry, longer clip
    t, (h0, w0), (hm, wm) = 9, fr.shape[1:3], mk.shape[1:3]
    fr2, mk2 = synth.frames(t, h0, w0, seed=5), synth.masks(t, hm, wm, seed=6, salt=0.001)
    vvd_max[0] = 80
    vvd = _install_adapters(seed=t)
    try:
        got = vvd.run_infill_on_frames(list(fr2), list(mk2), mask_dilation_iter=n, max_img_size=80, feather_px=f)
    finally:
        vvd.propainter = None
    want = ofp.run(list(fr2), list(mk2), _flow_fn_np(t), mask_dilation_iter=n, max_img_size=80, feather_px=f)
    assert np.array_equal(np.stack(got), np.stack(want))


def test_mask_row_bounds_kernel(ops):
    t, h, w = 6, 75, 200
    dil = np.zeros((t, h, w), np.uint8)
    dil[0, 10:20, 5:9] = 255
    dil[1, 0, 199] = 255                      # first row, last column
    dil[2, 74, 0] = 255                       # last row
    dil[3] = 255
    dil[5, 30, 64] = dil[5, 50, 3] = 255      # frame 4 stays empty
    wp = (w + 31) // 32
    padded = np.zeros((t, h, wp * 32), np.uint8)
    padded[:, :, :w] = dil > 0
    bits = dev(np.packbits(padded.reshape(t, h, wp, 32), axis=-1, bitorder="little").view(np.int32).reshape(t, h, wp))
    for margin in (0, 3):
        got = host(ops.mask_row_bounds(bits, margin))
        for i in range(t):
            ys = np.nonzero(dil[i].any(axis=1))[0]
            want = (0, 0) if len(ys) == 0 else (max(0, ys[0] - margin), min(h, ys[-1] + 1 + margin))
            assert tuple(got[i]) == want, (i, margin)


@pytest.mark.parametrize("t,h0,w0,size,dilate,feather", [(12, 180, 320, 160, 5, 3), (9, 90, 160, 80, 2, 6.5), (5, 72, 128, 64, 3, 0)])
def test_device_resident_row_bounded_results(ops, t, h0, w0, size, dilate, feather):
    """One-object masks: only the rows the dilated mask (+ feather radius) reaches are downloaded, the rest of every
    finished frame is copied from the caller's input frame on the host - same bytes as the full download and as the
    composition of the stage oracles."""
    vvd_max[0] = size
    vvd = _install_adapters(seed=t)
    fr, mk = synth.frames(t, h0, w0, seed=t + 1), synth.masks(t, h0, w0, seed=t + 2, salt=0.0)
    mk[t // 2] = 0                                               # a frame without any mask: copied entirely on the host
    want = ofp.run(list(fr), list(mk), _flow_fn_np(t), mask_dilation_iter=dilate, max_img_size=size, feather_px=feather)
    try:
        assert vvd.ROW_BOUNDED_RESULTS
        out = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=dilate, max_img_size=size, feather_px=feather)
        info = dict(vvd.last_call_info)
        vvd.ROW_BOUNDED_RESULTS = False
        full = vvd.run_infill_on_frames(list(fr), list(mk), mask_dilation_iter=dilate, max_img_size=size, feather_px=feather)
        info_full = dict(vvd.last_call_info)
    finally:
        vvd.ROW_BOUNDED_RESULTS = True
        vvd.propainter = None
    assert info["row_bounded"] and 0 < info["rows_downloaded"] < 0.75 * t * h0
    assert not info_full["row_bounded"] and info_full["rows_downloaded"] == t * h0
    assert all(o.dtype == np.uint8 and o.flags.c_contiguous and o.shape == (h0, w0, 3) for o in out)
    assert np.array_equal(np.stack(out), np.stack(want))
    assert np.array_equal(np.stack(full), np.stack(want))


def test_wrapper_adapters_keep_the_upstream_host_signatures(ops):
    """The adapters also work as plain host-list models behind the reference's calls (diffuerase.py:52-57, :62-67)."""
    from videovanish_b200 import wrappers
    t, h0, w0, size = 9, 72, 128, 64
    fr, mk = synth.frames(t, h0, w0, seed=3), synth.masks(t, h0, w0, seed=4, salt=0.0005)
    dil = op.ref_binarize_dilate(list(mk), 3)
    fn = _flow_fn_np(5)
    prior = wrappers.ProPainterPrior(flow_fn=lambda s, l: tuple(dev(x) for x in fn(s.cpu().numpy(), l.cpu().numpy())),
                                     network_fn=lambda upd, um, low, ids, refs: upd[ids[0]:ids[-1] + 1], max_img_size=size)
    pri = prior.forward(list(fr), dil, ref_stride=10, neighbor_length=10, subvideo_length=50, mask_dilation=0)
    h, w = op.inference_size(h0, w0, size)
    small = np.stack([op.ref_resize_linear(f, h, w) for f in fr])
    low = np.stack([op.ref_resize_nearest(m, h, w) for m in dil])
    ref_pri = ofp.propainter_prior(small, low, *fn(small, low))
    assert np.array_equal(np.stack(pri), np.stack([op.ref_resize_linear(p, h0, w0) for p in ref_pri]))
    eraser = wrappers.DiffuEraserWrapper(network_fn=lambda masked, m, priors: priors)
    got = eraser.forward(list(fr), dil, pri, max_img_size=size, mask_dilation_iter=0, guidance_scale=None)
    pri_small = [op.ref_resize_linear(p, h, w) for p in pri]
    m = [ow.ref_wrapper_mask(x, 0) for x in low]
    want = [ow.ref_wrapper_compose(pri_small[i], small[i], m[i], True) for i in range(t)]
    assert np.array_equal(np.stack(got), np.stack(want))


def test_non_uint8_masks_are_binarised_before_the_cast(ops):
    """A float mask in (0,1) or an int16 value of 256 is 'set' for the reference (m > 0 on the original dtype,
    diffuerase.py:29); a plain uint8 cast would drop it."""
    from videovanish_b200 import hostpipe
    h0, w0 = 40, 64
    m_f = np.zeros((h0, w0, 3), np.float32)
    m_f[10:20, 10:30, 1] = 0.25
    m_i = np.zeros((h0, w0, 3), np.int16)
    m_i[5:9, 40:50, 2] = 256
    pipe = hostpipe.HostPipeline(h0, w0)
    dil = pipe.pre([m_f, m_i], 2)
    pipe.close()
    assert np.array_equal(np.stack(dil), np.stack(op.ref_binarize_dilate([m_f, m_i], 2)))


# ------------------------------------------------------------------------------- N1: device-resident frame I/O
def test_n1_device_resident_frame_io(ops, tmp_path):
    """tools.load_video_frames_from_path(device='cuda') == the host loader (decode -> pinned ring -> H2D -> swap on
    the device), the writer takes the device clip (NEAREST fix-up + swap on the device, double-buffered download),
    and run_infill_on_frames takes DeviceFrames in and gives DeviceFrames out without touching the host."""
    pytest.importorskip("cv2")
    from videovanish_b200 import tools
    rng = np.random.default_rng(12)
    t, h0, w0 = 70, 72, 128                          # more than two ring blocks of 32 frames
    frames = list(rng.integers(0, 256, (t, h0, w0, 3), dtype=np.uint8))
    path, path2, path3 = (str(tmp_path / n) for n in ("a.mkv", "b.mkv", "c.mkv"))
    try:
        tools.write_video_frames_to_path(path, frames, 25.0, h0, w0)
    except AssertionError:
        pytest.skip("FFV1 writer unavailable in this OpenCV build")
    host_frames, fps = tools.load_video_frames_from_path(path, start_frame=3, max_frames=66)
    clip, fps2 = tools.load_video_frames_from_path(path, start_frame=3, max_frames=66, device="cuda")
    assert fps == fps2 and len(clip) == len(host_frames) == 66
    assert np.array_equal(host(clip.tensor), np.stack(host_frames)) and np.array_equal(np.stack(host_frames), np.stack(frames[3:69]))
    assert np.array_equal(clip[5], host_frames[5])                              # list behaviour (lazy download)
    tools.write_video_frames_to_path(path2, clip, fps, h0, w0)
    back, _ = tools.load_video_frames_from_path(path2)
    assert np.array_equal(np.stack(back), np.stack(host_frames))
    tools.write_video_frames_to_path(path3, clip, fps, 36, 64)                   # the writer's NEAREST fix-up (tools.py:41-42)
    small, _ = tools.load_video_frames_from_path(path3)
    assert np.array_equal(np.stack(small), np.stack([op.ref_resize_nearest(f, 36, 64) for f in host_frames]))

    # device clip in -> device clip out through the drop-in
    vvd_max[0] = 64
    vvd = _install_adapters(seed=4)
    mk = synth.masks(66, h0, w0, seed=8, salt=0.0008)
    try:
        out = vvd.run_infill_on_frames(clip, tools.DeviceFrames(dev(mk)), mask_dilation_iter=3, max_img_size=64)
    finally:
        vvd.propainter = None
    assert isinstance(out, tools.DeviceFrames)
    want = ofp.run(host_frames, list(mk), _flow_fn_np(4), mask_dilation_iter=3, max_img_size=64)
    assert np.array_equal(host(out.tensor), np.stack(want))


def test_loader_pinned_budget(ops, tmp_path, monkeypatch):
    """Beyond tools.PINNED_BUDGET the loader hands out ordinary host memory (no unbounded pinning)."""
    pytest.importorskip("cv2")
    from videovanish_b200 import tools
    frames = list(np.random.default_rng(1).integers(0, 256, (40, 32, 48, 3), dtype=np.uint8))
    path = str(tmp_path / "p.mkv")
    try:
        tools.write_video_frames_to_path(path, frames, 25.0, 32, 48)
    except AssertionError:
        pytest.skip("FFV1 writer unavailable in this OpenCV build")
    monkeypatch.setattr(tools, "PINNED_BUDGET", 32 * 32 * 48 * 3)               # exactly one block
    got, _ = tools.load_video_frames_from_path(path)
    assert np.array_equal(np.stack(got), np.stack(frames))
    pinned = [torch.from_numpy(g).is_pinned() for g in got]
    assert all(pinned[:32]) and not any(pinned[32:])


def test_chunked_driver_device_resident(ops):
    """run_infill_on_frames_chunked with the adapters installed: every chunk runs device resident, the overlaps are
    cross-faded in HBM, chunk c+1 uploads while chunk c downloads - and the clip equals the per-chunk full-path
    oracle outputs stitched by the oracle's blend (row A11)."""
    from oracle import chunk_blend as ocb
    from videovanish_b200 import chunking
    t, h0, w0, size, chunk, ov = 47, 72, 128, 64, 20, 6
    vvd_max[0] = size
    vvd = _install_adapters(seed=9)
    fr, mk = synth.frames(t, h0, w0, seed=111), synth.masks(t, h0, w0, seed=112, salt=0.0008)
    try:
        got = vvd.run_infill_on_frames_chunked(list(fr), list(mk), chunk=chunk, overlap=ov, mask_dilation_iter=3,
                                               max_img_size=size)
    finally:
        vvd.propainter = None
    plan = chunking.chunk_plan(t, chunk, ov)
    per_chunk = [np.stack(ofp.run(list(fr[s:e]), list(mk[s:e]), _flow_fn_np(9), mask_dilation_iter=3, max_img_size=size))
                 for s, e in plan]
    assert len(got) == t and np.array_equal(np.stack(got), ocb.stitch_chunks(per_chunk, plan, ov))

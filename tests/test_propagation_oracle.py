"""Row A10 oracle self-consistency (CPU): the explicit numpy model on the packed u8 state
== the torch restatement that uses torch's own CPU F.grid_sample.  PARITY UNPINNED by the
reference (no upstream code / vectors in /root/reference); see oracle/propagation.py."""
import numpy as np
import pytest

from oracle import propagation as opp
from videovanish_b200 import synth

torch = pytest.importorskip("torch")

# torch's CPU grid_sample has ISA-specific code paths (FMA contraction); on the build
# container the match is exact.  Allow a vanishing fraction elsewhere, expect 0.
MAX_MISMATCH_FRACTION = 1e-5


def clip(t, h, w, seed, big_motion=False):
    fr = synth.frames(t, h, w, seed=seed)
    mk = synth.masks(t, h, w, seed=seed + 1, salt=0.002)
    m = (mk.max(axis=3) > 0).astype(np.uint8) * 255
    ff, fb = synth.flows(t, h, w, seed=seed + 2)
    if big_motion:                         # push samples out of the frame to exercise zero padding
        ff[..., 0] += 9.0
        fb[..., 0] -= 9.0
    return fr, m, ff, fb


@pytest.mark.parametrize("t,h,w,big", [(6, 48, 64, False), (5, 40, 56, True), (12, 72, 96, False), (2, 33, 47, False)])
def test_model_matches_torch(t, h, w, big):
    fr, m, ff, fb = clip(t, h, w, seed=t * 100 + h, big_motion=big)
    ref_frames, ref_masks = opp.img_propagation_torch(fr, m, ff, fb)
    got_frames, got_masks = opp.decode_state(opp.model_propagate(fr, m, ff, fb))
    assert (got_masks != ref_masks).mean() <= MAX_MISMATCH_FRACTION
    assert (got_frames != ref_frames).mean() <= MAX_MISMATCH_FRACTION
    # the scan does something: holes shrink, never grow
    assert got_masks.sum() < (m > 0).sum()
    assert np.all(got_masks <= (m > 0))


def test_single_frame_and_no_holes():
    fr, m, ff, fb = clip(1, 24, 32, seed=5)
    p = opp.model_propagate(fr, m, ff, fb)
    assert np.array_equal(p, opp.pack_state(fr, m))
    fr, m, ff, fb = clip(4, 24, 32, seed=6)
    m[:] = 0
    p = opp.model_propagate(fr, m, ff, fb)
    assert np.array_equal(p, opp.pack_state(fr, m))


def test_subvideo_plan_matches_upstream_loop():
    assert opp.subvideo_plan(40) == [(0, 40, 0, 0)]
    assert opp.subvideo_plan(120) == [(0, 60, 0, 10), (40, 110, 10, 10), (90, 120, 10, 0)]
    plan = opp.subvideo_plan(300)
    assert len(plan) == 6 and plan[0] == (0, 60, 0, 10) and plan[-1] == (240, 300, 10, 0)
    kept = sum(e - s - ps - pe for s, e, ps, pe in plan)
    assert kept == 300


def test_clip_level_model_matches_torch():
    fr, m, ff, fb = clip(14, 32, 40, seed=9)
    u, um = opp.propagate_clip_torch(fr, m, ff, fb, subvideo_length=5, pad_len=2)
    g, gm = opp.decode_state(opp.model_propagate_clip(fr, m, ff, fb, subvideo_length=5, pad_len=2))
    assert u.shape == g.shape == (14, 3, 32, 40)
    assert (g != u).mean() <= MAX_MISMATCH_FRACTION and (gm != um).mean() <= MAX_MISMATCH_FRACTION


def test_zero_padding_fill_state():
    """A hole on the left border whose small outward flow samples the zero padding is filled
    with the float 0.0 (state ZERO without HOLE), exactly as torch's grid_sample does."""
    t, h, w = 3, 24, 32
    fr = synth.frames(t, h, w, seed=5)
    m = np.zeros((t, h, w), np.uint8)
    m[1, 4:20, 0:2] = 255
    ff = np.zeros((t - 1, h, w, 2), np.float32)
    fb = np.zeros((t - 1, h, w, 2), np.float32)
    ff[..., 0] = -0.61
    fb[..., 0] = 0.58
    p = opp.model_propagate(fr, m, ff, fb)
    assert ((p >> 24) == opp.ZERO).sum() == 16 and ((p >> 24) & opp.HOLE).sum() == 0
    rf, rm = opp.img_propagation_torch(fr, m, ff, fb)
    gf, gm = opp.decode_state(p)
    assert np.array_equal(gf, rf) and np.array_equal(gm, rm)
    assert (rf[1, :, 4:20, 0] == 0).all()

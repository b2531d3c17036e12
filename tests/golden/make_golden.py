"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/diffuerase.py through oracle/reference_harness.py) on small seeded
clips.  Run in the build container only:  python tests/golden/make_golden.py

Every file stores the inputs' generator arguments (so tests regenerate the same
inputs from videovanish_b200.synth) plus the reference's outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh          # noqa: E402
from videovanish_b200 import synth                  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, T, H0, W0, h, w, dilation N, feather F, keep_unmasked)
CASES = [
    ("c1_small", 2, 360, 640, 176, 320, 8, 3, True),            # BASELINE config 1 shape, first 2 frames
    ("odd_sizes", 3, 97, 131, 40, 56, 3, 3, True),              # ragged: W not a multiple of 16/4
    ("dil1_f5", 2, 120, 160, 56, 80, 1, 5, True),
    ("dil25_f2", 2, 144, 256, 72, 128, 25, 2, True),            # GUI maximum dilation
    ("dil0_fill", 2, 64, 96, 32, 48, 0, 3, True),               # iterations=0 -> until convergence (KAT T2)
    ("hard_alpha", 2, 90, 160, 40, 80, 4, 0, True),             # feather_px <= 0 -> hard composite
    ("no_keep", 2, 90, 160, 40, 80, 4, 3, False),               # keep_unmasked_original=False: resize only
    ("same_size", 2, 72, 128, 72, 128, 2, 3, True),             # no resize-back branch
    ("frac_feather", 2, 80, 112, 40, 56, 2, 2.5, True),         # float feather_px
    ("hd_crop", 1, 270, 480, 136, 240, 8, 3, True),
    ("feather12_5", 2, 120, 176, 56, 88, 3, 12.5, True),        # beyond 8: the two-pass chamfer table is not symmetric any more
    ("feather32", 2, 150, 200, 72, 96, 5, 32, True),            # VV_MAX_FEATHER
]


def inputs(t, h0, w0, h, w, seed):
    fr = synth.frames(t, h0, w0, seed=seed)
    mk = synth.masks(t, h0, w0, seed=seed + 1, salt=0.002)
    inp = synth.noise_frames(t, h, w, seed=seed + 2)
    return fr, mk, inp


def main():
    assert rh.available(), "needs /root/reference"
    for ci, (name, t, h0, w0, h, w, n, f, keep) in enumerate(CASES):
        seed = 100 + 10 * ci
        fr, mk, inp = inputs(t, h0, w0, h, w, seed)
        if name == "dil0_fill":
            mk[1] = 0                                             # an empty frame must stay empty
        outs, dil = rh.ref_post_all_frames(list(fr), list(mk), list(inp), mask_dilation_iter=n,
                                           keep_unmasked_original=keep, feather_px=f)
        lit, _, kw = rh.ref_run_literal(list(fr), list(mk), list(inp), mask_dilation_iter=n,
                                        keep_unmasked_original=keep, feather_px=f)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            args=np.array([t, h0, w0, h, w, n, seed], np.int64), feather=np.float64(f), keep=np.bool_(keep),
            empty_frame1=np.bool_(name == "dil0_fill"),
            dilated=np.stack(dil), out=np.stack(outs), literal_frame0=lit[0],
            literal_rest_raw=np.bool_(all(np.array_equal(a, b) for a, b in zip(lit[1:], inp[1:]))))
        print(name, "dilated px", int(np.stack(dil).astype(bool).sum()), "out", np.stack(outs).shape)


def main_mask_size():
    """tests/golden/other_mask_size.npz: masks of ANOTHER size than the frames - the reference dilates them at their own
    size and fits them with INTER_NEAREST in the post loop (diffuerase.py:85-86).  Same reference calls as `main`."""
    t, h0, w0, h, w, hm, wm, n, f, seed = 3, 90, 160, 40, 80, 67, 101, 3, 3, 300
    fr, _, inp = inputs(t, h0, w0, h, w, seed)
    mk = synth.masks(t, hm, wm, seed=seed + 1, salt=0.002)
    outs, dil = rh.ref_post_all_frames(list(fr), list(mk), list(inp), mask_dilation_iter=n, keep_unmasked_original=True,
                                       feather_px=f)
    np.savez_compressed(os.path.join(OUT, "other_mask_size.npz"), args=np.array([t, h0, w0, h, w, n, seed], np.int64),
                        mask_hw=np.array([hm, wm], np.int64), feather=np.float64(f), dilated=np.stack(dil), out=np.stack(outs))
    print("other_mask_size", np.stack(dil).shape, np.stack(outs).shape)


def main_paint():
    """tests/golden/paint.npz: the UNMODIFIED sam2_masker.run_sam2_on_frames (stub SAM2 predictor that
    replays seeded logits) -> colour-painted mask frames (SURVEY next row N3)."""
    from oracle import painter
    rng = np.random.default_rng(5)
    t, k, mh, mw, h0, w0 = 4, 3, 45, 80, 90, 160
    obj_ids = [1, 2, 7]
    frames = [np.zeros((h0, w0, 3), np.uint8) for _ in range(t)]
    logits = {i: (rng.normal(size=(k, 1, mh, mw)) - 0.8).astype(np.float32) for i in range(t)}
    for i in range(t):
        logits[i][0, 0, 5:30, 10:50] = 2.0
        logits[i][1, 0, 20:40, 30:70] = 1.0
        logits[i][2, 0, 25:35, 40:45] = 3.0
    out, _ = painter.reference_paint(frames, logits, obj_ids)
    logits2 = {i: (rng.normal(size=(k, 1, h0, w0)) - 1.0).astype(np.float32) for i in range(2)}
    out2, _ = painter.reference_paint(frames[:2], logits2, obj_ids)
    np.savez_compressed(os.path.join(OUT, "paint.npz"), logits=np.stack([logits[i] for i in range(t)]),
                        obj_ids=np.array(obj_ids), out=np.stack(out),
                        logits_same=np.stack([logits2[i] for i in range(2)]), out_same=np.stack(out2))
    print("paint", np.stack(out).shape)


if __name__ == "__main__":
    main()
    main_paint()
    main_mask_size()

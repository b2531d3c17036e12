"""CPU oracle for the WHOLE stage sequence of one ``run_infill_on_frames`` call, networks stubbed.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference's call (/root/reference/diffuerase.py:20-114) runs, around its two networks, these pixel
stages; the in-tree ones are restated in ``prepost.py`` (pinned by the reference-run goldens), the ones
inside the un-vendored model wrappers in ``propagation.py`` / ``wrapper.py`` (PARITY UNPINNED):

    K1  binarise + dilate                                   diffuerase.py:28-31
    -- Propainter.forward (call site :52-57) --------------------------------------------------
    K2  frames -> processing size (INTER_LINEAR), masks NEAREST          row A9
        [network: RAFT + flow completion -> flows]                       stub: ``flow_fn``
    K4  bidirectional image propagation, 50+10+10 windows                row A10
    N2  to [-1,1] float, per neighbour window [network] -> 0.5/0.5 u8 merge
                                                                         stub: identity on the propagated frames
    -- DiffuEraser.forward (call site :62-67) -------------------------------------------------
    N4  read_mask (erode 3x3, dilate x iter), masked frames
        [network: diffusion]                                             stub: returns the priors
    N4  (blurred) compose
    K3  resize back + feather + composite                  diffuerase.py:70-112

``run`` is what ``tests/`` compare ``videovanish_b200`` against, and what ``bench.py`` times on the host
cores as the reference arm (``per_frame_pool`` spreads the independent per-frame stages over threads).
"""
import numpy as np

from . import prepost as op
from . import propagation as opp
from . import wrapper as ow


def _map(pool, fn, n):
    return list(pool.map(fn, range(n))) if pool is not None else [fn(i) for i in range(n)]


def propainter_prior(small, low, flows_f, flows_b, neighbor_length=10, ref_stride=10, subvideo_length=50, pool=None):
    """Image propagation + the compose loop with an identity network (pred = propagated frames)."""
    t = len(small)
    if t > 1:
        updated, _ = opp.propagate_clip_torch(small, low, flows_f, flows_b, subvideo_length)      # f32 [T,3,h,w]
    else:
        updated = opp.decode_state(opp.pack_state(small, low))[0]
    plan = opp.neighbor_plan(t, neighbor_length, ref_stride, subvideo_length)
    preds = [updated[ids[0]:ids[-1] + 1] for ids, _ in plan]
    return opp.ref_neighbor_merge(preds, plan, (low > 0).astype(np.uint8), small)


def run(frames_rgb, mask_frames, flow_fn, mask_dilation_iter=8, max_img_size=960, keep_unmasked_original=True,
        feather_px=3, blended=True, pool=None, stages=None, infer_size=None):
    """The full stage set on the host.  ``flow_fn(small u8 [T,h,w,3], low u8 [T,h,w]) -> (flows_f, flows_b)``.
    ``stages``: optional dict that receives the intermediate results by name.  ``infer_size`` overrides the
    wrapper's multiple-of-8 rule (bench.py: BASELINE's named configuration infers at 960x540)."""
    t = len(frames_rgb)
    h0, w0 = frames_rgb[0].shape[:2]
    h, w = infer_size if infer_size is not None else op.inference_size(h0, w0, max_img_size)
    dil = _map(pool, lambda i: op.ref_binarize_dilate([mask_frames[i]], mask_dilation_iter)[0], t)          # K1
    small = np.stack(_map(pool, lambda i: op.ref_resize_linear(frames_rgb[i], h, w), t))                     # K2
    low = np.stack(_map(pool, lambda i: op.ref_resize_nearest(dil[i], h, w), t))
    flows_f, flows_b = flow_fn(small, low) if t > 1 else (None, None)
    priors = propainter_prior(small, low, flows_f, flows_b, pool=pool)                                       # K4 + N2
    m = _map(pool, lambda i: ow.ref_wrapper_mask(low[i], 0), t)                                              # N4
    masked = _map(pool, lambda i: ow.ref_masked_frame(small[i], m[i]), t)
    images = priors                                              # diffusion stub: the priors are the "inpainted" frames
    comp = _map(pool, lambda i: ow.ref_wrapper_compose(images[i], small[i], m[i], blended), t)
    out = _map(pool, lambda i: op.ref_post_frame(comp[i], frames_rgb[i], dil[i], keep_unmasked_original, feather_px), t)   # K3
    if stages is not None:
        stages.update(dil=dil, small=small, low=low, priors=priors, wrapper_mask=m, masked=masked, comp=comp)
    return out

"""Run the UNMODIFIED reference ``diffuerase.run_infill_on_frames``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``/root/reference/diffuerase.py:8-9`` imports two un-vendored packages
(``diffueraser.diffueraser.DiffuEraser``, ``propainter.inference.{Propainter,
get_device}``).  This harness injects stub modules for exactly those names so
that the reference's own lines ``:26-31`` (binarise + dilate) and ``:70-112``
(resize-back, feather, composite) execute verbatim.  It only works where
``/root/reference`` exists (the build container), never on the GPU box; it is
used by ``tests/golden/make_golden.py`` and by the pinning tests (skipped when
the reference is absent).

Early-return note: ``diffuerase.py:114`` returns inside the ``for`` loop, so the
literal function post-processes frame 0 only.  ``ref_post_all_frames`` drives
the reference one frame per call (T=1 lists) so that its literal loop body is
applied to every frame; ``ref_run_literal`` keeps the multi-frame call and
therefore reproduces the bug (frames 1.. returned raw).
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("VV_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "diffuerase.py"))


class _Capture:
    """What the stub models saw / should return for the current call."""
    inpainted = None      # list handed back by DiffuEraser.forward
    seen_masks = None     # dilated masks the reference passed to the model
    seen_kwargs = None
    prior_calls = 0


def _install_stubs():
    cap = _Capture

    class DiffuEraser:                       # stands in for diffueraser.diffueraser.DiffuEraser
        def __init__(self, *a, **k):
            pass

        def forward(self, frames, masks, priors, **kw):
            cap.seen_masks = [m.copy() for m in masks]
            cap.seen_kwargs = dict(kw)
            return [f.copy() for f in cap.inpainted]

    class Propainter:                        # stands in for propainter.inference.Propainter
        def __init__(self, *a, **k):
            pass

        def forward(self, frames, masks, **kw):
            cap.prior_calls += 1
            return [f.copy() for f in frames]

    def get_device():
        return "cpu"

    m_d = types.ModuleType("diffueraser")
    m_dd = types.ModuleType("diffueraser.diffueraser")
    m_dd.DiffuEraser = DiffuEraser
    m_d.diffueraser = m_dd
    m_p = types.ModuleType("propainter")
    m_pi = types.ModuleType("propainter.inference")
    m_pi.Propainter = Propainter
    m_pi.get_device = get_device
    m_p.inference = m_pi
    sys.modules.setdefault("diffueraser", m_d)
    sys.modules.setdefault("diffueraser.diffueraser", m_dd)
    sys.modules.setdefault("propainter", m_p)
    sys.modules.setdefault("propainter.inference", m_pi)


_ref_mod = None


def load_reference():
    """Import ``/root/reference/diffuerase.py`` (unmodified) once."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    _install_stubs()
    saved = list(sys.path)
    sys.path.insert(0, REFERENCE_DIR)           # so that its `import tools` resolves to the reference's tools.py
    try:
        spec = importlib.util.spec_from_file_location(
            "_vv_reference_diffuerase", os.path.join(REFERENCE_DIR, "diffuerase.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
        sys.modules.pop("tools", None)          # do not leak the reference's `tools` into the caller
    _ref_mod = mod
    return mod


def ref_run_literal(frames, masks, inpainted, **kw):
    """One literal multi-frame call (reproduces the early return at :114)."""
    mod = load_reference()
    _Capture.inpainted = inpainted
    out = mod.run_infill_on_frames(frames, masks, propainer_frames=list(frames), **kw)
    return out, _Capture.seen_masks, _Capture.seen_kwargs


def ref_pre(masks, mask_dilation_iter=8):
    """Reference lines :26-31 on every mask frame -> list of u8 HxW in {0,255}."""
    mod = load_reference()
    import numpy as np
    H, W = masks[0].shape[:2]
    dummy = [np.zeros((H, W, 3), np.uint8) for _ in masks]
    _Capture.inpainted = dummy
    mod.run_infill_on_frames(dummy, masks, mask_dilation_iter=mask_dilation_iter,
                             propainer_frames=dummy, keep_unmasked_original=False)
    return _Capture.seen_masks


def ref_post_all_frames(frames, masks, inpainted, mask_dilation_iter=8,
                        keep_unmasked_original=True, feather_px=3):
    """Reference loop body :71-112 applied to EVERY frame (one T=1 call each)."""
    mod = load_reference()
    outs, dil = [], []
    for f, m, inp in zip(frames, masks, inpainted):
        _Capture.inpainted = [inp]
        o = mod.run_infill_on_frames([f], [m], mask_dilation_iter=mask_dilation_iter,
                                     propainer_frames=[f],
                                     keep_unmasked_original=keep_unmasked_original,
                                     feather_px=feather_px)
        outs.append(o[0])
        dil.append(_Capture.seen_masks[0])
    return outs, dil

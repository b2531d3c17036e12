"""Device-resident operators: thin torch-tensor front-ends of the C ABI.

torch is used for device memory and streams only; every byte of arithmetic happens in
``csrc/*.cu``.  All tensors are contiguous uint8 NHWC on a CUDA device; kernels are enqueued on
``torch.cuda.current_stream()``.  Names and argument meaning follow the reference's stages
(/root/reference/diffuerase.py:26-31, :70-112) and SURVEY.md section 8b.
"""
import ctypes

import torch

from . import _lib
from ._lib import VV_INTER_LINEAR, VV_INTER_NEAREST, lib

INTER_NEAREST = VV_INTER_NEAREST      # same numeric values as cv2.INTER_NEAREST / cv2.INTER_LINEAR
INTER_LINEAR = VV_INTER_LINEAR


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("videovanish_b200.ops: expected a CUDA tensor (there is no CPU fallback)")
        if not t.is_contiguous():
            raise ValueError("videovanish_b200.ops: tensors must be contiguous")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def inference_size(h0, w0, max_img_size=960):
    """(h, w) the model wrapper resizes to."""
    h, w = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.vv_inference_size(h0, w0, max_img_size, ctypes.byref(h), ctypes.byref(w)), "vv_inference_size")
    return h.value, w.value


def binarize_dilate(mask, iterations=8, lowres_size=None, return_bits=False):
    """K1.  mask u8 [T,H,W,C] (or [T,H,W]) -> u8 [T,H,W] in {0,255}   (diffuerase.py:28-31).
    With ``lowres_size=(h, w)`` also returns the INTER_NEAREST down-sized mask [T,h,w].
    ``return_bits=True`` returns ``(out, low_or_None, bits)`` where ``bits`` is the dilated mask as the 1-bit
    plane i32 [T,H,ceil(W/32)] the pass computes anyway (``upscale_feather_composite(mask_bits=...)``)."""
    if mask.dim() == 3:
        mask = mask.unsqueeze(-1)
    _require_cuda(mask)
    if mask.dtype != torch.uint8 or mask.dim() != 4:
        raise ValueError("binarize_dilate: mask must be uint8 [T,H,W,C]")
    t, h, w, c = mask.shape
    with torch.cuda.device(mask.device):
        out = torch.empty((t, h, w), dtype=torch.uint8, device=mask.device)
        low = None
        lh = lw = 0
        if lowres_size is not None:
            lh, lw = int(lowres_size[0]), int(lowres_size[1])
            low = torch.empty((t, lh, lw), dtype=torch.uint8, device=mask.device)
        bits = torch.empty((t, h, (w + 31) // 32), dtype=torch.int32, device=mask.device) if return_bits else None
        nbytes = lib.vv_binarize_dilate_workspace_bytes(t, h, w)
        ws = _ws(nbytes, mask.device)
        _lib.check(lib.vv_binarize_dilate_ex(_ptr(mask), t, h, w, c, int(iterations), _ptr(out), _ptr(low), lh, lw,
                                             _ptr(bits), _ptr(ws), nbytes, _stream()), "vv_binarize_dilate")
    if return_bits:
        return out, low, bits
    return out if low is None else (out, low)


def resize(src, h, w, interpolation=INTER_LINEAR):
    """K2.  u8 [T,H,W,C] -> u8 [T,h,w,C], cv2.resize semantics (bit-exact)."""
    squeeze = src.dim() == 3
    if squeeze:
        src = src.unsqueeze(-1)
    _require_cuda(src)
    if src.dtype != torch.uint8 or src.dim() != 4:
        raise ValueError("resize: src must be uint8 [T,H,W,C]")
    t, hh, ww, c = src.shape
    with torch.cuda.device(src.device):
        dst = torch.empty((t, int(h), int(w), c), dtype=torch.uint8, device=src.device)
        nbytes = lib.vv_resize_workspace_bytes(int(h), int(w))
        ws = _ws(nbytes, src.device)
        _lib.check(lib.vv_resize(_ptr(src), t, hh, ww, c, _ptr(dst), int(h), int(w), int(interpolation), _ptr(ws),
                                 nbytes, _stream()), "vv_resize")
    return dst.squeeze(-1) if squeeze else dst


def upscale_feather_composite(inpainted, orig, mask, feather_px=3, keep_unmasked_original=True, out=None, mask_bits=None,
                              chain_previous=False):
    """K3.  inpainted u8 [T,h,w,3], orig u8 [T,H0,W0,3], mask u8 [T,H0,W0] -> u8 [T,H0,W0,3]
    (diffuerase.py:70-112 applied to every frame).  ``mask_bits``: the same mask as K1's 1-bit plane
    (``binarize_dilate(return_bits=True)``); kernels that can use it skip the u8 mask.
    ``chain_previous=True``: this call directly follows another K3 call on the same stream that produced OTHER frames
    (a clip composited in parts); its launch may then begin while the previous one drains."""
    _require_cuda(inpainted, orig, mask, mask_bits)
    t, h, w, _ = inpainted.shape
    if keep_unmasked_original:
        if orig is None or mask is None:
            raise ValueError("upscale_feather_composite: orig and mask are required when keep_unmasked_original")
        h0, w0 = orig.shape[1:3]
        if tuple(mask.shape) != (t, h0, w0) or orig.shape[0] != t:
            raise ValueError("upscale_feather_composite: shape mismatch")
        if mask_bits is not None and (tuple(mask_bits.shape) != (t, h0, (w0 + 31) // 32) or mask_bits.dtype != torch.int32):
            raise ValueError("upscale_feather_composite: mask_bits must be int32 [T,H0,ceil(W0/32)]")
    else:
        h0, w0 = (orig.shape[1:3] if orig is not None else mask.shape[1:3])
    with torch.cuda.device(inpainted.device):
        if out is None:
            out = torch.empty((t, h0, w0, 3), dtype=torch.uint8, device=inpainted.device)
        nbytes = lib.vv_composite_workspace_bytes(h0, w0)
        ws = _ws(nbytes, inpainted.device)
        if chain_previous:
            _lib.set_option("k3_chain", 1)
        try:
            _lib.check(lib.vv_upscale_feather_composite_bits(
                _ptr(inpainted), t, h, w, _ptr(orig) if keep_unmasked_original else None,
                _ptr(mask) if keep_unmasked_original else None, _ptr(mask_bits) if keep_unmasked_original else None,
                h0, w0, float(feather_px), 1 if keep_unmasked_original else 0, _ptr(out), _ptr(ws), nbytes, _stream()),
                "vv_upscale_feather_composite")
        finally:
            if chain_previous:
                _lib.set_option("k3_chain", 0)
    return out


def mask_row_bounds(mask_bits, margin=0):
    """Per frame, the rows a composite with this (dilated) mask can change: i32 [T,2] = (lo, hi), hi exclusive, from K1's
    1-bit plane i32 [T,H,ceil(W/32)]; ``margin`` rows (the feather radius) are added either side; (0, 0) = empty mask."""
    _require_cuda(mask_bits)
    if mask_bits.dtype != torch.int32 or mask_bits.dim() != 3:
        raise ValueError("mask_row_bounds: mask_bits must be int32 [T,H,ceil(W/32)]")
    t, h, wp = mask_bits.shape
    with torch.cuda.device(mask_bits.device):
        out = torch.empty((t, 2), dtype=torch.int32, device=mask_bits.device)
        _lib.check(lib.vv_mask_row_bounds(_ptr(mask_bits), t, h, wp, int(margin), _ptr(out), _stream()), "vv_mask_row_bounds")
    return out


def subvideo_plan(video_length, subvideo_length=50, pad_len=10):
    """Windows of propainter/inference.py's image-propagation loop: (s_f, e_f, pad_s, pad_e)."""
    sub = min(100, subvideo_length)
    if video_length <= sub:
        return [(0, video_length, 0, 0)]
    plan = []
    for f in range(0, video_length, sub):
        s_f = max(0, f - pad_len)
        e_f = min(video_length, f + sub + pad_len)
        plan.append((s_f, e_f, f - s_f, e_f - min(video_length, f + sub)))
    return plan


def propagate(frames, masks, flows_f, flows_b, subvideo_length=50, pad_len=10, keep_pads=False, out=None):
    """K4.  frames u8 [N,h,w,3], masks u8 [N,h,w] (>0 = hole), flows f32 [N-1,h,w,2] ->
    packed u32 [N,h,w] (R | G<<8 | B<<16 | state<<24) of the forward propagation pass.  The pad
    frames of every sub-video window are discarded like upstream does: the kernels keep their state
    in the workspace and write the kept frames straight into the [N,h,w] result
    (``keep_pads=True`` returns the raw windows, pads included, concatenated)."""
    _require_cuda(frames, masks, flows_f, flows_b, out)
    n, h, w, _ = frames.shape
    plan = subvideo_plan(n, subvideo_length, pad_len)
    ints = lambda v: (ctypes.c_int * len(plan))(*v)
    starts, lens = ints([p[0] for p in plan]), ints([p[1] - p[0] for p in plan])
    total = sum(p[1] - p[0] for p in plan)
    if keep_pads:
        kstart = klen = None
        n_out = total
    else:
        kstart, klen = ints([p[2] for p in plan]), ints([p[1] - p[0] - p[2] - p[3] for p in plan])
        n_out = n
    with torch.cuda.device(frames.device):
        if out is None:
            out = torch.empty((n_out, h, w), dtype=torch.int32, device=frames.device)
        elif tuple(out.shape) != (n_out, h, w) or out.dtype != torch.int32:
            raise ValueError("propagate: out must be int32 [%d,%d,%d]" % (n_out, h, w))
        nbytes = lib.vv_propagate_workspace_bytes(total, total - n_out, h, w)
        ws = _ws(nbytes, frames.device)
        _lib.check(lib.vv_propagate(_ptr(frames), _ptr(masks), _ptr(flows_f) if n > 1 else None,
                                    _ptr(flows_b) if n > 1 else None, n, h, w, starts, lens, kstart, klen, len(plan),
                                    _ptr(out), _ptr(ws), nbytes, _stream()), "vv_propagate")
    return out


def propagate_unpack(packed, zero_level=127):
    """packed u32 [N,h,w] -> (rgb u8 [N,h,w,3], hole mask u8 [N,h,w] in {0,255})."""
    _require_cuda(packed)
    n, h, w = packed.shape
    with torch.cuda.device(packed.device):
        rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device=packed.device)
        hole = torch.empty((n, h, w), dtype=torch.uint8, device=packed.device)
        _lib.check(lib.vv_propagate_unpack(_ptr(packed), n * h * w, int(zero_level), _ptr(rgb), _ptr(hole), _stream()),
                   "vv_propagate_unpack")
    return rgb, hole


def chunk_blend(tail, head, k0=0, overlap_total=None, out=None):
    """K5.  tail/head u8 [O,H,W,C] (earlier chunk's last O frames / later chunk's first O)
    -> blended u8 [O,H,W,C].  Either input may be an ``int`` device address instead of a tensor:
    a peer GPU's buffer mapped through CUDA IPC, read in place over NVLink."""
    ref = tail if isinstance(tail, torch.Tensor) else head
    if not isinstance(ref, torch.Tensor):
        ref = out
    _require_cuda(ref, out, *(x for x in (tail, head) if isinstance(x, torch.Tensor)))
    o = ref.shape[0]
    frame_bytes = ref[0].numel() * ref.element_size()
    if overlap_total is None:
        overlap_total = k0 + o
    with torch.cuda.device(ref.device):
        if out is None:
            out = torch.empty_like(ref)
        pa = tail if isinstance(tail, int) else tail.data_ptr()
        pb = head if isinstance(head, int) else head.data_ptr()
        _lib.check(lib.vv_chunk_blend(ctypes.c_void_p(pa), ctypes.c_void_p(pb), o, frame_bytes, int(k0),
                                      int(overlap_total), _ptr(out), _stream()), "vv_chunk_blend")
    return out


def paint_masks(masks, colors, out_size=None):
    """N3.  masks u8 / f32-logits [T,K,mh,mw] (object k in ascending id order), colors K x 3 ints ->
    u8 [T,H0,W0,3] colour-painted mask frames (sam2_masker.py:151-175; highest object wins)."""
    _require_cuda(masks)
    if masks.dim() != 4 or masks.dtype not in (torch.uint8, torch.float32, torch.bool):
        raise ValueError("paint_masks: masks must be uint8 / bool / float32 [T,K,mh,mw]")
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    t, k, mh, mw = masks.shape
    h0, w0 = (mh, mw) if out_size is None else (int(out_size[0]), int(out_size[1]))
    cols = (ctypes.c_ubyte * (3 * k))(*[int(v) & 255 for c in colors for v in c])
    with torch.cuda.device(masks.device):
        out = torch.empty((t, h0, w0, 3), dtype=torch.uint8, device=masks.device)
        nbytes = lib.vv_paint_masks_workspace_bytes(h0, w0)
        ws = _ws(nbytes, masks.device)
        _lib.check(lib.vv_paint_masks(_ptr(masks), 1 if masks.dtype == torch.float32 else 0, t, k, mh, mw, cols, _ptr(out),
                                      h0, w0, _ptr(ws), nbytes, _stream()), "vv_paint_masks")
    return out


def wrapper_mask(mask, dilation_iter=0):
    """N4.  The DiffuEraser wrapper's read_mask on u8 [T,h,w]: (mask > 0), 3x3 erode once, 3x3 dilate
    ``dilation_iter`` times, {0,255} (the step inside the model call at diffuerase.py:62-67)."""
    _require_cuda(mask)
    if mask.dim() != 3 or mask.dtype != torch.uint8:
        raise ValueError("wrapper_mask: mask must be uint8 [T,h,w]")
    mask = mask.contiguous()
    t, h, w = mask.shape
    with torch.cuda.device(mask.device):
        out = torch.empty_like(mask)
        _lib.check(lib.vv_wrapper_mask(_ptr(mask), t, h, w, int(dilation_iter), _ptr(out), _stream()), "vv_wrapper_mask")
    return out


def wrapper_compose(img, frames, mask255, blended=True, out=None):
    """N4.  The wrapper's compose: model output ``img`` over the resized originals ``frames`` (u8 [T,h,w,3])
    through the (optionally 21x21-Gaussian-softened) mask u8 [T,h,w]."""
    _require_cuda(img)
    if img.shape != frames.shape or img.dim() != 4 or img.shape[3] != 3 or mask255.shape != img.shape[:3]:
        raise ValueError("wrapper_compose: img / frames must be [T,h,w,3] and mask255 [T,h,w]")
    if img.dtype != torch.uint8 or frames.dtype != torch.uint8 or mask255.dtype != torch.uint8:
        raise ValueError("wrapper_compose: uint8 tensors required")
    img, frames, mask255 = img.contiguous(), frames.contiguous(), mask255.contiguous()
    t, h, w, _ = img.shape
    with torch.cuda.device(img.device):
        if out is None:
            out = torch.empty_like(img)
        _lib.check(lib.vv_wrapper_compose(_ptr(img), _ptr(frames), _ptr(mask255), t, h, w, 1 if blended else 0, _ptr(out),
                                          _stream()), "vv_wrapper_compose")
    return out


def propagate_to_float(packed, want_mask=True):
    """N2.  K4's packed state [N,h,w] -> (f32 [N,3,h,w] in [-1,1], f32 hole mask [N,h,w])."""
    _require_cuda(packed)
    n, h, w = packed.shape
    with torch.cuda.device(packed.device):
        rgb = torch.empty((n, 3, h, w), dtype=torch.float32, device=packed.device)
        hole = torch.empty((n, h, w), dtype=torch.float32, device=packed.device) if want_mask else None
        _lib.check(lib.vv_propagate_to_float(_ptr(packed), n, h, w, _ptr(rgb), _ptr(hole), _stream()),
                   "vv_propagate_to_float")
    return (rgb, hole) if want_mask else rgb


def neighbor_merge(pred, mask, ori, comp, first):
    """N2.  One sliding window of the ProPainter network output merged into the running result, in place:
    pred f32 [L,3,h,w] in [-1,1], mask u8 [L,h,w] (>0 = masked), ori / comp u8 [L,h,w,3];
    ``first[l]`` = frame l has not been composed before (then comp = img, else the 0.5 / 0.5 average)."""
    _require_cuda(pred, mask, ori, comp)
    l, c, h, w = pred.shape
    if c != 3 or pred.dtype != torch.float32 or tuple(mask.shape) != (l, h, w) or tuple(ori.shape) != (l, h, w, 3) \
            or tuple(comp.shape) != (l, h, w, 3) or len(first) != l:
        raise ValueError("neighbor_merge: pred f32 [L,3,h,w], mask u8 [L,h,w], ori / comp u8 [L,h,w,3], first [L]")
    if mask.dtype != torch.uint8 or ori.dtype != torch.uint8 or comp.dtype != torch.uint8:
        raise ValueError("neighbor_merge: uint8 tensors required")
    bits = sum(1 << i for i, f in enumerate(first) if f)
    with torch.cuda.device(pred.device):
        _lib.check(lib.vv_neighbor_merge(_ptr(pred), _ptr(mask), _ptr(ori), _ptr(comp), l, h, w, bits, _stream()),
                   "vv_neighbor_merge")
    return comp


def apply_mask(frames, mask, out=None):
    """N4.  frames u8 [T,h,w,3] with the pixels of mask u8 [T,h,w] > 0 set to zero (frame * (1 - m))."""
    _require_cuda(frames, mask, out)
    t, h, w, c = frames.shape
    if c != 3 or tuple(mask.shape) != (t, h, w) or frames.dtype != torch.uint8 or mask.dtype != torch.uint8:
        raise ValueError("apply_mask: frames u8 [T,h,w,3], mask u8 [T,h,w]")
    with torch.cuda.device(frames.device):
        if out is None:
            out = torch.empty_like(frames)
        _lib.check(lib.vv_apply_mask(_ptr(frames), _ptr(mask), t, h, w, _ptr(out), _stream()), "vv_apply_mask")
    return out


def swap_rb(frames, out=None):
    """N1.  BGR <-> RGB on u8 [...,3] (tools.py:21 / :43); ``out`` may be ``frames`` itself."""
    _require_cuda(frames, out)
    if frames.dtype != torch.uint8 or frames.shape[-1] != 3:
        raise ValueError("swap_rb: uint8 [...,3] required")
    with torch.cuda.device(frames.device):
        if out is None:
            out = torch.empty_like(frames)
        _lib.check(lib.vv_swap_rb(_ptr(frames), _ptr(out), frames.numel() // 3, _stream()), "vv_swap_rb")
    return out

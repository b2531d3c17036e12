"""Host-buffer pipeline front-end (``vv_pipeline_*`` in include/vvb200.h).

The reference's interface is lists of per-frame numpy arrays; this class
turns such lists into arrays of host pointers and lets the native runtime in
``csrc/pipeline.cu`` overlap H2D copies, kernels and D2H copies over several streams.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib


# results larger than this are returned in pageable memory (page-locking tens of GB starves the host)
PINNED_RESULT_LIMIT = 8 << 30


def _ptr_array(frames, shape=None, is_mask=False):
    """Per-frame host pointers.  Returns (ctypes array, keep-alive list).  Masks that are not uint8 are
    binarised on their own dtype first (the reference tests ``m > 0`` before any cast, diffuerase.py:29:
    a float mask in (0,1) or an int16 value of 256 must stay "set")."""
    keep = []
    arr = (ctypes.c_void_p * len(frames))()
    for i, f in enumerate(frames):
        if isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.flags.c_contiguous:
            a = f
        elif is_mask and np.asarray(f).dtype != np.uint8:
            a = np.ascontiguousarray(np.asarray(f) > 0, dtype=np.uint8)
        else:
            a = np.ascontiguousarray(f, dtype=np.uint8)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError("frame %d has shape %s, expected %s" % (i, a.shape, tuple(shape)))
        keep.append(a)
        arr[i] = a.ctypes.data
    return arr, keep


def pinned_frames(t, shape):
    """``t`` uint8 frames of ``shape`` as numpy views of ONE page-locked block: D2H copies land in
    them directly and they satisfy the reference's contract (C-contiguous uint8 HxW[x3] arrays,
    kept alive by the views themselves)."""
    if not torch.cuda.is_available():
        raise RuntimeError("videovanish_b200: CUDA device required (there is no CPU fallback)")
    nbytes = t * int(np.prod(shape))
    whole = None
    if nbytes <= PINNED_RESULT_LIMIT:
        try:
            whole = torch.empty((t,) + tuple(shape), dtype=torch.uint8, pin_memory=True).numpy()
        except RuntimeError:
            whole = None        # page-locking refused (ulimit / memory pressure)
    if whole is None:
        # very long clips (BASELINE config 5: 5 000 frames = 31 GB per stream): ordinary host memory; the
        # native pipeline then copies results out of its pinned ring with its memcpy pool
        whole = np.empty((t,) + tuple(shape), np.uint8)
    return [whole[i] for i in range(t)]


class HostPipeline:
    def __init__(self, h0, w0, device=None, frames_per_batch=8, n_slots=3):
        if not torch.cuda.is_available():
            raise RuntimeError("videovanish_b200: CUDA device required (there is no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.geometry = (int(h0), int(w0))
        handle = ctypes.c_void_p()
        _lib.check(lib.vv_pipeline_create(ctypes.byref(handle), self.device, int(h0), int(w0),
                                          int(frames_per_batch), int(n_slots)), "vv_pipeline_create")
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            lib.vv_pipeline_destroy(self._h)
            self._h = None

    __del__ = close

    def pre(self, mask_frames, iterations, lowres_size=None):
        """diffuerase.py:28-31 over a list of HxWxC masks -> list of HxW u8 {0,255}; with
        ``lowres_size=(h, w)`` also the INTER_NEAREST down-sized masks."""
        h0, w0 = self.geometry
        c = 1 if mask_frames[0].ndim == 2 else mask_frames[0].shape[2]
        shape = (h0, w0) if mask_frames[0].ndim == 2 else (h0, w0, c)
        src, keep = _ptr_array(mask_frames, shape, is_mask=True)
        t = len(mask_frames)
        dil = pinned_frames(t, (h0, w0))
        dptr, _ = _ptr_array(dil)
        low, lptr, lh, lw = None, None, 0, 0
        if lowres_size is not None:
            lh, lw = int(lowres_size[0]), int(lowres_size[1])
            low = pinned_frames(t, (lh, lw))
            lptr = _ptr_array(low)[0]
        _lib.check(lib.vv_pipeline_pre(self._h, src, t, c, int(iterations), dptr, lptr, lh, lw), "vv_pipeline_pre")
        del keep
        return dil if low is None else (dil, low)

    def downsize(self, frames, h, w):
        """Row A9: list of H0xW0x3 frames -> list of hxwx3 frames (cv2 INTER_LINEAR semantics)."""
        h0, w0 = self.geometry
        src, keep = _ptr_array(frames, (h0, w0, 3))
        out = pinned_frames(len(frames), (int(h), int(w), 3))
        _lib.check(lib.vv_pipeline_downsize(self._h, src, len(frames), int(h), int(w), _ptr_array(out)[0]),
                   "vv_pipeline_downsize")
        del keep
        return out

    def post(self, inpainted, orig, dilated=None, feather_px=3, keep_unmasked_original=True):
        """diffuerase.py:70-112 over lists; ``dilated=None`` reuses the masks left on the device
        by the preceding ``pre`` call."""
        h0, w0 = self.geometry
        t = len(inpainted)
        h, w = inpainted[0].shape[:2]
        iptr, k1 = _ptr_array(inpainted, (h, w, 3))
        optr, k2 = _ptr_array(orig, (h0, w0, 3)) if keep_unmasked_original else (None, None)
        mptr, k3 = _ptr_array(dilated, (h0, w0), is_mask=True) if (dilated is not None and keep_unmasked_original) \
            else (None, None)
        out = pinned_frames(t, (h0, w0, 3))
        _lib.check(lib.vv_pipeline_post(self._h, iptr, int(h), int(w), optr, mptr, t, float(feather_px),
                                        1 if keep_unmasked_original else 0, _ptr_array(out)[0]), "vv_pipeline_post")
        del k1, k2, k3
        return out

    # ---- device-resident clips (wrappers.py): lists of host frames <-> one contiguous device tensor
    def upload(self, frames, shape, out=None, is_mask=False):
        """List of host arrays of ``shape`` -> u8 device tensor [T, *shape].  The copies are ordered before
        whatever is enqueued on the current torch stream afterwards; page-locked sources are read by DMA
        asynchronously, so the list must stay alive until that stream has been synchronised."""
        if hasattr(frames, "tensor"):                 # tools.DeviceFrames: already resident
            if tuple(frames.tensor.shape[1:]) != tuple(shape):
                raise ValueError("device clip has shape %s, expected [T,%s]" % (tuple(frames.tensor.shape), tuple(shape)))
            return frames.tensor
        t = len(frames)
        src, keep = _ptr_array(frames, shape, is_mask=is_mask)
        nbytes = int(np.prod(shape))
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((t,) + tuple(shape), dtype=torch.uint8, device="cuda")
            _lib.check(lib.vv_pipeline_upload(self._h, src, t, nbytes, ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "vv_pipeline_upload")
        self._inflight = keep        # keep converted copies alive until the next call
        return out

    def last_rows(self):
        """(rows per direction the last ``post`` moved over PCIe, frames x H0): equal unless it was row bounded."""
        a, b = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(lib.vv_pipeline_last_rows(self._h, ctypes.byref(a), ctypes.byref(b)), "vv_pipeline_last_rows")
        return int(a.value), int(b.value)

    def host_rows_begin(self, dst, src, lo, hi):
        """Starts copying, in the background, the rows OUTSIDE [lo[i], hi[i]) of every input frame ``src[i]`` into the
        result frame ``dst[i]`` (lists of C-contiguous uint8 host arrays of one shape).  ``download_rows`` joins it."""
        t = len(dst)
        shape = tuple(dst[0].shape)
        row_bytes = int(np.prod(shape[1:]))
        lo_a = (ctypes.c_int * t)(*[int(v) for v in lo])
        hi_a = (ctypes.c_int * t)(*[int(v) for v in hi])
        src_arr, keep = _ptr_array(src, shape)
        self._rows_keep = (keep, src, dst)                 # sources AND destinations stay alive until the copy has been joined
        _lib.check(lib.vv_pipeline_host_rows_begin(self._h, t, shape[0], row_bytes, _ptr_array(dst)[0], src_arr, lo_a, hi_a),
                   "vv_pipeline_host_rows_begin")

    def download_rows(self, tensor, dst, lo, hi):
        """Rows [lo[i], hi[i]) of every frame of the u8 device tensor [T,H,...] -> the page-locked host arrays ``dst``;
        waits for the work enqueued on the current torch stream and for ``host_rows_begin``'s copy."""
        t = tensor.shape[0]
        shape = tuple(tensor.shape[1:])
        lo_a = (ctypes.c_int * t)(*[int(v) for v in lo])
        hi_a = (ctypes.c_int * t)(*[int(v) for v in hi])
        with torch.cuda.device(self.device):
            _lib.check(lib.vv_pipeline_download_rows(self._h, ctypes.c_void_p(tensor.data_ptr()), t, shape[0],
                                                     int(np.prod(shape[1:])), _ptr_array(dst)[0], lo_a, hi_a,
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "vv_pipeline_download_rows")
        self._rows_keep = None
        return dst

    def download(self, tensor):
        """u8 device tensor [T, ...] -> list of T host arrays (page-locked when the budget allows); waits for
        the work enqueued on the current torch stream."""
        t = tensor.shape[0]
        shape = tuple(tensor.shape[1:])
        out = pinned_frames(t, shape)
        with torch.cuda.device(self.device):
            _lib.check(lib.vv_pipeline_download(self._h, ctypes.c_void_p(tensor.data_ptr()), t, int(np.prod(shape)),
                                                _ptr_array(out)[0],
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "vv_pipeline_download")
        return out

"""ctypes binding of ``csrc/libvvb200.so`` (C ABI: ``include/vvb200.h``).

There is no fallback: if the library has not been built, importing this module raises with
the build command; if a call fails, ``check`` raises ``RuntimeError`` carrying
``vv_last_error()`` (the reference reports errors as Python exceptions too).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_ubyte, c_uint32, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvvb200.so")

VV_INTER_NEAREST = 0
VV_INTER_LINEAR = 1
VV_MAX_FEATHER = 32.0

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        "videovanish_b200: %s is missing - build it with `make -C %s` (or `python -c 'import __graft_entry__ as g; "
        "g.build()'`); there is no CPU fallback" % (LIB_PATH, os.path.join(_HERE, "csrc")))

lib = ctypes.CDLL(LIB_PATH)

_u8p = c_void_p          # device or host byte pointers are passed as integers
_pp = POINTER(c_void_p)  # arrays of per-frame host pointers

_SIGNATURES = {
    "vv_version": (c_int, []),
    "vv_last_error": (c_char_p, []),
    "vv_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "vv_launch_count": (c_ulonglong, []),
    "vv_reset_launch_count": (None, []),
    "vv_set_option": (c_int, [c_char_p, c_int]),
    "vv_get_option": (c_int, [c_char_p, POINTER(c_int)]),
    "vv_binarize_dilate_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vv_binarize_dilate": (c_int, [_u8p, c_int, c_int, c_int, c_int, c_int, _u8p, _u8p, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "vv_binarize_dilate_ex": (c_int, [_u8p, c_int, c_int, c_int, c_int, c_int, _u8p, _u8p, c_int, c_int, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "vv_resize_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vv_resize": (c_int, [_u8p, c_int, c_int, c_int, c_int, _u8p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vv_inference_size": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "vv_composite_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vv_upscale_feather_composite": (c_int, [_u8p, c_int, c_int, c_int, _u8p, _u8p, c_int, c_int, c_float, c_int, _u8p,
                                             c_void_p, c_size_t, c_void_p]),
    "vv_upscale_feather_composite_bits": (c_int, [_u8p, c_int, c_int, c_int, _u8p, _u8p, c_void_p, c_int, c_int, c_float,
                                                  c_int, _u8p, c_void_p, c_size_t, c_void_p]),
    "vv_propagate_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vv_propagate": (c_int, [_u8p, _u8p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int),
                             POINTER(c_int), POINTER(c_int), c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vv_propagate_unpack": (c_int, [c_void_p, c_size_t, c_ubyte, _u8p, _u8p, c_void_p]),
    "vv_paint_masks_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vv_paint_masks": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, _u8p, c_int, c_int, c_void_p,
                               c_size_t, c_void_p]),
    "vv_propagate_to_float": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vv_wrapper_mask": (c_int, [_u8p, c_int, c_int, c_int, c_int, _u8p, c_void_p]),
    "vv_wrapper_compose": (c_int, [_u8p, _u8p, _u8p, c_int, c_int, c_int, c_int, _u8p, c_void_p]),
    "vv_neighbor_merge": (c_int, [c_void_p, _u8p, _u8p, _u8p, c_int, c_int, c_int, c_ulonglong, c_void_p]),
    "vv_apply_mask": (c_int, [_u8p, _u8p, c_int, c_int, c_int, _u8p, c_void_p]),
    "vv_swap_rb": (c_int, [_u8p, _u8p, c_size_t, c_void_p]),
    "vv_chunk_blend": (c_int, [_u8p, _u8p, c_int, c_size_t, c_int, c_int, _u8p, c_void_p]),
    "vv_halo_blend": (c_int, [_u8p, c_int, c_size_t, c_int, _u8p, _u8p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "vv_pipeline_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]),
    "vv_pipeline_destroy": (None, [c_void_p]),
    "vv_pipeline_pre": (c_int, [c_void_p, _pp, c_int, c_int, c_int, _pp, _pp, c_int, c_int]),
    "vv_pipeline_downsize": (c_int, [c_void_p, _pp, c_int, c_int, c_int, _pp]),
    "vv_pipeline_post": (c_int, [c_void_p, _pp, c_int, c_int, _pp, _pp, c_int, c_float, c_int, _pp]),
    "vv_pipeline_upload": (c_int, [c_void_p, _pp, c_int, c_size_t, _u8p, c_void_p]),
    "vv_pipeline_download": (c_int, [c_void_p, _u8p, c_int, c_size_t, _pp, c_void_p]),
    "vv_pipeline_last_rows": (c_int, [c_void_p, POINTER(ctypes.c_longlong), POINTER(ctypes.c_longlong)]),
    "vv_chamfer_table": (c_int, [c_int, POINTER(c_float)]),
    "vv_mask_row_bounds": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vv_pipeline_host_rows_begin": (c_int, [c_void_p, c_int, c_int, c_size_t, _pp, _pp, POINTER(c_int), POINTER(c_int)]),
    "vv_pipeline_download_rows": (c_int, [c_void_p, _u8p, c_int, c_int, c_size_t, _pp, POINTER(c_int), POINTER(c_int), c_void_p]),
    "vv_ipc_get_handle": (c_int, [c_void_p, c_void_p, POINTER(c_size_t)]),
    "vv_ipc_open_handle": (c_int, [c_void_p, POINTER(c_void_p)]),
    "vv_ipc_close_handle": (c_int, [c_void_p]),
}

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED = tuple(_SIGNATURES)


def last_error():
    msg = lib.vv_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libvvb200 %s failed (code %d): %s" % (what, rc, last_error()))


def launch_count():
    return int(lib.vv_launch_count())


def reset_launch_count():
    lib.vv_reset_launch_count()


def set_option(name, value):
    check(lib.vv_set_option(name.encode(), int(value)), "vv_set_option")


def get_option(name):
    v = c_int()
    check(lib.vv_get_option(name.encode(), ctypes.byref(v)), "vv_get_option")
    return v.value

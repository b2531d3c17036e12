"""CPU oracle for rows A1-A9 of SURVEY.md section 8a.  TEST INFRASTRUCTURE ONLY.

Two layers, both pure CPU:

* ``ref_*`` functions restate the reference's lines with the SAME library calls
  it makes (``scipy.ndimage.binary_dilation``, ``cv2.resize``,
  ``cv2.distanceTransform`` ...).  Each cites the ``/root/reference`` line it
  follows.  They are what ``bench.py``'s CPU baseline times.
* ``model_*`` functions are closed-form numpy models of those library calls
  (L1 diamond, 11-bit fixed-point bilinear, windowed chamfer, non-FMA fp32
  blend).  They are the arithmetic spec of the CUDA kernels, and
  ``tests/test_oracle.py`` proves ``model_* == ref_*`` bit for bit.
"""
import math

import numpy as np

try:                       # the reference's own third-party deps (present in this image)
    import cv2
    import scipy.ndimage
except Exception:          # pragma: no cover - model_* still work without them
    cv2 = None
    scipy = None

INTER_LINEAR = 1
INTER_NEAREST = 0

# --------------------------------------------------------------------------------------
# A1 + A2: mask binarise + dilate                      /root/reference/diffuerase.py:26-31
# --------------------------------------------------------------------------------------


def ref_binarize_dilate(mask_frames, mask_dilation_iter=8):
    """diffuerase.py:28-31, verbatim per frame."""
    out = []
    for m in mask_frames:
        m = np.any(m > 0, axis=2).astype(np.uint8)                                  # :29
        m = scipy.ndimage.binary_dilation(m > 0, iterations=mask_dilation_iter).astype(np.uint8) * 255   # :30
        out.append(m)
    return out


def model_binarize(mask):
    """A1 (diffuerase.py:29): any channel > 0 -> {0,1}.  Accepts HxW or HxWxC."""
    if mask.ndim == 3:
        return (mask.max(axis=2) > 0).astype(np.uint8)
    return (mask > 0).astype(np.uint8)


def model_dilate_l1(b, n):
    """A2 (diffuerase.py:30) closed form.

    ``binary_dilation`` with the default cross structure, ``border_value=0`` and
    ``iterations=n`` is the L1 ball of radius n (SURVEY KAT T1); ``n < 1`` repeats
    until nothing changes, i.e. the whole (4-connected) frame fills if any pixel is
    set (KAT T2).  Computed here as n rounds of the cross so that the model does
    not share code with the kernel's decomposition.
    """
    b = b.astype(bool)
    if n < 1:
        return np.full(b.shape, 255 if b.any() else 0, np.uint8)
    cur = b.copy()
    for _ in range(n):
        nxt = cur.copy()
        nxt[1:, :] |= cur[:-1, :]
        nxt[:-1, :] |= cur[1:, :]
        nxt[:, 1:] |= cur[:, :-1]
        nxt[:, :-1] |= cur[:, 1:]
        cur = nxt
    return cur.astype(np.uint8) * 255


def model_binarize_dilate(mask_frames, n=8):
    return [model_dilate_l1(model_binarize(m), n) for m in mask_frames]


# --------------------------------------------------------------------------------------
# A9: inference size (un-vendored DiffuEraser wrapper; PARITY UNPINNED)
# --------------------------------------------------------------------------------------


def inference_size(h0, w0, max_img_size=960):
    """(h, w) the model wrapper resizes to: longest side <= max_img_size, then each
    side floored to a multiple of 8.  [recalled-upstream DiffuEraser read_video /
    resize_frames; call site diffuerase.py:62-64]"""
    w, h = w0, h0
    if max(h0, w0) > max_img_size:
        r = max(h0, w0) / max_img_size
        w, h = int(w0 / r), int(h0 / r)
    w, h = w - w % 8, h - h % 8
    return h, w


# --------------------------------------------------------------------------------------
# A3 / A9: bilinear resize (cv2.resize default INTER_LINEAR)        diffuerase.py:72-73
# --------------------------------------------------------------------------------------

_COEF_BITS = 11
_COEF_ONE = 1 << _COEF_BITS


def linear_coeffs(dst, src, clamp_coeff):
    """cv2 ``resize`` INTER_LINEAR tap table for one axis (u8 fixed-point path).

    Returns (ofs int32[dst], w0 int16[dst], w1 int16[dst]).  Mirrors the published
    algorithm of OpenCV 4.x ``resize.cpp`` (third-party, pinned de facto by the cv2
    4.13.0 in this image): ``scale = 1/(dst/src)`` in double,
    ``f = float((d+0.5)*scale-0.5)``, ``s = floor(f)``, ``f -= s``; the x axis clamps
    (s<0 -> s=0,f=0; s>=src-1 -> s=src-1,f=0) while the y axis keeps f and only
    clips the two ROW INDICES; weights are ``rint(float32(w)*2048)`` as int16.
    """
    scale = 1.0 / (float(dst) / float(src))
    ofs = np.empty(dst, np.int32)
    w0 = np.empty(dst, np.int16)
    w1 = np.empty(dst, np.int16)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(float(f)))
        f = np.float32(f - np.float32(s))
        if clamp_coeff:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0)
        c0 = np.float32(1.0) - f
        w0[d] = int(np.rint(np.float32(c0 * np.float32(_COEF_ONE))))
        w1[d] = int(np.rint(np.float32(f * np.float32(_COEF_ONE))))
        ofs[d] = s
    return ofs, w0, w1


def model_resize_linear(src, dh, dw):
    """Bit-exact model of ``cv2.resize(src, (dw, dh))`` for uint8 HxW[xC] (KAT T4/T5).

    Horizontal pass in int32 (``S[sx]*a0 + S[sx+1]*a1``), vertical pass
    ``(((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2``.  The exact x2 down-scale
    that cv2 reroutes to INTER_AREA (2x2 box, ``(a+b+c+d+2)>>2``) is the same number.
    """
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    sh, sw = src.shape[:2]
    if (sh, sw) == (dh, dw):
        out = src.copy()
        return out[..., 0] if squeeze else out
    xo, xa0, xa1 = linear_coeffs(dw, sw, True)
    yo, yb0, yb1 = linear_coeffs(dh, sh, False)
    S = src.astype(np.int32)
    x1 = np.minimum(xo + 1, sw - 1)
    hrow = S[:, xo] * xa0.astype(np.int32)[None, :, None] + S[:, x1] * xa1.astype(np.int32)[None, :, None]
    y0 = np.clip(yo, 0, sh - 1)
    y1 = np.clip(yo + 1, 0, sh - 1)
    b0 = yb0.astype(np.int32)[:, None, None]
    b1 = yb1.astype(np.int32)[:, None, None]
    out = (((b0 * (hrow[y0] >> 4)) >> 16) + ((b1 * (hrow[y1] >> 4)) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out[..., 0] if squeeze else out


def ref_resize_linear(src, dh, dw):
    """diffuerase.py:72-73 (``cv2.resize(f, (W0, H0))``, default INTER_LINEAR)."""
    if src.shape[0] == dh and src.shape[1] == dw:
        return src
    return cv2.resize(src, (dw, dh))


def nearest_ofs(dst, src):
    """cv2 INTER_NEAREST source index: min(floor(d * (1/(dst/src))), src-1) (KAT T9)."""
    scale = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst) * scale).astype(np.int64), src - 1).astype(np.int32)


def model_resize_nearest(src, dh, dw):
    """Bit-exact model of ``cv2.resize(..., interpolation=cv2.INTER_NEAREST)``
    (diffuerase.py:86, tools.py:42; the down-sized model masks of row A9)."""
    yo = nearest_ofs(dh, src.shape[0])
    xo = nearest_ofs(dw, src.shape[1])
    return src[yo][:, xo].copy()


def ref_resize_nearest(src, dh, dw):
    if src.shape[0] == dh and src.shape[1] == dw:
        return src
    return cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST)


# --------------------------------------------------------------------------------------
# A4 + A5: mask re-prep + feather alpha                               diffuerase.py:77-103
# --------------------------------------------------------------------------------------

_CHAMFER_A = np.float32(1.0)      # cv2.DIST_L2, maskSize 5: a, b, c  (OpenCV distransform.cpp)
_CHAMFER_B = np.float32(1.4)
_CHAMFER_C = np.float32(2.1969)
_cost_cache = {}


def chamfer_cost_table(radius):
    """float32 [2R+1, 2R+1]: ``tab[dy + R, dx + R]`` = what ``cv2.distanceTransform(m, DIST_L2, 5)`` yields at a pixel p
    whose only zero pixel sits at p + (dy, dx).

    The transform is the two-pass raster chamfer transform with float32 steps a = 1, b = 1.4 (diagonal), c = 2.1969
    (knight): a forward pass (rows top -> bottom, pixels left -> right; neighbours (-1, -2..2), (-2, +-1) and (0, -1)) and
    its mirror image backwards, every step one float32 addition.  Its value at p is therefore the minimum, over the zero
    pixels q and over the step sequences the two passes can realise from q to p, of the LEFT-TO-RIGHT float32 sum of the
    steps - and since float32 addition is monotone, that equals ``min_q tab[q - p]`` with the table of ONE zero pixel
    (out-of-image pixels are non-zero, "far": KAT T6).  Floating-point addition is not associative, so the table is NOT
    symmetric from d ~ 12 on (the raster order decides in which order long paths add up their steps); a
    direction-blind shortest-path table (round 1 used Dijkstra) is exact only for d < 9.  Verified bit-exact against cv2
    4.13.0 (IPP 2022.2) for every feather_px <= 32 (tests/test_oracle.py); from d >= 32 on IPP leaves a family of offsets
    next to the knight diagonals of one quadrant one ulp above this two-pass value, hence VV_MAX_FEATHER = 32.
    """
    if radius in _cost_cache:
        return _cost_cache[radius]
    n = 2 * radius + 1 + 8                    # margin: the optimal path to an offset never leaves its bounding box by more
    c = n // 2
    d = np.full((n + 4, n + 4), np.inf, np.float32)
    d[c + 2, c + 2] = 0
    up = [(-1, -2, _CHAMFER_C), (-1, -1, _CHAMFER_B), (-1, 0, _CHAMFER_A), (-1, 1, _CHAMFER_B), (-1, 2, _CHAMFER_C),
          (-2, -1, _CHAMFER_C), (-2, 1, _CHAMFER_C)]

    def row_pass(i, sgn):
        row = d[i]
        for di, dj, cost in up:               # the row(s) before, in pass direction
            cand = (np.roll(d[i + sgn * di], -sgn * dj) + cost).astype(np.float32)
            np.minimum(row[2:n + 2], cand[2:n + 2], out=row[2:n + 2])
        js = range(2, n + 2) if sgn > 0 else range(n + 1, 1, -1)
        for j in js:                          # then the pixel before, in pass direction (sequential)
            t = np.float32(row[j - sgn] + _CHAMFER_A)
            if t < row[j]:
                row[j] = t

    for i in range(2, n + 2):
        row_pass(i, 1)
    for i in range(n + 1, 1, -1):
        row_pass(i, -1)
    src = d[2:-2, 2:-2][c - radius:c + radius + 1, c - radius:c + radius + 1]     # value at offset (oy, ox) FROM the zero pixel
    tab = src[::-1, ::-1].copy()              # pixel p looks at a zero pixel at p + (dy, dx): offset of p from it is -(dy, dx)
    _cost_cache[radius] = tab
    return tab


def feather_radius(feather_px):
    """Window radius outside which the chamfer distance is >= feather_px (every
    step covers at most 2 Chebyshev units for >= 2.1969, i.e. >= 1 per unit)."""
    return max(int(math.ceil(float(feather_px))) - 1, 0)


def _window_min_dist(zero, radius):
    """min over zero pixels q within Chebyshev radius of cost(p-q); +inf if none."""
    h, w = zero.shape
    tab = chamfer_cost_table(radius)
    out = np.full((h, w), np.inf, np.float32)
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            c = tab[dy + radius, dx + radius]
            ys0, ys1 = max(0, -dy), min(h, h - dy)
            xs0, xs1 = max(0, -dx), min(w, w - dx)
            if ys0 >= ys1 or xs0 >= xs1:
                continue
            sub = out[ys0:ys1, xs0:xs1]
            np.minimum(sub, np.where(zero[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx], c, np.float32(np.inf)), out=sub)
    return out


def model_feather_alpha(m, feather_px):
    """Closed-form model of diffuerase.py:89-103 (KAT T6/T7).  ``m``: u8 HxW, >0 = masked."""
    inside = m > 0
    if not feather_px > 0:
        return inside.astype(np.float32)                                            # :101-103
    r = feather_radius(feather_px)
    d_in = _window_min_dist(~inside, r)      # distance of masked pixels to nearest unmasked (0 outside)
    d_out = _window_min_dist(inside, r)      # distance of unmasked pixels to nearest masked (0 inside)
    big = np.float32(8192.0)                  # any value >= feather_px saturates the clip
    d_in = np.where(np.isinf(d_in), big, d_in).astype(np.float32)
    d_out = np.where(np.isinf(d_out), big, d_out).astype(np.float32)
    div = np.float32(2.0 * float(feather_px))
    alpha = np.float32(0.5) + (d_in - d_out).astype(np.float32) / div              # :99 (fp32, IEEE divide)
    return np.clip(alpha, np.float32(0), np.float32(1)).astype(np.float32)          # :100


def ref_feather_alpha(m, feather_px):
    """diffuerase.py:77-103 verbatim (m is the dilated mask of frame i)."""
    if m.ndim == 3:
        m = np.any(m > 0, axis=2).astype(np.uint8)                                  # :79-80
    else:
        m = (m > 0).astype(np.uint8)                                                # :82
    _, m_bin = cv2.threshold(m, 0, 255, cv2.THRESH_BINARY)                          # :89
    inv_bin = cv2.bitwise_not(m_bin)                                                # :90
    if feather_px > 0:
        d_in = cv2.distanceTransform(m_bin, cv2.DIST_L2, 5)                         # :95
        d_out = cv2.distanceTransform(inv_bin, cv2.DIST_L2, 5)                      # :96
        alpha = 0.5 + (d_in - d_out) / (2.0 * float(feather_px))                    # :99
        alpha = np.clip(alpha, 0.0, 1.0).astype(np.float32)                         # :100
    else:
        alpha = (m_bin > 0).astype(np.float32)                                      # :103
    return alpha


# --------------------------------------------------------------------------------------
# A6: composite                                                       diffuerase.py:105-112
# --------------------------------------------------------------------------------------


def model_composite(alpha, out_u8, orig_u8):
    """fp32, two rounded products + one rounded add, round-half-even (KAT T8)."""
    a3 = alpha.astype(np.float32)[..., None]
    prod0 = (a3 * out_u8.astype(np.float32)).astype(np.float32)
    prod1 = ((np.float32(1.0) - a3).astype(np.float32) * orig_u8.astype(np.float32)).astype(np.float32)
    return np.clip(np.rint((prod0 + prod1).astype(np.float32)), 0, 255).astype(np.uint8)


def ref_composite(alpha, out_u8, orig_u8):
    """diffuerase.py:105-112 verbatim."""
    alpha3 = alpha[..., None]                                                       # :105
    out = out_u8
    orig = orig_u8
    if orig.dtype != out.dtype:
        orig = orig.astype(out.dtype)                                               # :109-110
    return np.clip(np.rint(alpha3 * out + (1.0 - alpha3) * orig), 0, 255).astype(np.uint8)   # :112


# --------------------------------------------------------------------------------------
# A3-A7: the post loop body, and the whole call
# --------------------------------------------------------------------------------------


def ref_post_frame(inp, orig, dil_mask, keep_unmasked_original=True, feather_px=3):
    """Loop body diffuerase.py:71-112 for one frame, using the reference's library calls."""
    h0, w0 = orig.shape[:2]
    f = ref_resize_linear(inp, h0, w0)                                              # :72-73
    if not keep_unmasked_original:                                                  # :75
        return f
    m = dil_mask
    if m.shape[:2] != (h0, w0):
        m2 = (np.any(m > 0, axis=2) if m.ndim == 3 else (m > 0)).astype(np.uint8)
        m = cv2.resize(m2, (w0, h0), interpolation=cv2.INTER_NEAREST)               # :85-86
    alpha = ref_feather_alpha(m, feather_px)
    return ref_composite(alpha, f, orig)


def model_post_frame(inp, orig, dil_mask, keep_unmasked_original=True, feather_px=3):
    """Same frame through the closed-form models only (no cv2 / scipy)."""
    h0, w0 = orig.shape[:2]
    f = model_resize_linear(inp, h0, w0)
    if not keep_unmasked_original:
        return f
    m = model_binarize(dil_mask)
    if m.shape[:2] != (h0, w0):
        m = model_resize_nearest(m, h0, w0)
    alpha = model_feather_alpha(m, feather_px)
    return model_composite(alpha, f, orig)


def ref_run_infill_on_frames(frames_rgb, mask_frames, diffueraser_forward, propainter_forward=None,
                             mask_dilation_iter=8, propainer_frames=None, max_img_size=960,
                             keep_unmasked_original=True, feather_px=3, prog=None, bug_compat=False):
    """Restatement of ``run_infill_on_frames`` (diffuerase.py:20-114) with the two
    model ``forward`` calls injected.  ``bug_compat=True`` reproduces the literal
    early return at :114 (only frame 0 post-processed); the default applies the
    loop body to every frame."""
    if prog is not None:
        prog(5, "dilating frames")
    dilated = ref_binarize_dilate(mask_frames, mask_dilation_iter)
    if prog is not None:
        prog(10, "loading weights")
    if propainer_frames is None:
        if prog is not None:
            prog(20, "running propainter prior")
        propainer_frames = propainter_forward(frames_rgb, dilated, ref_stride=10, neighbor_length=10,
                                              subvideo_length=50, mask_dilation=0, progress=prog)
    if prog is not None:
        prog(50, "running DiffuEraser")
    inpainted = diffueraser_forward(frames_rgb, dilated, propainer_frames, max_img_size=max_img_size,
                                    mask_dilation_iter=0, guidance_scale=None, progress=prog)
    if prog is not None:
        prog(90, "resizing and merging finished frames")
    for i, f in enumerate(inpainted):
        inpainted[i] = ref_post_frame(f, frames_rgb[i], dilated[i], keep_unmasked_original, feather_px)
        if bug_compat:
            return inpainted                                                        # :114
    return inpainted

"""CPU oracle for SURVEY "next" row N4: the DiffuEraser wrapper's mask preparation and blended
compose, the pixel steps immediately inside the model call at /root/reference/diffuerase.py:62-67.
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The wrapper (``diffueraser/diffueraser.py``) is not vendored in /root/reference
and the reference holds no test or vector for it ("[recalled-upstream]").
The spec below restates the public upstream DiffuEraser wrapper as recalled:

  read_mask      m = (mask > 0) ; cv2.erode(m, 3x3 rect, iterations=1) ;
                 cv2.dilate(m, 3x3 rect, iterations=mask_dilation_iter) ; mask = m * 255
                 (the reference passes mask_dilation_iter=0: diffuerase.py:65)
  compose        mask_blurred = cv2.GaussianBlur(mask, (21, 21), 0) / 255.            (float64)
                 soft = 1 - (1 - mask / 255.) * (1 - mask_blurred)                     (float64)
                 alpha_u8 = (soft * 255).astype(uint8)                                 (truncation)
                 a = alpha_u8.astype(float32) / 255.
                 out = (img.astype(uint8) * a + frame.astype(uint8) * (1 - a)).astype(uint8)   (float32, truncation)

What IS pinned: the OpenCV primitives.  ``model_*`` below are closed forms of cv2.erode / cv2.dilate
(3x3 rect, default borders) and of the bit-exact u8 cv2.GaussianBlur((21, 21), 0), checked against
cv2 itself in tests/test_oracle.py; the CUDA kernels implement the closed forms.
"""
import cv2
import numpy as np

f32 = np.float32

# cv2.getGaussianKernel(21, 0) in OpenCV's bit-exact Q0.8 form (getGaussianKernelBitExact: sigma = 3.5,
# rounding error diffused from the ends towards the centre, centre = 256 - rest); recovered from impulse
# responses of cv2.GaussianBlur and checked in the tests.
GAUSS21_Q8 = np.array([0, 2, 2, 4, 6, 11, 15, 20, 25, 28, 30, 28, 25, 20, 15, 11, 6, 4, 2, 2, 0], np.int64)


# ---------------------------------------------------------------------------- restatement (cv2 / numpy)
def ref_wrapper_mask(mask, dilation_iter=0):
    """read_mask of the upstream wrapper for one single-channel mask: u8 HxW -> u8 HxW in {0, 255}."""
    m = np.array(np.asarray(mask) > 0).astype(np.uint8)
    m = cv2.erode(m, cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3)), iterations=1)
    if dilation_iter > 0:
        m = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3)), iterations=int(dilation_iter))
    return m * np.uint8(255)


def ref_soft_alpha(mask255):
    """Blurred compose mask: u8 HxW in {0,255} -> u8 HxW."""
    mask255 = np.asarray(mask255)
    mask_blurred = cv2.GaussianBlur(mask255, (21, 21), 0) / 255.
    soft = 1 - (1 - mask255 / 255.) * (1 - mask_blurred)
    return (soft * 255).astype(np.uint8)


def ref_wrapper_compose(img, frame, mask255, blended=True):
    """Compose of one frame: img = model output, frame = resized original, both u8 HxWx3."""
    alpha = ref_soft_alpha(mask255) if blended else np.asarray(mask255)
    a = np.expand_dims(alpha, 2).repeat(3, axis=2).astype(np.float32) / 255.
    return (np.asarray(img).astype(np.uint8) * a + np.asarray(frame).astype(np.uint8) * (1 - a)).astype(np.uint8)


# ---------------------------------------------------------------------------- closed forms (the kernel spec)
def _shift_or_and(b, fill):
    """3x3 window stack of a bool image with out-of-image pixels = fill."""
    p = np.pad(b, 1, constant_values=fill)
    h, w = b.shape
    return [p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)]


def model_wrapper_mask(mask, dilation_iter=0):
    """cv2.erode: out-of-image pixels do not erode (border = +inf); cv2.dilate: they do not dilate
    (border = -inf); N iterations of the 3x3 rect = one (2N+1)^2 square."""
    b = np.asarray(mask) > 0
    b = np.logical_and.reduce(_shift_or_and(b, True))
    for _ in range(int(dilation_iter)):
        b = np.logical_or.reduce(_shift_or_and(b, False))
    return b.astype(np.uint8) * np.uint8(255)


def _reflect101(idx, n):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def model_gaussian21_binary(mask255):
    """Bit-exact cv2.GaussianBlur(mask, (21, 21), 0) for a {0, 255} u8 mask with h, w >= 11:
    S = sum_y k_y sum_x k_x b(x, y) over the REFLECT_101-padded image (Q0.8 taps, exact integers),
    out = (255 * S + 2^15) >> 16."""
    b = (np.asarray(mask255) > 0).astype(np.int64)
    h, w = b.shape
    assert h >= 11 and w >= 11
    p = b[_reflect101(np.arange(-10, h + 10), h)][:, _reflect101(np.arange(-10, w + 10), w)]
    hs = sum(GAUSS21_Q8[d] * p[:, d:d + w] for d in range(21))
    s = sum(GAUSS21_Q8[d] * hs[d:d + h] for d in range(21))
    return ((255 * s + 32768) >> 16).astype(np.uint8)


def alpha_lut():
    """alpha_u8 for an unmasked pixel as a function of the blurred value, in float64 like numpy:
    u8((1 - (1 - 0/255.) * (1 - b/255.)) * 255) - NOT the identity (42 of 256 values truncate to b-1)."""
    b = np.arange(256, dtype=np.uint8)
    return ((1 - (1 - 0 / 255.) * (1 - b / 255.)) * 255).astype(np.uint8)


def model_soft_alpha(mask255):
    blur = model_gaussian21_binary(mask255)
    return np.where(np.asarray(mask255) > 0, np.uint8(255), alpha_lut()[blur]).astype(np.uint8)


def model_wrapper_compose(img, frame, mask255, blended=True):
    alpha = model_soft_alpha(mask255) if blended else np.asarray(mask255)
    a = (alpha.astype(f32) / f32(255.0))[..., None]
    one_minus = (f32(1.0) - a).astype(f32)
    v = (img.astype(f32) * a).astype(f32) + (frame.astype(f32) * one_minus).astype(f32)
    return v.astype(np.uint8)


def ref_masked_frame(frame, mask255):
    """read_mask's masked image: frame * (1 - m) with m = mask > 0 (broadcast over RGB)."""
    m = (np.asarray(mask255) > 0).astype(np.uint8)[..., None]
    return (np.asarray(frame) * (1 - m)).astype(np.uint8)

"""Synthetic clips of the shapes BASELINE.json names.

All generators are seeded ``np.random.default_rng(seed)`` so the CUDA path, the
oracle and the committed golden vectors see the same bytes.  There is no model in
the loop: the "inpainted" frames are independent noise at inference resolution.
"""
import numpy as np

# name -> (T, H0, W0, inference (h, w))            BASELINE.json configs[0..4]; (h, w) = oracle.prepost.inference_size
# with max_img_size 320 for c1 and 960 for the others, except c2, which BASELINE names at 960x540 (the wrapper's
# multiple-of-8 rule gives 960x536: bench.py's `c2_production_960x536` block and the GPU tests cover that too).
# tests/test_host_logic.py checks the table against vv_inference_size.
CONFIGS = {
    "c1_360p": (64, 360, 640, (176, 320)),
    "c2_1080p": (300, 1080, 1920, (540, 960)),
    "c3_720p_flow": (500, 720, 1280, (720, 1280)),
    "c4_4k_chunked": (600, 2160, 3840, (536, 960)),
    "c5_1080p_long": (5000, 1080, 1920, (536, 960)),
}


def _gradient(h, w):
    y = np.linspace(0.0, 255.0, h, dtype=np.float32)[:, None, None]
    x = np.linspace(0.0, 255.0, w, dtype=np.float32)[None, :, None]
    c = np.array([1.0, 0.5, 0.25], np.float32)[None, None, :]
    return (y * c + x * (1.0 - c))


def frames(t, h, w, seed=0):
    """u8 [t,h,w,3]: uniform noise blended 50/50 with a smooth gradient."""
    rng = np.random.default_rng(seed)
    g = _gradient(h, w)
    out = np.empty((t, h, w, 3), np.uint8)
    for i in range(t):
        n = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[i] = ((n.astype(np.float32) + g) * 0.5).astype(np.uint8)
    return out


def noise_frames(t, h, w, seed=1):
    """u8 [t,h,w,3] independent noise (stand-in for the model's inpainted frames)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)


def masks(t, h, w, seed=2, salt=0.001):
    """u8 [t,h,w,3] colour-painted masks as sam2_masker.py produces them: black, one
    (h/4 x w/5) box painted (0,0,255) moving (+2,+6) px/frame with wrap-around, plus
    ``salt`` fraction of isolated pixels in random non-zero colours."""
    rng = np.random.default_rng(seed)
    out = np.zeros((t, h, w, 3), np.uint8)
    bh, bw = max(h // 4, 1), max(w // 5, 1)
    for i in range(t):
        y0 = (h // 3 + 2 * i) % h
        x0 = (w // 4 + 6 * i) % w
        ys = (np.arange(bh) + y0) % h
        xs = (np.arange(bw) + x0) % w
        out[i][np.ix_(ys, xs)] = (0, 0, 255)
        k = int(h * w * salt)
        if k:
            py = rng.integers(0, h, k)
            px = rng.integers(0, w, k)
            col = rng.integers(0, 256, (k, 3), dtype=np.uint8)
            col[:, rng.integers(0, 3)] |= 1          # make sure each salt pixel is non-zero somewhere
            out[i][py, px] = col
    return out


def flows(t, h, w, seed=3):
    """fp32 forward / backward flows [t-1,h,w,2] (x, y offsets in pixels).

    fwd = (3.0, -1.5) + N(0, 0.05^2); bwd = -fwd + N(0, 0.05^2); 2 % of pixels get a
    large random offset so that the forward/backward consistency check rejects
    them.  Continuous values: no exact half-integer sample positions (KAT T10)."""
    rng = np.random.default_rng(seed)
    n = max(t - 1, 0)
    base = np.array([3.0, -1.5], np.float32)
    fwd = base + rng.normal(0.0, 0.05, (n, h, w, 2)).astype(np.float32)
    bwd = -fwd + rng.normal(0.0, 0.05, (n, h, w, 2)).astype(np.float32)
    bad = rng.random((n, h, w)) < 0.02
    fwd[bad] += rng.uniform(-20.0, 20.0, (int(bad.sum()), 2)).astype(np.float32)
    return np.ascontiguousarray(fwd), np.ascontiguousarray(bwd)

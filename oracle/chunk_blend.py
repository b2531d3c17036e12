"""CPU oracle for SURVEY row A11: chunk-overlap feather blend.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED - and necessarily so: the reference advertises chunked processing
(/root/reference/README.md:18) but lists it as a TODO (README.md:76) and contains no code for
it.  The spec below is builder-defined and mirrors the composite arithmetic of
diffuerase.py:112 (fp32 products, fp32 add, round-half-even, clip):

  chunks of L frames, overlap O, stride L-O; for overlap frame k in [0, O):
      w   = f32(k+1) / f32(O+1)
      out = u8(clip(rint(f32((1-w)*A[k]) + f32(w*B[k]))))   A = earlier chunk's tail, B = later chunk's head
"""
import numpy as np

f32 = np.float32


def chunk_plan(n_frames, chunk=80, overlap=16):
    """[(start, end)] with stride chunk-overlap; the last chunk is clipped to the clip end and
    dropped if it would add no new frame.  600 frames, 80/16 -> starts 0,64,...,576 (10 chunks)."""
    if chunk <= overlap:
        raise ValueError("chunk must be longer than overlap")
    if n_frames <= chunk:
        return [(0, n_frames)]
    plan, s = [], 0
    stride = chunk - overlap
    while True:
        e = min(s + chunk, n_frames)
        plan.append((s, e))
        if e == n_frames:
            break
        s += stride
    return plan


def blend_weights(overlap, k0=0, n=None):
    n = overlap - k0 if n is None else n
    return (np.arange(k0 + 1, k0 + n + 1, dtype=f32) / f32(overlap + 1)).astype(f32)


def blend_overlap(tail, head, k0=0, overlap_total=None):
    """tail/head u8 [O,...] -> blended u8 [O,...]."""
    o = tail.shape[0]
    total = k0 + o if overlap_total is None else overlap_total
    w = blend_weights(total, k0, o).reshape((o,) + (1,) * (tail.ndim - 1))
    nw = (f32(1.0) - w).astype(f32)
    v = (nw * tail.astype(f32)).astype(f32) + (w * head.astype(f32)).astype(f32)
    return np.clip(np.rint(v.astype(f32)), 0, 255).astype(np.uint8)


def stitch_chunks(chunk_outputs, plan, overlap):
    """Concatenate per-chunk outputs (list of u8 [len_i,H,W,3]) into the full clip, blending
    the `overlap` frames shared by consecutive chunks."""
    n = plan[-1][1]
    out = np.empty((n,) + chunk_outputs[0].shape[1:], np.uint8)
    for ci, ((s, e), frames) in enumerate(zip(plan, chunk_outputs)):
        lo = 0
        if ci > 0:
            ov = plan[ci - 1][1] - s                       # frames shared with the previous chunk
            prev = chunk_outputs[ci - 1]
            out[s:s + ov] = blend_overlap(prev[len(prev) - ov:], frames[:ov], 0, ov)
            lo = ov
        out[s + lo:e] = frames[lo:]
    return out

"""Frame-chunk planning, chunk-overlap blending and multi-GPU sharding.

The reference advertises chunked processing with overlap blending (README.md:18) but has no code
for it (README.md:76), so the scheme is builder-defined: chunks of ``chunk`` frames, ``overlap``
shared frames, stride ``chunk - overlap``; shared frames are cross-faded with
``w = (k+1)/(overlap+1)`` by kernel K5 (``ops.chunk_blend``).

Multi-GPU: one process per GPU, consecutive chunks on consecutive ranks.  Frames are independent
for K1/K2/K3, so there is no data-path collective; the only exchange is the overlap halo at a
boundary between two ranks.  Each side blends half of the overlap, so each rank sends
``overlap/2`` frames to each neighbour - either with NCCL send/recv (``mode="nccl"``) or not at
all: with ``mode="peer"`` the neighbour's frames are mapped through CUDA IPC and K5 reads them in
place over NVLink while it blends.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def chunk_plan(n_frames, chunk=80, overlap=16):
    """[(start, end)] - 600 frames, 80/16 -> starts 0, 64, ..., 576 (10 chunks)."""
    if chunk <= overlap:
        raise ValueError("chunk must be longer than overlap")
    if n_frames <= chunk:
        return [(0, n_frames)]
    plan, s, stride = [], 0, chunk - overlap
    while True:
        e = min(s + chunk, n_frames)
        plan.append((s, e))
        if e == n_frames:
            break
        s += stride
    return plan


def shard_chunks(plan, world_size):
    """Contiguous blocks of chunks per rank, sizes differing by at most one: [[chunk indices]]."""
    n = len(plan)
    base, extra = divmod(n, world_size)
    out, i = [], 0
    for r in range(world_size):
        k = base + (1 if r < extra else 0)
        out.append(list(range(i, i + k)))
        i += k
    return out


def _default_blend(tail, head, k0, total, out):
    from . import ops
    return ops.chunk_blend(tail, head, k0=k0, overlap_total=total, out=out)


def stitch_chunks(chunk_outputs, plan, blend_fn=_default_blend):
    """Single-device stitch: list of u8 [len_i,H,W,C] chunk outputs -> u8 [N,H,W,C]."""
    n = plan[-1][1]
    first = chunk_outputs[0]
    out = torch.empty((n,) + tuple(first.shape[1:]), dtype=first.dtype, device=first.device)
    for ci, ((s, e), frames) in enumerate(zip(plan, chunk_outputs)):
        lo = 0
        if ci > 0:
            ov = plan[ci - 1][1] - s
            prev = chunk_outputs[ci - 1]
            blend_fn(prev[prev.shape[0] - ov:], frames[:ov], 0, ov, out[s:s + ov])
            lo = ov
        out[s + lo:e].copy_(frames[lo:])
    return out


class PeerWindow:
    """CUDA-IPC mapping of every rank's halo source buffer (mode="peer")."""

    def __init__(self, tensor, group=None):
        self.group = group
        self.tensor = tensor
        handle = (ctypes.c_ubyte * 64)()
        offset = ctypes.c_size_t()
        _lib.check(_lib.lib.vv_ipc_get_handle(ctypes.c_void_p(tensor.data_ptr()), handle, ctypes.byref(offset)),
                   "vv_ipc_get_handle")
        mine = (bytes(handle), int(offset.value))
        everyone = [None] * dist.get_world_size(group)
        dist.all_gather_object(everyone, mine, group=group)
        self.rank = dist.get_rank(group)
        self.mapped = {}
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < len(everyone):
                raw, off = everyone[r]
                ptr = ctypes.c_void_p()
                _lib.check(_lib.lib.vv_ipc_open_handle(ctypes.create_string_buffer(raw, 64), ctypes.byref(ptr)),
                           "vv_ipc_open_handle")
                self.mapped[r] = (ptr.value, off)

    def ptr(self, rank, byte_offset=0):
        base, off = self.mapped[rank]
        return base + off + byte_offset

    def close(self):
        for base, _ in self.mapped.values():
            _lib.lib.vv_ipc_close_handle(ctypes.c_void_p(base))
        self.mapped = {}


def blend_rank_boundaries(out, overlap, group=None, mode="nccl", blend_fn=_default_blend, window=None):
    """``out`` u8 [T,H,W,C] is this rank's chunk; its last ``overlap`` frames coincide with the
    first ``overlap`` frames of the next rank's chunk.  Blends, in place, the half of each boundary
    this rank owns: overlap indices [0, overlap/2) of the boundary with the next rank (stored in
    ``out[T-overlap+k]``) and [overlap/2, overlap) of the boundary with the previous rank (stored
    in ``out[k]``).  Returns the number of bytes received from / read on peers."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t = out.shape[0]
    half = overlap // 2
    if world == 1 or overlap == 0:
        return 0
    frame_bytes = out[0].numel() * out.element_size()
    has_prev, has_next = rank > 0, rank < world - 1
    moved = 0
    if mode == "peer":
        if window is None:
            raise ValueError("mode='peer' needs a PeerWindow over `out`")
        dist.barrier(group)                       # neighbours' chunks are complete
        # read the ORIGINAL neighbour frames: each side only overwrites the half it owns, and the
        # halves read remotely are the ones the neighbour does not write
        if has_next and half > 0:
            blend_fn(out[t - overlap:t - overlap + half], window.ptr(rank + 1, 0), 0, overlap,
                     out[t - overlap:t - overlap + half])
            moved += half * frame_bytes
        if has_prev and overlap - half > 0:
            src = window.ptr(rank - 1, (t - overlap + half) * frame_bytes)
            blend_fn(src, out[half:overlap], half, overlap, out[half:overlap])
            moved += (overlap - half) * frame_bytes
        torch.cuda.current_stream().synchronize()
        dist.barrier(group)                       # nobody reuses `out` while a peer still reads it
        return moved
    # mode == "nccl" (or gloo in the CPU tests): explicit halo send/recv, then blend locally
    reqs, recv_next, recv_prev = [], None, None
    if has_next and half > 0:
        recv_next = torch.empty_like(out[:half])                                  # next rank's head, k < half
        reqs.append(dist.P2POp(dist.irecv, recv_next, rank + 1, group))
        if overlap - half > 0:
            reqs.append(dist.P2POp(dist.isend, out[t - overlap + half:].contiguous(), rank + 1, group))
    if has_prev:
        if overlap - half > 0:
            recv_prev = torch.empty_like(out[:overlap - half])                    # prev rank's tail, k >= half
            reqs.append(dist.P2POp(dist.irecv, recv_prev, rank - 1, group))
        if half > 0:
            reqs.append(dist.P2POp(dist.isend, out[:half].contiguous(), rank - 1, group))
    for w in dist.batch_isend_irecv(reqs):
        w.wait()
    if recv_next is not None:
        blend_fn(out[t - overlap:t - overlap + half], recv_next, 0, overlap, out[t - overlap:t - overlap + half])
        moved += recv_next.numel()
    if recv_prev is not None:
        blend_fn(recv_prev, out[half:overlap], half, overlap, out[half:overlap])
        moved += recv_prev.numel()
    return moved

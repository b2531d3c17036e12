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
import os

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
    """Contiguous blocks of chunks per rank, sizes differing by at most one: [[chunk indices]].  The ranks at the
    END get the extra chunks: the last chunk of a clip is usually clipped short, and a rank's block must hold at
    least 2 x overlap frames for the halo blend."""
    n = len(plan)
    base, extra = divmod(n, world_size)
    out, i = [], 0
    for r in range(world_size):
        k = base + (1 if r >= world_size - extra else 0)
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


_ipc_cache = {}          # 64-byte handle -> [mapped base pointer, reference count]: a handle may only be opened once per process


def _ipc_open(raw):
    ent = _ipc_cache.get(raw)
    if ent is None:
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib.vv_ipc_open_handle(ctypes.create_string_buffer(raw, 64), ctypes.byref(ptr)), "vv_ipc_open_handle")
        ent = _ipc_cache[raw] = [ptr.value, 0]
    ent[1] += 1
    return ent[0]


def _ipc_close(raw):
    ent = _ipc_cache.get(raw)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        _lib.lib.vv_ipc_close_handle(ctypes.c_void_p(ent[0]))
        del _ipc_cache[raw]


def _ipc_export(tensor):
    handle = (ctypes.c_ubyte * 64)()
    offset = ctypes.c_size_t()
    _lib.check(_lib.lib.vv_ipc_get_handle(ctypes.c_void_p(tensor.data_ptr()), handle, ctypes.byref(offset)), "vv_ipc_get_handle")
    return bytes(handle), int(offset.value)


class PeerWindow:
    """CUDA-IPC mapping of the neighbour ranks' halo source buffers and handshake flags (mode="peer").

    Every rank contributes its frame buffer ``tensor`` [T,H,W,C] (T may differ between ranks: the last chunk of a
    clip is clipped and ``shard_chunks`` sizes differ by one), its frame count and a small zero-initialised flag
    block; the neighbours' buffers and flags are mapped into this process."""

    FLAG_WORDS = 8

    def __init__(self, tensor, group=None):
        self.group = group
        self.tensor = tensor
        self.flags = torch.zeros(self.FLAG_WORDS, dtype=torch.int32, device=tensor.device)
        self.epoch = 0
        torch.cuda.current_stream().synchronize()              # the zeros are in memory before anybody maps them
        mine = (_ipc_export(tensor), _ipc_export(self.flags), int(tensor.shape[0]))
        everyone = [None] * dist.get_world_size(group)
        dist.all_gather_object(everyone, mine, group=group)
        self.rank = dist.get_rank(group)
        self.frames = [e[2] for e in everyone]                 # T of every rank
        self.mapped, self._raw = {}, []
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < len(everyone):
                (raw_t, off_t), (raw_f, off_f), _ = everyone[r]
                self.mapped[r] = (_ipc_open(raw_t) + off_t, _ipc_open(raw_f) + off_f)
                self._raw += [raw_t, raw_f]

    def ptr(self, rank, byte_offset=0):
        return self.mapped[rank][0] + byte_offset

    def flags_ptr(self, rank):
        return self.mapped[rank][1]

    def error(self):
        """True when a halo kernel of this rank gave up waiting for a neighbour."""
        return bool(int(self.flags[4].item()))

    def close(self):
        for raw in self._raw:
            _ipc_close(raw)
        self.mapped, self._raw = {}, []


def blend_rank_boundaries(out, overlap, group=None, mode="nccl", blend_fn=_default_blend, window=None):
    """``out`` u8 [T,H,W,C] is this rank's block of frames; its last ``overlap`` frames coincide with the
    first ``overlap`` frames of the next rank's block.  Blends, in place, the half of each boundary
    this rank owns: overlap indices [0, overlap/2) of the boundary with the next rank (stored in
    ``out[T-overlap+k]``) and [overlap/2, overlap) of the boundary with the previous rank (stored
    in ``out[k]``).  Returns the number of bytes received from / read on peers.

    ``mode="peer"``: one kernel per rank and step reads the neighbours' frames in place over NVLink and
    synchronises with their kernels through flags in IPC-mapped memory (``vv_halo_blend``) - no host barrier,
    no stream synchronisation.  ``mode="nccl"`` (gloo in the CPU tests): explicit halo send/recv, then K5."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t = out.shape[0]
    half = overlap // 2
    if world == 1 or overlap == 0:
        return 0
    if t < 2 * overlap:
        raise ValueError("blend_rank_boundaries: a rank needs at least 2 x overlap frames (%d < %d): the halves it "
                         "overwrites would overlap the frames a neighbour still reads" % (t, 2 * overlap))
    frame_bytes = out[0].numel() * out.element_size()
    has_prev, has_next = rank > 0, rank < world - 1
    moved = 0
    if mode == "peer":
        if window is None:
            raise ValueError("mode='peer' needs a PeerWindow over `out`")
        if window.tensor.data_ptr() != out.data_ptr() or window.frames[rank] != t:
            raise ValueError("the PeerWindow was built over another buffer")
        if overlap < 2:
            raise ValueError("mode='peer' needs overlap >= 2")
        window.epoch += 1
        # the previous rank's tail starts at ITS frame T_prev - overlap + half
        prev_tail = window.ptr(rank - 1, (window.frames[rank - 1] - overlap + half) * frame_bytes) if has_prev else None
        next_head = window.ptr(rank + 1, 0) if has_next else None
        vp = ctypes.c_void_p
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib.vv_halo_blend(vp(out.data_ptr()), t, frame_bytes, overlap, vp(next_head), vp(prev_tail),
                                              vp(window.flags.data_ptr()),
                                              vp(window.flags_ptr(rank + 1)) if has_next else None,
                                              vp(window.flags_ptr(rank - 1)) if has_prev else None,
                                              window.epoch & 0xffffffff, vp(torch.cuda.current_stream().cuda_stream)),
                       "vv_halo_blend")
        return (half if has_next else 0) * frame_bytes + ((overlap - half) if has_prev else 0) * frame_bytes
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


_side_streams = {}
HALO_CTAS_OVERLAPPED = int(os.environ.get("VV_HALO_CTAS", "128"))


def produce_and_blend_boundaries(out, overlap, produce, group=None, mode="nccl", blend_fn=_default_blend, window=None,
                                 stream=None, events=None):
    """``blend_rank_boundaries`` for a block that is still being produced, with the exchange hidden behind the
    production: ``produce(lo, hi)`` writes ``out[lo:hi]`` on the current stream.  The frames the halo exchange touches
    (this rank's first / last ``overlap`` frames - the only ones a neighbour reads or this rank blends) are produced
    first; the exchange (NVLink peer reads + blend, or send/recv + blend) then runs on a side stream while ``produce``
    fills the rest, and the current stream joins it at the end.  ``events`` = optional (start, end) CUDA events recorded
    around the exchange on the side stream.  Returns the bytes received from / read on peers."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    t = out.shape[0]
    if world == 1 or overlap == 0 or not out.is_cuda:
        produce(0, t)
        return blend_rank_boundaries(out, overlap, group, mode=mode, blend_fn=blend_fn, window=window) if world > 1 else 0
    if t < 2 * overlap:
        raise ValueError("produce_and_blend_boundaries: a rank needs at least 2 x overlap frames (%d < %d)" % (t, 2 * overlap))
    has_prev, has_next = rank > 0, rank < world - 1
    main = torch.cuda.current_stream(out.device)
    side = stream
    if side is None:
        side = _side_streams.get(out.device.index)
        if side is None:
            # high priority: its few CTAs take the next SM slots the producing kernel frees instead of queueing behind its grid
            side = _side_streams[out.device.index] = torch.cuda.Stream(out.device, priority=-1)
    if has_prev:
        produce(0, overlap)
    if has_next:
        produce(t - overlap, t)
    ready = torch.cuda.Event()
    ready.record(main)
    side.wait_event(ready)
    with torch.cuda.stream(side):
        if events is not None:
            events[0].record(side)
        # a small grid: the exchange is NVLink-latency work and must not take the SM slots of the producing kernel
        full = _lib.get_option("k5_halo_ctas")
        _lib.set_option("k5_halo_ctas", HALO_CTAS_OVERLAPPED)
        try:
            moved = blend_rank_boundaries(out, overlap, group, mode=mode, blend_fn=blend_fn, window=window)
        finally:
            _lib.set_option("k5_halo_ctas", full)
        if events is not None:
            events[1].record(side)
        done = torch.cuda.Event()
        done.record(side)
    produce(overlap if has_prev else 0, t - overlap if has_next else t)
    main.wait_event(done)
    return moved


def run_sharded(n_frames, chunk, overlap, process_chunk, group=None, mode="peer"):
    """BASELINE config 4 as a system: ``chunk_plan`` -> ``shard_chunks`` -> every rank runs ``process_chunk(ci, s, e)
    -> u8 [e-s,H,W,C]`` (device tensor) on its own chunks, stitches them locally with K5 and blends the boundaries it
    shares with the neighbour ranks (halo only, no other exchange).

    Returns ``(block, first_frame, owned)``: this rank's frames [first_frame, first_frame + len(block)) of the clip
    and the slice of ``block`` this rank is authoritative for (each rank owns half of each shared overlap), so that
    concatenating ``block[owned]`` over the ranks gives the stitched clip."""
    plan = chunk_plan(n_frames, chunk, overlap)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if len(plan) < world:
        raise ValueError("run_sharded: %d chunks cannot be spread over %d ranks" % (len(plan), world))
    mine = shard_chunks(plan, world)[rank]
    first = plan[mine[0]][0]
    outs = [process_chunk(ci, *plan[ci]) for ci in mine]
    block = stitch_chunks(outs, [(plan[ci][0] - first, plan[ci][1] - first) for ci in mine])
    del outs
    t = block.shape[0]
    half = overlap // 2
    has_prev, has_next = rank > 0, rank < world - 1
    if world > 1:
        window = PeerWindow(block, group) if mode == "peer" else None
        blend_rank_boundaries(block, overlap, group, mode=mode, window=window)
        if window is not None:
            torch.cuda.current_stream().synchronize()
            failed = window.error()
            dist.barrier(group)                   # nobody unmaps while a neighbour's kernel may still poll
            window.close()
            if failed:
                raise RuntimeError("halo blend: a neighbour rank did not show up")
    owned = slice(half if has_prev else 0, t - (overlap - half) if has_next else t)
    return block, first, owned

"""World-size-2 (and 3) gloo test of the rank-boundary halo blend: the same
host logic that runs over NCCL on the GPUs, exercised on CPU tensors with the oracle's blend."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import chunk_blend as ocb


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, overlap, t, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from videovanish_b200 import chunking

    def blend(tail, head, k0, total, out):
        out.copy_(torch.from_numpy(ocb.blend_overlap(tail.numpy(), head.numpy(), k0, total)))

    rng = np.random.default_rng(100 + rank)
    t = t + 3 * rank                    # ranks hold different numbers of frames (clipped last chunk, uneven shards)
    mine = torch.from_numpy(rng.integers(0, 256, (t, 5, 7, 3), dtype=np.uint8))
    orig = mine.clone()
    moved = chunking.blend_rank_boundaries(mine, overlap, mode="nccl", blend_fn=blend)
    # the produce-then-exchange form (on CPU tensors: production, then the same exchange) gives the same block
    mine2 = torch.zeros_like(orig)
    moved2 = chunking.produce_and_blend_boundaries(mine2, overlap, lambda lo, hi: mine2[lo:hi].copy_(orig[lo:hi]), mode="nccl",
                                                   blend_fn=blend)
    assert torch.equal(mine2, mine) and moved2 == moved
    ret[rank] = (orig.numpy(), mine.numpy(), moved)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap", [(2, 4), (3, 6), (2, 5)])
def test_rank_boundary_blend(world, overlap):
    t = 12
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), overlap, t, ret), nprocs=world, join=True)
    half = overlap // 2
    for r in range(world - 1):
        a_orig, a_new, _ = ret[r]
        b_orig, b_new, _ = ret[r + 1]
        t = len(a_orig)
        full = ocb.blend_overlap(a_orig[t - overlap:], b_orig[:overlap])
        assert np.array_equal(a_new[t - overlap:t - overlap + half], full[:half])       # owned by rank r
        assert np.array_equal(b_new[half:overlap], full[half:])                         # owned by rank r+1
    # interior frames untouched; bytes moved = the halves received
    o0, n0, moved0 = ret[0]
    t = len(o0)
    assert np.array_equal(o0[:t - overlap], n0[:t - overlap])
    assert moved0 == half * 5 * 7 * 3


def _sharded_worker(rank, world, port, n_frames, chunk, overlap, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from videovanish_b200 import chunking

    def blend(tail, head, k0, total, out):
        out.copy_(torch.from_numpy(ocb.blend_overlap(tail.numpy(), head.numpy(), k0, total)))

    def process_chunk(ci, s, e):          # chunk-dependent content, so that the cross-fade matters
        rng = np.random.default_rng(1000 + ci)
        return torch.from_numpy(rng.integers(0, 256, (e - s, 4, 6, 3), dtype=np.uint8))

    real_stitch, real_blend = chunking.stitch_chunks, chunking.blend_rank_boundaries
    chunking.stitch_chunks = lambda outs, plan: real_stitch(outs, plan, blend_fn=blend)
    chunking.blend_rank_boundaries = lambda out, ov, group=None, mode="nccl", window=None: real_blend(
        out, ov, group, mode="nccl", blend_fn=blend)
    block, first, owned = chunking.run_sharded(n_frames, chunk, overlap, process_chunk, mode="nccl")
    ret[rank] = (first, block.numpy()[owned], (owned.start, owned.stop))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,chunk,overlap", [(2, 100, 20, 6), (3, 75, 16, 4), (2, 47, 20, 6)])
def test_config4_sharded_clip_equals_single_device_stitch(world, n_frames, chunk, overlap):
    """BASELINE config 4 as a system on CPU ranks: chunk plan -> contiguous shards -> per-rank stitch -> halo blend
    at the rank boundaries; the ranks' owned frames concatenate to exactly the single-device stitch."""
    from videovanish_b200 import chunking
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, _free_port(), n_frames, chunk, overlap, ret), nprocs=world, join=True)
    plan = chunking.chunk_plan(n_frames, chunk, overlap)
    outs = [np.random.default_rng(1000 + ci).integers(0, 256, (e - s, 4, 6, 3), dtype=np.uint8) for ci, (s, e) in enumerate(plan)]
    want = ocb.stitch_chunks(outs, plan, overlap)
    got = np.concatenate([ret[r][1] for r in range(world)])
    assert got.shape == want.shape and np.array_equal(got, want)
    firsts = [ret[r][0] + ret[r][2][0] for r in range(world)]
    assert firsts == sorted(firsts) and firsts[0] == 0

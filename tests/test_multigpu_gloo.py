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
    mine = torch.from_numpy(rng.integers(0, 256, (t, 5, 7, 3), dtype=np.uint8))
    orig = mine.clone()
    moved = chunking.blend_rank_boundaries(mine, overlap, mode="nccl", blend_fn=blend)
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
        full = ocb.blend_overlap(a_orig[t - overlap:], b_orig[:overlap])
        assert np.array_equal(a_new[t - overlap:t - overlap + half], full[:half])       # owned by rank r
        assert np.array_equal(b_new[half:overlap], full[half:])                         # owned by rank r+1
    # interior frames untouched; bytes moved = the halves received
    o0, n0, moved0 = ret[0]
    assert np.array_equal(o0[:t - overlap], n0[:t - overlap])
    assert moved0 == half * 5 * 7 * 3

#!/usr/bin/env python
"""Multi-GPU check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Verifies the rank-boundary halo blend (NCCL send/recv and CUDA-IPC peer reads inside K5) against the
oracle, and times both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import chunk_blend as ocb  # noqa: E402
from videovanish_b200 import chunking  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t, h, w, overlap = 40, 270, 480, 16
    rngs = [np.random.default_rng(100 + r) for r in range(world)]
    clips = [g.integers(0, 256, (t, h, w, 3), dtype=np.uint8) for g in rngs]       # every rank knows all inputs
    half = overlap // 2
    expect = clips[rank].copy()
    if rank < world - 1:
        full = ocb.blend_overlap(clips[rank][t - overlap:], clips[rank + 1][:overlap])
        expect[t - overlap:t - overlap + half] = full[:half]
    if rank > 0:
        full = ocb.blend_overlap(clips[rank - 1][t - overlap:], clips[rank][:overlap])
        expect[half:overlap] = full[half:]
    ok = True
    for mode in ("nccl", "peer"):
        mine = torch.from_numpy(clips[rank]).to(dev)
        window = chunking.PeerWindow(mine) if mode == "peer" else None
        moved = chunking.blend_rank_boundaries(mine, overlap, mode=mode, window=window)
        torch.cuda.synchronize()
        good = np.array_equal(mine.cpu().numpy(), expect)
        ok &= good
        # timing at 1080p, 16-frame overlap
        big = torch.randint(0, 256, (32, 1080, 1920, 3), dtype=torch.uint8, device=dev)
        bwin = chunking.PeerWindow(big) if mode == "peer" else None
        for _ in range(3):
            chunking.blend_rank_boundaries(big, overlap, mode=mode, window=bwin)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            chunking.blend_rank_boundaries(big, overlap, mode=mode, window=bwin)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("rank %d mode %-4s parity %s moved %d B; 1080p halo blend %.3f ms" % (rank, mode, good, moved, ms), flush=True)
        if window is not None:
            window.close()
        if bwin is not None:
            bwin.close()
        dist.barrier()
    t_ok = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU CHECK", "PASS" if int(t_ok.item()) else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t_ok.item()) else 1)


if __name__ == "__main__":
    main()

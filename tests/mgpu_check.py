#!/usr/bin/env python
"""Multi-GPU check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Verifies against the oracle, on real GPUs:
  1. the rank-boundary halo blend, NCCL send/recv and CUDA-IPC peer reads with the device-side handshake, with a
     DIFFERENT number of frames on every rank, over several epochs enqueued without any host synchronisation;
  2. BASELINE config 4 as a system (chunk plan -> shards -> per-rank stitch -> halo blend): the ranks' owned frames
     equal the single-GPU stitch byte for byte;
and times both halo modes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import chunk_blend as ocb  # noqa: E402
from videovanish_b200 import chunking, ops  # noqa: E402


def expected_after_blend(clips, rank, overlap):
    world, half = len(clips), overlap // 2
    t = len(clips[rank])
    expect = clips[rank].copy()
    if rank < world - 1:
        full = ocb.blend_overlap(clips[rank][t - overlap:], clips[rank + 1][:overlap])
        expect[t - overlap:t - overlap + half] = full[:half]
    if rank > 0:
        tp = len(clips[rank - 1])
        full = ocb.blend_overlap(clips[rank - 1][tp - overlap:], clips[rank][:overlap])
        expect[half:overlap] = full[half:]
    return expect


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h, w, overlap = 270, 480, 16
    frames_of = [40 + 8 * ((3 * r) % 5) for r in range(world)]                    # uneven T per rank
    clips = [np.random.default_rng(100 + r).integers(0, 256, (frames_of[r], h, w, 3), dtype=np.uint8) for r in range(world)]
    expect = expected_after_blend(clips, rank, overlap)
    ok = True
    for mode in ("nccl", "peer"):
        mine = torch.from_numpy(clips[rank]).to(dev)
        window = chunking.PeerWindow(mine) if mode == "peer" else None
        moved = chunking.blend_rank_boundaries(mine, overlap, mode=mode, window=window)
        torch.cuda.synchronize()
        good = np.array_equal(mine.cpu().numpy(), expect)
        if mode == "peer":
            # several more epochs back to back, no host synchronisation in between: the buffer is restored by a
            # device copy and blended again; the handshake alone keeps neighbours from reading half-written frames
            src = torch.from_numpy(clips[rank]).to(dev)
            for _ in range(5):
                mine.copy_(src)
                chunking.blend_rank_boundaries(mine, overlap, mode=mode, window=window)
            torch.cuda.synchronize()
            good = good and np.array_equal(mine.cpu().numpy(), expect) and not window.error()
        # the overlapped form: the block is "produced" (a device copy here) boundary frames first, the exchange runs
        # on a side stream underneath the production of the rest; same bytes, several epochs back to back
        src2 = torch.from_numpy(clips[rank]).to(dev)
        for _ in range(3):
            mine.fill_(0)
            chunking.produce_and_blend_boundaries(mine, overlap, lambda lo, hi: mine[lo:hi].copy_(src2[lo:hi]), mode=mode,
                                                  window=window)
        torch.cuda.synchronize()
        good_ov = np.array_equal(mine.cpu().numpy(), expect) and (window is None or not window.error())
        print("rank %d mode %-4s overlapped production + exchange parity %s" % (rank, mode, good_ov), flush=True)
        good = good and good_ov
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
        print("rank %d mode %-4s T=%d parity %s moved %d B; 1080p halo blend %.3f ms" % (rank, mode, frames_of[rank], good, moved, ms),
              flush=True)
        dist.barrier()
        if window is not None:
            window.close()
        if bwin is not None:
            bwin.close()
        dist.barrier()

    # ---- config 4 as a system, small frames: 150 frames, chunk 20 / overlap 6 over all ranks
    n_frames, chunk, ov, fh, fw = 30 * max(world, 3) + 7, 20, 6, 72, 128

    def process_chunk(ci, s, e):
        g = torch.Generator(device=dev)
        g.manual_seed(5000 + ci)
        inp = torch.randint(0, 256, (e - s, fh // 2, fw // 2, 3), dtype=torch.uint8, device=dev, generator=g)
        fr = torch.stack([torch.full((fh, fw, 3), (f * 7) % 256, dtype=torch.uint8, device=dev) for f in range(s, e)])
        mk = torch.zeros((e - s, fh, fw), dtype=torch.uint8, device=dev)
        mk[:, 20:50, 30:90] = 255
        return ops.upscale_feather_composite(inp, fr, mk, 3)

    for mode in ("nccl", "peer"):
        block, first, owned = chunking.run_sharded(n_frames, chunk, ov, process_chunk, mode=mode)
        plan = chunking.chunk_plan(n_frames, chunk, ov)
        whole = chunking.stitch_chunks([process_chunk(ci, s, e) for ci, (s, e) in enumerate(plan)], plan)
        lo, hi = first + owned.start, first + owned.stop
        good = torch.equal(block[owned], whole[lo:hi])
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        good = good and spans[0][0] == 0 and spans[-1][1] == n_frames and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        print("rank %d config-4 sharded (%s): frames [%d,%d) of %d equal the single-GPU stitch: %s" % (rank, mode, lo, hi, n_frames, good),
              flush=True)
        ok &= good
    t_ok = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU CHECK", "PASS" if int(t_ok.item()) else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t_ok.item()) else 1)


if __name__ == "__main__":
    main()

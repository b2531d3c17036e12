#!/usr/bin/env python
"""Host-side cost of enqueueing the hot path (no synchronisation inside the loop): if a step takes longer
to ENQUEUE than to run, the bench is launch-bound.  Usage: python tools/launch_cost.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videovanish_b200 import _lib, ops, synth  # noqa: E402

H0, W0, HS, WS, T = 1080, 1920, 540, 960, 300
dev = torch.device("cuda", 0)
fr = torch.from_numpy(np.tile(synth.frames(12, H0, W0, seed=1), (T // 12, 1, 1, 1))).to(dev)
mk = torch.from_numpy(np.tile(synth.masks(12, H0, W0, seed=3), (T // 12, 1, 1, 1))).to(dev)
inp = torch.from_numpy(np.tile(synth.noise_frames(12, HS, WS, seed=2), (T // 12, 1, 1, 1))).to(dev)
g = torch.Generator(device=dev).manual_seed(5)
ff = torch.randn((T - 1, HS, WS, 2), device=dev, generator=g) * 0.05 + torch.tensor([3.0, -1.5], device=dev)
fb = -ff + torch.randn((T - 1, HS, WS, 2), device=dev, generator=g) * 0.05
out = torch.empty_like(fr)
dil, low = ops.binarize_dilate(mk, 8, lowres_size=(HS, WS))
small = ops.resize(fr, HS, WS)


def measure(name, fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-34s enqueue %.3f ms/call   gpu %.3f ms/call" % (name, (t1 - t0) * 1e3 / n, e0.elapsed_time(e1) / n), flush=True)


measure("K1 binarize_dilate", lambda: ops.binarize_dilate(mk, 8, lowres_size=(HS, WS)))
measure("K2 resize", lambda: ops.resize(fr, HS, WS))
measure("K3 composite", lambda: ops.upscale_feather_composite(inp, fr, dil, 3, out=out))
for lean in (5, 0):
    _lib.set_option("k4_lean", lean)
    measure("K4 propagate (k4_lean=%d)" % lean, lambda: ops.propagate(small, low, ff, fb))
_lib.set_option("k4_lean", 5)

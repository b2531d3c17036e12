#!/usr/bin/env python
"""K4 at the BASELINE config-2 shape (300 frames, 960x540, windows 50+10+10) over the step-chain options:
python tools/k4_sweep.py [--frames N].  One JSON object per line; every variant is checked against the first."""
import argparse
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videovanish_b200 import _lib, ops, synth  # noqa: E402

H0, W0, HS, WS = 1080, 1920, 540, 960


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--box", action="store_true", help="moving-box mask only (no salt)")
    args = ap.parse_args()
    t = args.frames
    dev = torch.device("cuda", 0)
    fr = torch.from_numpy(np.tile(synth.frames(16, HS, WS, seed=1), ((t + 15) // 16, 1, 1, 1))[:t]).to(dev)
    mk = torch.from_numpy(synth.masks(t, H0, W0, seed=3, salt=0.0 if args.box else 0.001)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    ff = torch.tensor([3.0, -1.5], device=dev) + 0.05 * torch.randn((t - 1, HS, WS, 2), device=dev, generator=g)
    fb = -ff + 0.05 * torch.randn((t - 1, HS, WS, 2), device=dev, generator=g)
    bad = torch.rand((t - 1, HS, WS), device=dev, generator=g) < 0.02
    ff[bad] += (torch.rand((int(bad.sum()), 2), device=dev, generator=g) - 0.5) * 40.0
    _, low = ops.binarize_dilate(mk, 8, lowres_size=(HS, WS))
    del mk
    pbuf = torch.empty((t, HS, WS), dtype=torch.int32, device=dev)
    peak = 6545.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    alg = t * 56 * HS * WS
    ref = None
    for streams, ctas, lean, pre in itertools.chain(
            [(1, 5, 5, 0)], itertools.product((2, 3, 4), (7, 8, 9, 10), (5,), (0,)), [(2, 8, 5, 1), (2, 10, 6, 0)]):
        for k, v in dict(k4_streams=streams, k4_chain_ctas=ctas, k4_lean=lean, k4_precheck=pre).items():
            _lib.set_option(k, v)
        med, best = timeit(lambda: ops.propagate(fr, low, ff, fb, out=pbuf), reps=7, warm=2)
        same = True
        if ref is None:
            ref = pbuf.clone()
        else:
            same = bool(torch.equal(ref, pbuf))
        print(json.dumps(dict(k4_streams=streams, k4_chain_ctas=ctas, k4_lean=lean, k4_precheck=pre, ms=round(med, 4),
                              ms_best=round(best, 4), frac=round(alg / (med * 1e-3) / 1e9 / peak, 3), same=same)), flush=True)


if __name__ == "__main__":
    main()

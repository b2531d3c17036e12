#!/usr/bin/env python
"""Per-operator timing at the BASELINE config-2 shape (CUDA events, device-resident inputs larger
than L2).  Used for A/B-ing kernel variants in one GPU call:  python tools/microbench.py [--frames N]
Prints one JSON object per line."""
import argparse
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
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--ops", default="all")
    args = ap.parse_args()
    t = args.frames
    dev = torch.device("cuda", 0)
    base = min(t, 16)
    reps = (t + base - 1) // base
    fr = torch.from_numpy(np.tile(synth.frames(base, H0, W0, seed=1), (reps, 1, 1, 1))[:t]).to(dev)
    inp = torch.from_numpy(np.tile(synth.noise_frames(base, HS, WS, seed=2), (reps, 1, 1, 1))[:t]).to(dev)
    mk = torch.from_numpy(synth.masks(t, H0, W0, seed=3)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    ff = torch.tensor([3.0, -1.5], device=dev) + 0.05 * torch.randn((t - 1, HS, WS, 2), device=dev, generator=g)
    fb = -ff + 0.05 * torch.randn((t - 1, HS, WS, 2), device=dev, generator=g)
    bad = torch.rand((t - 1, HS, WS), device=dev, generator=g) < 0.02
    ff[bad] += (torch.rand((int(bad.sum()), 2), device=dev, generator=g) - 0.5) * 40.0
    px, spx = H0 * W0, HS * WS
    dil, low = ops.binarize_dilate(mk, 8, lowres_size=(HS, WS))
    small = ops.resize(fr, HS, WS)
    out = torch.empty_like(fr)
    empty_mask = torch.zeros_like(dil)
    full_mask = torch.full_like(dil, 255)
    peak = 6545.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    wanted = None if args.ops == "all" else [w.strip().lower() for w in args.ops.split(",")]

    def report(name, alg_bytes, fn, **extra):
        if wanted is not None and not any(w in name.lower() for w in wanted):
            return
        med, best = timeit(fn)
        gbs = alg_bytes / (med * 1e-3) / 1e9
        print(json.dumps(dict(op=name, frames=t, ms=med, ms_best=best, us_per_frame=med * 1e3 / t, GBps=gbs,
                              frac=gbs / peak, **extra)), flush=True)

    print(json.dumps(dict(mask_fraction=float((dil > 0).float().mean()))), flush=True)
    for exact in (1, 0):
        _lib.set_option("k1b_exact", exact)
        report("K1 dilate8 + half-res mask", t * (4 * px + spx), lambda: ops.binarize_dilate(mk, 8, lowres_size=(HS, WS)),
               k1b_exact=exact)
        report("K1 dilate8", t * 4 * px, lambda: ops.binarize_dilate(mk, 8), k1b_exact=exact)
    _lib.set_option("k1b_exact", 1)
    report("K1 dilate8 + generic low-res (536)", t * (4 * px + 536 * 960),
           lambda: ops.binarize_dilate(mk, 8, lowres_size=(536, 960)))
    for diag in (2, 1):
        _lib.set_option("k1b_diag", diag)
        report("K1 dilate8 + half-res mask", t * (4 * px + spx), lambda: ops.binarize_dilate(mk, 8, lowres_size=(HS, WS)), k1b_diag=diag)
    for diag in (1, 0):
        _lib.set_option("k1b_diag", diag)
        for n in (12, 16, 25):
            report("K1 dilate%d" % n, t * 4 * px, lambda: ops.binarize_dilate(mk, n), k1b_diag=diag)
    _lib.set_option("k1b_diag", 2)
    report("K2 resize 1080p->540p", t * (3 * px + 3 * spx), lambda: ops.resize(fr, HS, WS))
    report("K2 resize 1080p->536p", t * (3 * px + 3 * 536 * 960), lambda: ops.resize(fr, 536, 960))
    report("K2 nearest mask 1080p->540p", t * (px + spx), lambda: ops.resize(dil, HS, WS, ops.INTER_NEAREST))
    # ---- K3: the k3_fast kernel (default) with / without K1's bit plane and per rows-per-task, against the
    # round-1 kernels, on four masks; and the production 960x536 geometry
    dil_b, _, bits = ops.binarize_dilate(mk, 8, return_bits=True)
    full_bits = torch.full_like(bits, -1)
    box_dil, _, box_bits = ops.binarize_dilate(torch.from_numpy(synth.masks(t, H0, W0, seed=3, salt=0.0)).to(dev), 8, return_bits=True)
    k3b = t * (7 * px + 3 * spx)
    K3 = lambda m, b=None, i=None: (lambda: ops.upscale_feather_composite(inp if i is None else i, fr, m, 3, out=out, mask_bits=b))
    empty_bits = torch.zeros_like(bits)
    inp536 = inp[:, :536].contiguous()
    k3b536 = t * (7 * px + 3 * 536 * 960)
    for x2 in (3, 2):                                       # 3 = k3_fastw (word tasks, default), 2 = k3_fast (rolling 16-pixel tasks)
        _lib.set_option("k3_x2", x2)
        report("K3 composite (synthetic mask)", k3b, K3(dil, bits), k3_x2=x2, bits=1)
        report("K3 composite (synthetic mask)", k3b, K3(dil), k3_x2=x2, bits=0)
        report("K3 composite (full mask)", k3b, K3(full_mask, full_bits), k3_x2=x2, bits=1)
        report("K3 composite (empty mask)", k3b, K3(empty_mask, empty_bits), k3_x2=x2, bits=1)
        report("K3 composite (box mask)", k3b, K3(box_dil, box_bits), k3_x2=x2, bits=1)
        report("K3 composite 960x536 (synthetic mask)", k3b536, K3(dil, bits, inp536), k3_x2=x2, bits=1)
    _lib.set_option("k3_x2", 3)
    for thr, rows_list in ((384, (16, 8, 6)), (512, (12, 14)), (256, (16, 7))):
        _lib.set_option("k3_tma_threads", thr)
        for rows in rows_list:
            _lib.set_option("k3_tma_rows", rows)
            report("K3 composite (synthetic mask)", k3b, K3(dil, bits), k3_x2=3, bits=1, rows=rows, threads=thr)
            if thr == 384:
                report("K3 composite 960x536 (synthetic mask)", k3b536, K3(dil, bits, inp536), k3_x2=3, bits=1, rows=rows, threads=thr)
                report("K3 composite (box mask)", k3b, K3(box_dil, box_bits), k3_x2=3, bits=1, rows=rows, threads=thr)
    _lib.set_option("k3_tma_rows", 16)
    _lib.set_option("k3_tma_threads", 512)
    for x2 in (1, 0):                                       # round-1 kernels
        _lib.set_option("k3_x2", x2)
        report("K3 composite (synthetic mask)", k3b, K3(dil), k3_x2=x2, bits=0)
    report("K3 composite 960x536 (synthetic mask)", k3b536, K3(dil, None, inp536), k3_x2=0, bits=0)
    _lib.set_option("k3_x2", 3)
    report("K3 composite feather 5 (generic path)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 5, out=out))
    report("K3 composite feather 8 (generic path)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 8, out=out))
    _lib.set_option("k3_big_from", 3)
    report("K3 composite feather 5 (k3_bigfeather)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 5, out=out))
    report("K3 composite feather 8 (k3_bigfeather)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 8, out=out))
    _lib.set_option("k3_big_from", 8)
    report("K3 composite feather 16 (k3_bigfeather)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 16, out=out))
    report("K3 composite feather 32 (k3_bigfeather)", k3b, lambda: ops.upscale_feather_composite(inp, fr, dil, 32, out=out))
    report("K3 composite feather 16, box mask (k3_bigfeather)", k3b, lambda: ops.upscale_feather_composite(inp, fr, box_dil, 16, out=out))
    del dil_b, full_bits, box_dil, box_bits, inp536, empty_bits
    # ---- K4: step-kernel variants
    pbuf = torch.empty((t, HS, WS), dtype=torch.int32, device=dev)
    k4b = t * 56 * spx
    for persist, sc in ((1, 5), (1, 4), (0, 5)):
        _lib.set_option("k4_persist", persist)
        _lib.set_option("k4_step_ctas", sc)
        report("K4 propagate 50+10 windows", k4b, lambda: ops.propagate(small, low, ff, fb, out=pbuf), k4_persist=persist, k4_step_ctas=sc)
    _lib.set_option("k4_persist", 0)
    for pre, lean, npt, sc, spec in ((1, 5, 1, 5, 1), (0, 5, 1, 5, 1), (1, 6, 1, 6, 1), (1, 8, 1, 8, 1), (1, 8, 1, 6, 1), (1, 5, 1, 5, 0),
                                     (1, 4, 2, 4, 1), (1, 3, 2, 3, 1), (1, 6, 1, 8, 1)):
        for k, v in dict(k4_precheck=pre, k4_lean=lean, k4_npt=npt, k4_step_ctas=sc, k4_speculate=spec).items():
            _lib.set_option(k, v)
        report("K4 propagate 50+10 windows", k4b, lambda: ops.propagate(small, low, ff, fb, out=pbuf), k4_precheck=pre, k4_lean=lean,
               k4_npt=npt, k4_step_ctas=sc, k4_speculate=spec)
    for k, v in dict(k4_precheck=0, k4_lean=5, k4_npt=1, k4_step_ctas=5, k4_speculate=1, k4_persist=0).items():
        _lib.set_option(k, v)
    report("K4 propagate 50+10 windows (fresh output tensor)", k4b, lambda: ops.propagate(small, low, ff, fb))
    del pbuf
    report("K5 chunk blend 16 frames", 16 * 9 * px, lambda: ops.chunk_blend(fr[:16], fr[16:32], out=out[:16]))
    report("copy (torch) 1080p frames", t * 6 * px, lambda: out.copy_(fr))
    # next rows: N3 painter (3 objects at inference resolution painted onto the 1080p canvas), N2 state -> float
    if wanted is None or any(w_ in ("n3", "n2", "next") for w_ in wanted):
        tn = min(t, 60)
        logits = torch.randn((tn, 3, HS, WS), device=dev, generator=g) - 1.0
        report("N3 paint 3 objects 540p -> 1080p", tn * (3 * 4 * spx + 3 * px),
               lambda: ops.paint_masks(logits, [(55, 255, 208), (148, 255, 55), (182, 255, 55)], out_size=(H0, W0)), frames_n=tn)
        # N4 at inference resolution: read_mask (erode + 0 / 4 dilations) and the blurred compose
        for nd in (0, 4):
            report("N4 wrapper mask (erode + %d dilate) 540p" % nd, t * 2 * spx, lambda: ops.wrapper_mask(low, nd))
        wm = ops.wrapper_mask(low, 0)
        comp_out = torch.empty_like(small)
        report("N4 wrapper compose (blurred) 540p", t * 10 * spx, lambda: ops.wrapper_compose(inp, small, wm, True, out=comp_out))
        box = torch.zeros_like(wm)
        box[:, 180:360, 300:620] = 255                    # one object, 11 % of the frame: the realistic case
        report("N4 wrapper compose (blurred, one object) 540p", t * 10 * spx,
               lambda: ops.wrapper_compose(inp, small, box, True, out=comp_out))
        report("N4 wrapper compose (hard) 540p", t * 10 * spx, lambda: ops.wrapper_compose(inp, small, wm, False, out=comp_out))
        packed = ops.propagate(small[:tn], low[:tn], ff[:tn - 1], fb[:tn - 1])
        report("N2 state -> float CHW", tn * (4 + 16) * spx, lambda: ops.propagate_to_float(packed), frames_n=tn)
        upd, _ = ops.propagate_to_float(packed)
        comp = torch.empty_like(small[:11])
        report("N2 neighbour merge (11-frame window)", 11 * (12 + 1 + 3 + 3 + 3) * spx,
               lambda: ops.neighbor_merge(upd[:11], low[:11], small[:11], comp, [False] * 11), frames_n=11)
        report("N4 masked frames 540p", t * 7 * spx, lambda: ops.apply_mask(small, wm, out=comp_out))
        report("N1 channel swap 1080p", t * 6 * px, lambda: ops.swap_rb(fr, out=out))

    # BASELINE config 4 shape: 4K frames, inference 960x536 (x ratio exactly 4, y ratio 4.03)
    if wanted is None or any("4k" in w_ for w_ in wanted):
        del fr, inp, mk, dil, low, small, out, empty_mask, full_mask, ff, fb
        torch.cuda.empty_cache()
        t4, H4, W4, h4, w4 = 48, 2160, 3840, 536, 960
        fr4 = torch.from_numpy(np.tile(synth.frames(4, H4, W4, seed=5), (12, 1, 1, 1))).to(dev)
        mk4 = torch.from_numpy(synth.masks(t4, H4, W4, seed=6)).to(dev)
        inp4 = torch.from_numpy(np.tile(synth.noise_frames(4, h4, w4, seed=7), (12, 1, 1, 1))).to(dev)
        px4, spx4 = H4 * W4, h4 * w4
        dil4, _, bits4 = ops.binarize_dilate(mk4, 8, return_bits=True)
        out4 = torch.empty_like(fr4)
        report("4K K1 dilate8 + low-res 960x536", t4 * (4 * px4 + spx4), lambda: ops.binarize_dilate(mk4, 8, lowres_size=(h4, w4)), frames_4k=t4)
        report("4K K2 resize ->960x536", t4 * (3 * px4 + 3 * spx4), lambda: ops.resize(fr4, h4, w4), frames_4k=t4)
        report("4K K3 composite (synthetic mask)", t4 * (7 * px4 + 3 * spx4),
               lambda: ops.upscale_feather_composite(inp4, fr4, dil4, 3, out=out4, mask_bits=bits4), frames_4k=t4)
        report("4K K5 chunk blend 16 frames", 16 * 9 * px4, lambda: ops.chunk_blend(fr4[:16], fr4[16:32], out=out4[:16]), frames_4k=t4)


if __name__ == "__main__":
    main()

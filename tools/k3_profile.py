#!/usr/bin/env python
"""K3 launches for ncu: each (variant, mask) once, in a fixed order (see the printed index)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videovanish_b200 import _lib, ops, synth  # noqa: E402

H0, W0, HS, WS, T = 1080, 1920, 540, 960, 60
dev = torch.device("cuda", 0)
fr = torch.from_numpy(np.tile(synth.frames(12, H0, W0, seed=1), (5, 1, 1, 1))).to(dev)
inp = torch.from_numpy(np.tile(synth.noise_frames(12, HS, WS, seed=2), (5, 1, 1, 1))).to(dev)
dil = ops.binarize_dilate(torch.from_numpy(synth.masks(T, H0, W0, seed=3)).to(dev), 8)
masks = {"synthetic": dil, "empty": torch.zeros_like(dil), "full": torch.full_like(dil, 255)}
out = torch.empty_like(fr)
i = 0
for tma, nt in ((1, 1), (0, 2)):
    _lib.set_option("k3_tma", tma)
    _lib.set_option("k3_nt", nt)
    for name, m in masks.items():
        ops.upscale_feather_composite(inp, fr, m, 3, out=out)
        torch.cuda.synchronize()
        print("k3 launch %d: tma=%d nt=%d mask=%s" % (i, tma, nt, name))
        i += 1

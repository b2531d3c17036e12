#!/usr/bin/env python
"""K3 launches for ncu: the default kernel (k3_fastw) on the bench mask with K1's bit plane, 60 frames 1080p <- 960x540,
then the same on 960x536.  ncu -k regex:k3_fast -s 1 -c 2 captures launch 1 (540) and 2 (536); launch 0 is warm-up."""
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
dil, _, bits = ops.binarize_dilate(torch.from_numpy(synth.masks(T, H0, W0, seed=3)).to(dev), 8, return_bits=True)
out = torch.empty_like(fr)
inp536 = inp[:, :536].contiguous()
if len(sys.argv) > 1:
    _lib.set_option("k3_x2", int(sys.argv[1]))
for i, src in enumerate((inp, inp, inp536)):
    ops.upscale_feather_composite(src, fr, dil, 3, out=out, mask_bits=bits)
    torch.cuda.synchronize()
    print("k3 launch %d: %s" % (i, tuple(src.shape)))

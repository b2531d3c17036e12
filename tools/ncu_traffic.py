#!/usr/bin/env python
"""Aggregate an ncu CSV (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
one bench.py step) into per-stage DRAM traffic -> profiles/ncu_traffic.json.
Usage: python tools/ncu_traffic.py gpurun_out/step_metrics.csv FRAMES"""
import collections
import csv
import json
import os
import sys

STAGE = {"k1a": "K1_binarize_dilate", "k1b": "K1_binarize_dilate", "k1c": "K1_binarize_dilate", "k1d": "K1_binarize_dilate",
         "k2_resize": "K2_resize_down", "k2_make": None, "k3_": "K3_upscale_feather_composite", "k4_": "K4_propagate",
         "k5_": "K5_halo_blend"}


def main():
    path, frames = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: {"dram_bytes": 0.0, "time_us": 0.0, "launches": 0})
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
             "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki]
        stage = None
        for k, v in STAGE.items():
            if k in name:
                stage = v
                break
        if stage is None:
            continue
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
        if r[mi].startswith("dram__bytes"):
            agg[stage]["dram_bytes"] += v
        elif r[mi].startswith("gpu__time_duration"):
            agg[stage]["time_us"] += v
            agg[stage]["launches"] += 1
    out = {"frames": frames, "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                        "--clock-control none python bench.py --steps 1 --warmup 3 (last step's launches)",
           "stages": agg}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

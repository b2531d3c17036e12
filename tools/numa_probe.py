#!/usr/bin/env python
"""Host topology and CONCURRENT pinned host<->device bandwidth of all ranks (torchrun, one rank per GPU).
Answers whether the end-to-end path's multi-GPU scaling is bound by the platform (PCIe root complexes, NUMA placement)
or by this repo's host pipeline.  python -m torch.distributed.run --nproc-per-node N tools/numa_probe.py"""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def node_of_cpu(cpu):
    for p in glob.glob("/sys/devices/system/node/node*/cpu%d" % cpu):
        return int(p.split("/node/node")[1].split("/")[0])
    return -1


def set_mempolicy_preferred(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (x86_64: 238); True on success."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    r = libc.syscall(238, 1, ctypes.byref(mask), 65)
    return r == 0


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = {"rank": rank}
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = sorted(64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        info["pci"] = bus
        try:
            info["gpu_numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read())
        except Exception as e:
            info["gpu_numa_node"] = str(e)[:60]
        info["ideal_cpus"] = "%d..%d (%d)" % (ideal[0], ideal[-1], len(ideal)) if ideal else "none"
    except Exception as e:
        info["nvml"] = str(e)[:80]
        ideal = []
    allowed = sorted(os.sched_getaffinity(0))
    info["allowed_cpus"] = "%d..%d (%d)" % (allowed[0], allowed[-1], len(allowed))
    info["allowed_nodes"] = sorted({node_of_cpu(c) for c in allowed})
    info["ideal_and_allowed"] = len(set(ideal) & set(allowed))
    if rank == 0:
        for cmd in ("nvidia-smi topo -m", "lscpu | grep -i -E 'numa|socket|model name'", "cat /sys/fs/cgroup/cpuset.mems.effective",
                    "cat /sys/fs/cgroup/cpuset.cpus.effective", "free -g | head -2"):
            try:
                print("$ " + cmd + "\n" + subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout, flush=True)
            except Exception as e:
                print(cmd, "failed", e)

    n = 1 << 30

    def bandwidth(tag):
        h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h1.fill_(1), h2.fill_(2)
        d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
        d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
        s2 = torch.cuda.Stream()

        def timed(fn, reps=4):
            fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return n * reps / (time.perf_counter() - t0) / 1e9
        res = {"h2d": timed(lambda: d1.copy_(h1, non_blocking=True)), "d2h": timed(lambda: h2.copy_(d2, non_blocking=True))}

        def both():
            d1.copy_(h1, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        res["duplex_each"] = timed(both)
        info[tag] = {k: round(v, 1) for k, v in res.items()}
        del h1, h2, d1, d2

    bandwidth("default")
    node = info.get("gpu_numa_node")
    if isinstance(node, int) and node >= 0:
        info["mempolicy_set"] = set_mempolicy_preferred(node)
        bandwidth("mem_on_gpu_node")
    if ideal and set(ideal) & set(allowed):
        os.sched_setaffinity(0, set(ideal) & set(allowed))
        bandwidth("cpu_and_mem_on_gpu_node")
    if world > 1:
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)
    else:
        allinfo = [info]
    if rank == 0:
        for i in allinfo:
            print(json.dumps(i), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

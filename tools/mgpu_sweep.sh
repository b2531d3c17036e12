#!/bin/bash
# Multi-GPU check on N GPUs of one box: byte parity (tests/mgpu_check.py), then bench.py with the halo exchange serial
# (VV_HALO_OVERLAP=0) and underneath K3 (default), optionally with other CTA budgets.  Usage: bash tools/mgpu_sweep.sh N [cfg ...]
P='import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],3), {k[:2]:round(v["ms"],3) for k,v in d["stages"].items()}, d["halo_parity"], d["c4_sharded"]["byte_exact"], d["halo_handshake_error"], "e2e", round(d["e2e"]["value"]))'
N=$1; shift
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/mgpu_check_${N}gpu.log 2>&1; grep "MGPU" gpurun_out/mgpu_check_${N}gpu.log
port=29540
for cfg in "VV_HALO_OVERLAP=0" "VV_HALO_OVERLAP=1" "$@"; do
  port=$((port+1))
  echo "== $cfg"
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${N}gpu_$cfg.json
  python -c "$P" < gpurun_out/bench_${N}gpu_$cfg.json
done

#!/bin/bash
# One gpurun call: parity tests, smoke, bench, microbench, ncu launch list + full captures.
# Usage (from the repo root on the GPU box):  bash tools/gpu_check.sh "tests bench micro ncu"
set -u
what=${1:-"tests bench micro ncu"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
if [[ $what == *tests* ]]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -40 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log
fi
if [[ $what == *micro* ]]; then
  timeout 900 python tools/microbench.py > gpurun_out/micro.log 2>&1
  cat gpurun_out/micro.log
fi
if [[ $what == *bench* ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
  echo "bench exit $?" >> gpurun_out/bench.log
  tail -5 gpurun_out/bench.log
fi
if [[ $what == *refarm* ]]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.log 2>&1
  tail -2 gpurun_out/bench_reference.log
fi
if [[ $what == *ncu* ]]; then
  B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
  # launch list + DRAM traffic of exactly one step (283 launches per step with two K4 step chains; skip the 3 warm-up steps): cold-cache,
  # serialised under the profiler -> compare SHARES with the bench's stage times, not absolutes
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k 'regex:^k[1-5]' --launch-skip 849 --launch-count 283 --csv --log-file gpurun_out/step_metrics.csv \
      $B > gpurun_out/bench_under_ncu.log 2>&1
  # full captures of the hot kernels inside the same command (4th step = after warm-up)
  for spec in "k3_fastw:3" "k1a_binarize:3" "k1b_dilate:3" "k2_resize_linear_ratio:3" "k4_pack:3" "k4_step_lean:880" "k4_step_lean:1060"; do
    k=${spec%%:*}; skip=${spec##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f \
        -o gpurun_out/prof_${k}_s$skip $B > gpurun_out/ncu_${k}_s$skip.log 2>&1
  done
  ls -la gpurun_out
fi

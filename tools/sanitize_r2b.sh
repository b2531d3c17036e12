# compute-sanitizer over the kernels / host paths changed last in round 2: K1 diagonal blocks, K4 step chains on two streams,
# masks of another size through the drop-in.  Usage (GPU box): bash tools/sanitize_r2b.sh
K='k1 or k4 or golden or dropin and not 540'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k \"$K\""
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "$K" 2>&1 | grep -v "^=========\s*$" | tail -4
done

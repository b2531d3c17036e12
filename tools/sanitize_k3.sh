for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python -m pytest tests/test_gpu_round2.py -m gpu -q -k 'k3_fast and (120-176-60-88 or 120-176-56-88 or 16-16-8-8 or 40-64-10-16)'"
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k 'k3_fast and (120-176-60-88 or 120-176-56-88 or 16-16-8-8 or 40-64-10-16)' 2>&1 | grep -v "^=========\s*$" | tail -6
done
echo "== compute-sanitizer --tool memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== compute-sanitizer --tool racecheck smoke"; timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

# compute-sanitizer over smoke() and the small parity tests: memcheck, racecheck, synccheck, initcheck.
# usage: bash scripts/gpu_sanitize.sh <tag>
TAG=$1
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_round2.py::test_verdict_fused_ragged_tiles_and_edge_cases tests/test_gpu_round2.py::test_compaction_of_bit_packed_survivors tests/test_gpu_round2.py::test_device_side_segment_source tests/test_gpu_parity.py::test_segcheck_f64_edge_cases tests/test_gpu_generate.py -k 'not ks and not soak'"
for tool in memcheck racecheck synccheck; do
  echo "== $tool: smoke()" 
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_${TAG}_${tool}_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${TAG}_${tool}_smoke.out 2>&1
  echo "rc=$?"; tail -3 gpurun_out/sanitize_${TAG}_${tool}_smoke.log; tail -1 gpurun_out/sanitize_${TAG}_${tool}_smoke.out
done
for tool in memcheck racecheck; do
  echo "== $tool: round-2 parity tests (small)"
  timeout 2400 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_${TAG}_${tool}_tests.log python -m pytest -x -q tests/test_gpu_round2.py -k "ragged or compaction or segment_source or dda_on" > gpurun_out/sanitize_${TAG}_${tool}_tests.out 2>&1
  echo "rc=$?"; tail -3 gpurun_out/sanitize_${TAG}_${tool}_tests.log; tail -2 gpurun_out/sanitize_${TAG}_${tool}_tests.out
done

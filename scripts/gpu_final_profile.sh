# usage: bash scripts/gpu_final_profile.sh <tag> <a|b>
#   a: GPU tests + full bench + reference arm + ncu launch list      b: ncu --set full of every hot kernel
# (two gpurun calls: one profiler run per call)
TAG=$1; STAGE=${2:-a}
mkdir -p gpurun_out
set -x
if [ "$STAGE" = a ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python scripts/show_bench.py gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
else
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:generate_kernel|segcheck_kernel|dda_kernel|gmm_sample_kernel" -s 15 -c 5 -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
fi

# usage: bash scripts/gpu_final_profile.sh <tag> <a|b>
#   a: GPU tests + full bench + reference arm + ncu launch list      b: ncu --set full of every hot kernel
# (two gpurun calls: one profiler run per call)
TAG=$1; STAGE=${2:-a}
mkdir -p gpurun_out
set -x
if [ "$STAGE" = a ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 2 --warmup 3 --passes 2 --no-e2e --no-cpu --no-secondary --no-config4 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --passes 2 --no-e2e --no-cpu --no-secondary --no-config4 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
else
python bench.py --steps 2 --warmup 3 --passes 2 --no-e2e --no-cpu --no-secondary --no-config4 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:generate_kernel|verdict_kernel|dda_kernel|gmm_sample_kernel|compact_bits" -s 40 -c 7 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --passes 2 --no-e2e --no-cpu --no-secondary --no-config4 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
fi

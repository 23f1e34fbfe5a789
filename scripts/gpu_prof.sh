# usage: bash scripts/gpu_prof.sh <tag> <kernel-regex> [skip] [count]
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=${3:-6}; CNT=${4:-3}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python scripts/show_bench.py gpurun_out/bench_quick.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log

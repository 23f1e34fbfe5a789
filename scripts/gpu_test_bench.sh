mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python scripts/show_bench.py gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err

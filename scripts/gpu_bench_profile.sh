mkdir -p gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 3000 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_a.json 2>&1; tail -c 1500 gpurun_out/bench_ref_a.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 20 -c 5 -o gpurun_out/prof_r1_b python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out

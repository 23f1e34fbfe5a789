mkdir -p gpurun_out
for dp in 1 2 3; do
python bench.py --steps 3 --warmup 3 --passes 8 --no-cpu --no-secondary --no-config4 --e2e-depth $dp > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; echo depth=$dp rc=$?; tail -2 gpurun_out/bench_dev.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dev.json'))
for k in ('e2e','e2e_generator_mode'):
    e=d[k]; print(k, '%.3e seg/s  %.2f ms/pass h2d %d d2h %d'%(e['value'],e['ms_per_pass'],e['h2d_bytes_per_pass'],e['d2h_bytes_per_pass']))
PY
done

mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; tail -3 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_latest.json'))
print('value %.3e seg/s  ms/step %.3f  valid/s %.3e  e2e %.3e (%.2f ms)' % (d['value'], d['ms_per_step'], d['valid_paths_per_s'], d['e2e']['value'], d['e2e']['ms_per_step']))
for k,v in d['kernels'].items(): print('  %-14s %.3f ms  share %.2f  %.0f GB/s  frac %.3f' % (k, v['ms'], v['share'], v['achieved_gbs'], v['frac']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'clocks', d['clocks'])
PY

mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary --no-config4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pass %.4f'%d['ms_per_pass'], {k:(round(d[k]['ms_per_pass'],3), round(d[k]['raw_copy_ms_per_pass'],3)) for k in ('e2e','e2e_generator_mode')})"
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv

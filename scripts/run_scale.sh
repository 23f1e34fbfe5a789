# usage: bash scripts/run_scale.sh N tag
N=$1; TAG=$2
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
fi
echo rc=$?; tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_${N}gpu.json'))
print('N=%d value %.3e ms/pass %.4f valid/s %.3e'%(d['n_gpus'],d['value'],d['ms_per_pass'],d['valid_paths_per_s']))
for k in ('e2e','e2e_generator_mode'):
    e=d[k]; print(k, '%.3e seg/s  %.2f ms/pass valid/s %.3e'%(e['value'],e['ms_per_pass'],e['valid_paths_per_s']))
c=d['config4']; print('config4 %.4f s %.3e maps/s digest %s counts %s'%(c['seconds'],c['maps_per_s'],c['digest'],c['counts_all_gathered']))
print(d['counts_all_gathered'], d['clocks'])
PY

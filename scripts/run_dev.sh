mkdir -p gpurun_out
for lib in "" $(ls ppnet_b200/lib/libvar_*.so 2>/dev/null); do
  if [ -n "$lib" ]; then export PPNET_B200_LIB=$PWD/$lib; else unset PPNET_B200_LIB; fi
  echo "== variant ${lib:-default}"; python bench.py --steps 3 --warmup 3 --passes 8 --no-e2e --no-cpu --no-config4 --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pass %.4f seq %.4f'%(d['ms_per_pass'], d['run_info']['sequential_pass_ms']), {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
done

for v in "" 9 10 12; do
  if [ -n "$v" ]; then export PPNET_B200_LIB=$PWD/ppnet_b200/lib/libvar_gen_$v.so; else unset PPNET_B200_LIB; fi
  echo "== variant $v"; python bench.py --steps 3 --warmup 3 --passes 8 --no-e2e --no-cpu --no-secondary --no-config4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pass %.4f'%d['ms_per_pass'], {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
done

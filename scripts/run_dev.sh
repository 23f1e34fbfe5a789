mkdir -p gpurun_out
for v in "" 6 7 8; do
  if [ -n "$v" ]; then export PPNET_B200_LIB=$PWD/ppnet_b200/lib/libvar_$v.so; fi
  echo "== variant $v"; python scripts/dev_verdict.py 2>&1 | grep -E "fused|old|mismatches [1-9]"
done | tee gpurun_out/dev_variants.log

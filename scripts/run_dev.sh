mkdir -p gpurun_out
python scripts/dev_verdict.py 2>&1 | grep -E "fused|old|mismatches [1-9]"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q 2>&1 | tail -3

mkdir -p gpurun_out
python scripts/dev_verdict.py 2>&1 | grep -E "old|mismatches [1-9]"
export PPNET_NEW_SEGCHECK=1
python scripts/dev_verdict.py 2>&1 | grep -E "old|mismatches [1-9]"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_round2.py -x -q 2>&1 | tail -4

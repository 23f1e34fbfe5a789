# compute-sanitizer is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.txt).  Substitute: the library built with
# PPNET_ASSERT bounds / protocol checks on every shared-memory queue, stage and scatter index (make debug), and the whole
# GPU parity suite + smoke() run against it.  A failed device assert aborts the process with file:line.
# usage: bash scripts/gpu_debug_bounds.sh <tag>
TAG=$1
mkdir -p gpurun_out
export PPNET_B200_LIB=$PWD/ppnet_b200/lib/libppnet_b200_dbg.so
{
echo "library: $PPNET_B200_LIB"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
} > gpurun_out/debug_bounds_$TAG.txt 2>&1
cat gpurun_out/debug_bounds_$TAG.txt

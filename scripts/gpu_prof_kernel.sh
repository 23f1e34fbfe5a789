# usage: bash scripts/gpu_prof_kernel.sh <tag> <which> <kernel-regex>
mkdir -p gpurun_out
TAG=$1; W=$2; RX=$3
python scripts/prof_kernel.py $W > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 1 -f -o gpurun_out/prof_$TAG python scripts/prof_kernel.py $W > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log

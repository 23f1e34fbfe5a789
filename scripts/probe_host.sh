# host / PCIe facts of the GPU box (no timing of our code)
mkdir -p gpurun_out
{
echo "== nvidia-smi topo -m"; nvidia-smi topo -m
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== lscpu"; lscpu | head -40
echo "== numa"; ls /sys/devices/system/node/ ; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
echo "== mem"; free -g
echo "== pci link"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist) $(readlink -f $d); fi; done
echo "== affinity"; python - <<'PY'
import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0)))
PY
} > gpurun_out/probe_host.txt 2>&1
tail -80 gpurun_out/probe_host.txt

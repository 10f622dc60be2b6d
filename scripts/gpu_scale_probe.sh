#!/bin/bash
# 8-GPU box: host topology, concurrent PCIe ceiling at 1/2/4/8 ranks, then the bench's end-to-end path at 8 and 4 ranks.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done; free -g | head -2) > gpurun_out/host_numa.txt 2>&1
timeout 400 python scripts/pcie_scaling.py --gpus $N > gpurun_out/pcie_scaling.json 2> gpurun_out/pcie_scaling.err; echo "pcie rc=$?"; tail -2 gpurun_out/pcie_scaling.err
for n in $N 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 2 --no-cpu --no-extras > gpurun_out/scale_e2e_$n.log 2> gpurun_out/scale_e2e_$n.err; echo "bench $n rc=$?"; tail -2 gpurun_out/scale_e2e_$n.err | cut -c1-300
done
cat gpurun_out/host_numa.txt; head -c 1500 gpurun_out/pcie_scaling.json; echo
python - <<P
import json
for n in ($N, 4):
    try:
        l=json.loads(open("gpurun_out/scale_e2e_%d.log"%n).read().strip().split("\n")[-1])
        print(n, "ranks: kernel %.3f ms value %.1f GMB/s | e2e %.2f ms %.2f GMB/s | binding %s" % (l["ms_per_step"], l["value"]/1e9, l["e2e"]["ms_per_step"], l["e2e"]["value"]/1e9, l["e2e"]["host_binding"]))
    except Exception as e: print(n, "failed", e)
P

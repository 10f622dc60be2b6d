#!/bin/bash
# GPU box: quick parity subset + kernel-only bench. Usage: gpu_bench.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<P
import json
l=json.loads(open("gpurun_out/bench_$tag.log").read().strip().split("\n")[-1])
print("$tag: kernel %.3f ms  value %.2f GMB/s  roofline %.3f  e2e %.2f GMB/s (%.2f ms)" % (l["ms_per_step"], l["value"]/1e9, l["roofline"]["frac"], l["e2e"]["value"]/1e9, l["e2e"]["ms_per_step"]))
P

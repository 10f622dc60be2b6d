#!/bin/bash
# GPU box: time a deliberately incomplete kernel build (no parity check). Usage: gpu_ablate.sh <tag>
tag=${1:-abl}; shift
mkdir -p gpurun_out
H264SCROLL_BENCH_NOCHECK=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<P
import json
l=json.loads(open("gpurun_out/bench_$tag.log").read().strip().split("\n")[-1])
print("$tag: kernel %.3f ms" % (l["ms_per_step"]))
P

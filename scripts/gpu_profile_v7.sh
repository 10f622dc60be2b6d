#!/bin/bash
# GPU box: plain run, then ncu launch list + one full capture of the compose kernel (small batch). Usage: gpu_profile_v7.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu --streams 32 --frames 64"
timeout 200 python bench.py $ARGS > gpurun_out/plain_$tag.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:compose_frames -s 2 -c 1 -o gpurun_out/prof_$tag python bench.py $ARGS > gpurun_out/ncu2_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log | cut -c1-300; ls -la gpurun_out/prof_$tag.ncu-rep

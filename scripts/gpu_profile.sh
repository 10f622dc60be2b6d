#!/bin/bash
# GPU box: plain run, then ncu launch list + full capture of the compose kernel (small batch). Usage: gpu_profile.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu --streams 32 --frames 64"
python bench.py $ARGS > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu1_$tag.log 2>&1
python bench.py $ARGS > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:compose_frames -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py $ARGS > gpurun_out/ncu2_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log; ls -la gpurun_out/

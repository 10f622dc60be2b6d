#!/bin/bash
# GPU box: plain run, ncu launch list, one full ncu capture of the compose kernel - all on the bench's own workload.
tag=${1:-prof}
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-cpu"
python bench.py $ARGS > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu1_$tag.log 2>&1
python bench.py $ARGS > gpurun_out/plain2_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:compose_frames -s 6 -c 1 -o gpurun_out/prof_$tag python bench.py $ARGS > gpurun_out/ncu2_$tag.log 2>&1
tail -c 400 gpurun_out/plain_$tag.log; ls -la gpurun_out/prof_$tag.ncu-rep

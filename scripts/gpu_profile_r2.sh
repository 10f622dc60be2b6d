#!/bin/bash
# GPU box: evidence for profiles/: launch list of the default bench command, one full ncu capture of the compose kernel on
# config 5 and one on config 4 (each only after the same command has exited 0 without ncu).
mkdir -p gpurun_out
A="--steps 2 --warmup 1 --no-cpu --no-extras --no-check"
python bench.py $A > gpurun_out/plain_r2_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py $A > gpurun_out/ncu_l.log 2>&1
python bench.py $A --streams 32 > gpurun_out/plain_r2_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:compose_frames -s 2 -c 1 -o gpurun_out/prof_r2_final python bench.py $A --streams 32 > gpurun_out/ncu_c5.log 2>&1
python bench.py $A --workload config4 > gpurun_out/plain_r2_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:compose_frames -s 2 -c 1 -o gpurun_out/prof_r2_config4 python bench.py $A --workload config4 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/prof_r2_config4.ncu-rep gpurun_out/r2_final_launches.csv

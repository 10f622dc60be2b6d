#!/bin/bash
# GPU box: DRAM bytes of ONE full-size compose launch (the bench's default workload) for profiles/traffic.json
mkdir -p gpurun_out
A="--steps 2 --warmup 1 --no-cpu --no-extras --no-check"
python bench.py $A > gpurun_out/plain_traffic.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:compose_frames -s 2 -c 1 --csv --log-file gpurun_out/r2_traffic.csv python bench.py $A > gpurun_out/ncu_traffic.log 2>&1
cat gpurun_out/r2_traffic.csv | tail -8

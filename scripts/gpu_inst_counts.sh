#!/bin/bash
# GPU box: dynamic instruction counts (and pipe shares) of the compose kernel for prebuilt variants (variants/<tag>.so).
# Usage: gpu_inst_counts.sh tag1 tag2 ...   -> gpurun_out/inst_<tag>.csv
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
A="--steps 2 --warmup 1 --no-cpu --no-extras --no-check --streams 32"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_xu.sum,gpu__time_duration.sum,sm__cycles_active.avg
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  timeout 300 ncu --metrics $M --clock-control none -k regex:compose_frames -s 2 -c 1 --csv --log-file gpurun_out/inst_$tag.csv python bench.py $A > gpurun_out/inst_$tag.log 2>&1
  python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/inst_$tag.csv")) if len(r)>10]
h=rows[0]; i=h.index("Metric Name"); v=h.index("Metric Value")
print("$tag:", "  ".join("%s=%s" % (r[i].replace("sm__inst_executed_pipe_","").replace(".sum",""), r[v]) for r in rows[1:]))
P
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

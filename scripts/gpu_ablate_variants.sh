#!/bin/bash
# GPU box: kernel time of deliberately incomplete builds (no parity check). Usage: gpu_ablate_variants.sh tag1 tag2 ...
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  H264SCROLL_BENCH_NOCHECK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
  python - <<P
import json
try:
    l=json.loads(open("gpurun_out/bench_$tag.log").read().strip().split("\n")[-1]); print("$tag: kernel %.3f ms" % l["ms_per_step"])
except Exception as e:
    print("$tag: failed", e); print(open("gpurun_out/bench_$tag.err").read()[-600:])
P
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

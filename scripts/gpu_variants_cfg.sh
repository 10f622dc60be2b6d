#!/bin/bash
# GPU box: kernel-only bench of prebuilt library variants on the other named configurations.
# Usage: gpu_variants_cfg.sh "config4 config1" tag1 tag2 ...
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
wls=$1; shift
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  for wl in $wls; do
    timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_$wl.log 2> gpurun_out/bench_${tag}_$wl.err
    python - <<P
import json
try:
    l=json.loads(open("gpurun_out/bench_${tag}_$wl.log").read().strip().split("\n")[-1])
    print("$tag $wl: kernel %.4f ms  roofline %.4f  bit_exact %s" % (l["ms_per_step"], l["roofline"]["frac"], l.get("bit_exact")))
except Exception as e:
    print("$tag $wl: bench failed", e)
P
  done
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

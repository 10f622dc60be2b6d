#!/bin/bash
# GPU box: rewrite / unescape / parse parity and benches for prebuilt variants (variants/<tag>.so). Usage: gpu_variants3.sh tag1 tag2 ...
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  timeout 600 python -m pytest tests/test_gpu_rewrite.py tests/test_gpu_unescape.py tests/test_gpu_parse.py -m gpu -x -q 2>&1 | tail -3
  for wl in rewrite unescape; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --workload $wl > gpurun_out/bench_${wl}_$tag.log 2> gpurun_out/bench_${wl}_$tag.err
    python - <<P
import json
try:
    l=json.loads(open("gpurun_out/bench_${wl}_$tag.log").read().strip().split("\n")[-1])
    print("$tag $wl: kernel %.4f ms  roofline %.3f" % (l["ms_per_step"], l["roofline"]["frac"]))
except Exception as e:
    print("$tag $wl: failed", e)
P
  done
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

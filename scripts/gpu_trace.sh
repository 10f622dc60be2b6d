#!/bin/bash
# GPU box: per-warp event trace of one CTA from a COMPOSE_TRACE=1 build (variants/<tag>.so) -> gpurun_out/trace_<tag>.bin
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  python bench.py --steps 1 --warmup 0 --no-cpu --no-extras --no-check > /dev/null 2>&1
  mv gpurun_out/trace.bin gpurun_out/trace_$tag.bin
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

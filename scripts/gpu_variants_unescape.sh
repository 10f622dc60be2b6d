#!/bin/bash
# GPU box: test + bench the unescape kernel of several prebuilt library variants (variants/<tag>.so)
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  timeout 300 python -m pytest tests/test_gpu_unescape.py -m gpu -x -q 2>&1 | tail -1
  timeout 120 python bench.py --workload unescape --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$tag unescape ms', l['ms_per_step'], l['roofline']['frac'])"
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

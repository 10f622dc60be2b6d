#!/bin/bash
# GPU box: side-kernel parity + bench for prebuilt library variants (variants/<tag>.so)
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  echo "== $tag"
  timeout 300 python -m pytest tests/test_gpu_rewrite.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  python bench.py --workload rewrite --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rewrite: %.4f ms  frac %.4f' % (l['ms_per_step'], l['roofline']['frac']))"
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

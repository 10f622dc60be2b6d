#!/bin/bash
# GPU box: side-kernel parity + bench for prebuilt library variants (variants/<tag>.so).  WL=rewrite|unescape gpu_side_variants.sh tag...
WL=${WL:-rewrite}
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  echo "== $tag"
  timeout 300 python -m pytest tests/test_gpu_rewrite.py tests/test_gpu_unescape.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so   # (the tests rebuild the library when the variant was built with other flags)
  python bench.py --workload $WL --steps 20 --warmup 5 2>/dev/null | WL=$WL python -c "
import sys,json,os
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('%s: %.4f ms  frac %.4f' % (os.environ['WL'], l['ms_per_step'], l['roofline']['frac']))"
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

#!/bin/bash
# GPU box: parity (compose + rewrite) and compose/rewrite benches for prebuilt variants (variants/<tag>.so). Usage: gpu_variants2.sh tag1 tag2 ...
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for tag in "$@"; do
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rewrite.py tests/test_gpu_legacy_api.py -m gpu -x -q 2>&1 | tail -3
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload rewrite > gpurun_out/bench_rw_$tag.log 2> gpurun_out/bench_rw_$tag.err
  python - <<P
import json
for pre in ("bench_", "bench_rw_"):
    try:
        l=json.loads(open("gpurun_out/%s$tag.log" % pre).read().strip().split("\n")[-1])
        print("$tag %s: kernel %.3f ms  roofline %.3f  e2e %s ms" % (pre, l["ms_per_step"], l["roofline"]["frac"], l["e2e"].get("ms_per_step")))
    except Exception as e:
        print("$tag %s: failed" % pre, e)
P
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

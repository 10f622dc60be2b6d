#!/bin/bash
# GPU box: smoke + parity subset + kernel-only bench for prebuilt library variants (variants/<tag>.so).
# Usage: gpu_variants_r3.sh tag1 tag2:nopar ...   (":nopar" = bench only).  A variant that hangs (timeout) ends the whole run.
mkdir -p gpurun_out
cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
for spec in "$@"; do
  tag=${spec%%:*}
  cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so
  if [ "$spec" = "$tag" ]; then
    timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1
    rc=$?
    tail -2 gpurun_out/smoke_$tag.log | cut -c1-300
    if [ $rc -eq 124 ]; then echo "$tag: smoke HUNG - stopping"; break; fi
    if [ $rc -ne 0 ]; then echo "$tag: smoke FAILED rc=$rc"; continue; fi
    timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_batch.py -m gpu -x -q > gpurun_out/parity_$tag.log 2>&1
    rc=$?
    tail -3 gpurun_out/parity_$tag.log | cut -c1-300
    if [ $rc -eq 124 ]; then echo "$tag: parity HUNG - stopping"; break; fi
    if [ $rc -ne 0 ]; then echo "$tag: parity FAILED rc=$rc"; fi
    cp variants/$tag.so h264-scroll-encoder_b200/libh264scroll_b200.so   # (the tests rebuild the library when the variant was built with other flags)
  fi
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
  python - <<P
import json
try:
    l=json.loads(open("gpurun_out/bench_$tag.log").read().strip().split("\n")[-1])
    oc=l.get("other_configs",{})
    print("$tag: kernel %.4f ms  roofline %.4f  e2e %.2f ms  bit_exact %s  others %s" % (l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["ms_per_step"], l.get("bit_exact"), {k:(v.get("roofline",{}).get("frac") if isinstance(v,dict) else v) for k,v in oc.items()}))
except Exception as e:
    print("$tag: bench failed", e)
P
done
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so

#!/bin/bash
# GPU box: the default bench line (with other_configs / side_kernels / latency) and one line per named configuration.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}_default.log 2> gpurun_out/bench_${tag}_default.err; echo "default rc=$?"; tail -3 gpurun_out/bench_${tag}_default.err | cut -c1-300
for wl in config4 config2 config1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl > gpurun_out/bench_${tag}_$wl.log 2> gpurun_out/bench_${tag}_$wl.err; echo "$wl rc=$?"; tail -3 gpurun_out/bench_${tag}_$wl.err | cut -c1-300
done
timeout 300 python bench.py --workload latency > gpurun_out/bench_${tag}_latency.log 2> gpurun_out/bench_${tag}_latency.err; echo "latency rc=$?"; tail -3 gpurun_out/bench_${tag}_latency.err | cut -c1-300
python - <<P
import json
for wl in ("default","config4","config2","config1"):
    try:
        l=json.loads(open("gpurun_out/bench_${tag}_%s.log"%wl).read().strip().split("\n")[-1])
        print("%s: kernel %.4f ms  %.2f GMB/s  frac %.4f  e2e %.2f ms (%.2f GMB/s)  bit_exact %s/%s  cpu %.1f MMB/s" % (wl, l["ms_per_step"], l["value"]/1e9, l["roofline"]["frac"], l["e2e"]["ms_per_step"], l["e2e"]["value"]/1e9, l.get("bit_exact"), l.get("frames_checked"), l.get("cpu_baseline",{}).get("value",0)/1e6))
        if wl=="default":
            for k,v in l.get("other_configs",{}).items(): print("   other", k, "frac %.4f kernel %.3f ms bit_exact %s" % (v["frac"], v["kernel_ms"], v["bit_exact"]))
            for k,v in l.get("side_kernels",{}).items(): print("   side", k, "frac %.4f ms %.3f" % (v["frac"], v["ms"]))
            print("   latency", l.get("latency"))
    except Exception as e: print(wl, "failed", e)
print(open("gpurun_out/bench_${tag}_latency.log").read()[:600])
P

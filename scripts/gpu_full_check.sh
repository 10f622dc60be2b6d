#!/bin/bash
# GPU box: the whole -m gpu suite, a fuzz soak, smoke and the kernel-only bench lines of the library as it is built in-tree.
# Usage: gpu_full_check.sh <tag> [fuzz cases] [fuzz seed]
tag=${1:-chk}; cases=${2:-1500}; seed=${3:-770000}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$tag.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/gputests_$tag.log | cut -c1-200
timeout 900 python scripts/gpu_fuzz.py $cases $seed > gpurun_out/fuzz_$tag.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz_$tag.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in compose config4 config2 config1 rewrite; do
  timeout 250 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/bench_${tag}_$wl.log 2> gpurun_out/bench_${tag}_$wl.err
  python - <<P
import json
try:
    l=json.loads(open("gpurun_out/bench_${tag}_$wl.log").read().strip().split("\n")[-1])
    print("$tag $wl: kernel %.4f ms  roofline %.4f  bit_exact %s" % (l["ms_per_step"], l["roofline"]["frac"], l.get("bit_exact")))
except Exception as e:
    print("$tag $wl: bench failed", e)
P
done

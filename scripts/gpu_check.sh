#!/bin/bash
# GPU box: parity tests then a bench run; logs under gpurun_out/<tag>.*
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.log

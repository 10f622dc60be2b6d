#!/bin/bash
# GPU box: end-to-end step time of the pipelined host-batch call against the number of frame chunks
for c in 4 6 8 10 12 16 24; do
  python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-check --e2e-chunks $c 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('chunks $c: e2e %.3f ms  (kernel %.4f ms)' % (l['e2e']['ms_per_step'], l['ms_per_step']))"
done

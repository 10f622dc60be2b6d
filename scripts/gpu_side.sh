#!/bin/bash
# GPU box: parity tests and bench lines of the side kernels (slice rewrite, EBSP -> RBSP)
timeout 600 python -m pytest tests/test_gpu_rewrite.py tests/test_gpu_unescape.py tests/test_gpu_parse.py tests/test_gpu_configs_full.py -m gpu -x -q 2>&1 | tail -2
for wl in rewrite unescape; do
  python bench.py --workload $wl --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$wl: %.4f ms  frac %.4f  bit_exact %s' % (l['ms_per_step'], l['roofline']['frac'], l.get('bit_exact')))"
done

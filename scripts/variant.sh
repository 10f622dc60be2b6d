#!/bin/bash
# build a kernel variant here (CPU box) : variant.sh "<nvcc extra flags>"
H264SCROLL_NVCC_EXTRA="$1" python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as ge
ge.load_package()
from h264_scroll_encoder_b200 import build
build.build(force=True, verbose=True)
" > /tmp/variant_build.log 2>&1 || { grep -B2 -A6 "error" /tmp/variant_build.log | head -40; exit 1; }
grep -A2 "compose_frames_kernel" /tmp/variant_build.log | grep -E "spill|Used"

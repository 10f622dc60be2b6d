#!/bin/bash
# GPU box: one full ncu capture each of the rewrite, unescape and raster kernels on their bench workloads. Usage: gpu_profile_side.sh <tag>
tag=${1:-side}
mkdir -p gpurun_out
python bench.py --workload rewrite --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_rw_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:rewrite_slices -s 4 -c 1 -o gpurun_out/prof_${tag}_rewrite python bench.py --workload rewrite --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_rw_$tag.log 2>&1
python bench.py --workload unescape --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_un_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:unescape_kernel -s 4 -c 1 -o gpurun_out/prof_${tag}_unescape python bench.py --workload unescape --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_un_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_regions -s 2 -c 1 -o gpurun_out/prof_${tag}_raster python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ra_$tag.log 2>&1
ls -la gpurun_out/prof_${tag}_*.ncu-rep

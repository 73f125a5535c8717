#!/usr/bin/env bash
# Round-2 GPU session R (one B200): ncu --set full of the tiled kernels on S3B (both A x + dual step and A'y + interaction).
set -u
mkdir -p gpurun_out/r2r
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_tile_kernel" -s 16 -c 4 -o gpurun_out/r2r/prof_tile_s3b python bench.py --workload S3B --steps 60 --warmup 45 --no-cpu-baseline > gpurun_out/r2r/ncu.log 2> gpurun_out/r2r/ncu.err
echo "exit $?"; tail -n 2 gpurun_out/r2r/ncu.err | cut -c1-200

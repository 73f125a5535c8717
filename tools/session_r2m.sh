#!/usr/bin/env bash
# Round-2 GPU session M (one B200): tiled shape with long-row segments and with a few unstaged tiles (forced), the default
# shape selection on S3 / S3B / S3D / S5, whole suite.
set -u
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=1800 run pytest_all python -m pytest tests -q -m gpu
B200PDLP_TIMING=1 run bench_s3b python bench.py --workload S3B --no-cpu-baseline --parity
B200PDLP_TILE=1 B200PDLP_TIMING=1 run bench_s3b_force python bench.py --workload S3B --no-cpu-baseline --parity
B200PDLP_TILE=1 B200PDLP_TIMING=1 run bench_s3d_force python bench.py --workload S3D --no-cpu-baseline
B200PDLP_TIMING=1 run bench_s3d python bench.py --workload S3D --no-cpu-baseline
B200PDLP_TIMING=1 run bench_s5 python bench.py --workload S5 --no-cpu-baseline --parity
B200PDLP_TIMING=1 run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
grep -h '"metric"' $O/bench_*.log | cut -c1-300
grep -h "SpMV shapes" $O/*.err | sort | uniq -c
tail -n 8 $O/pytest_all.log | cut -c1-300

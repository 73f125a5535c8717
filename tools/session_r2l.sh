#!/usr/bin/env bash
# Round-2 GPU session L (one B200): the tiled SpMV shape (shared-memory staged gathers; bulk copy vs cooperative loads) on the
# banded workloads, the S3 converged-parity test, the new tests.
set -u
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=600 run pytest_new python -m pytest tests/test_gpu_solve.py -q -m gpu -k "tiled or s3_converged or light"
B200PDLP_TIMING=1 run bench_s3b python bench.py --workload S3B --no-cpu-baseline
B200PDLP_TILE=1 B200PDLP_TIMING=1 run bench_s3b_tile1 python bench.py --workload S3B --no-cpu-baseline --parity
B200PDLP_TILE=2 B200PDLP_TIMING=1 run bench_s3b_tile2 python bench.py --workload S3B --no-cpu-baseline
B200PDLP_TILE=1 B200PDLP_TIMING=1 run bench_s3d_tile1 python bench.py --workload S3D --no-cpu-baseline
B200PDLP_TILE_A=1 B200PDLP_TIMING=1 run bench_s3b_tileA python bench.py --workload S3B --no-cpu-baseline
B200PDLP_TILE=1 B200PDLP_TIMING=1 run bench_s3_tile1 python bench.py --no-cpu-baseline
T=300 run ncu_tile env B200PDLP_TILE=1 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_tile_kernel" -s 20 -c 4 -o $O/prof_tile python bench.py --workload S3B --steps 60 --warmup 45 --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
grep -h "SpMV shapes" $O/*.err | sort | uniq -c
tail -n 15 $O/pytest_new.log | cut -c1-300

#!/usr/bin/env bash
# Round-2 GPU session N (one B200): final validation -- smoke(), the whole GPU suite, the bench lines of record with the
# final defaults (S3 in the driver's window and at 2000 steps, S3B / S3D with the tiled shape chosen by the prologue).
set -u
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=300 run smoke python __graft_entry__.py smoke
T=1800 run pytest_all python -m pytest tests -q -m gpu
run bench_s20 python bench.py --steps 20 --warmup 5
B200PDLP_TIMING=1 run bench_s3b python bench.py --workload S3B --no-cpu-baseline --parity
B200PDLP_TIMING=1 run bench_s3d python bench.py --workload S3D --no-cpu-baseline --parity
run bench_default python bench.py --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
grep -h "SpMV shapes" $O/*.err | sort | uniq -c | cut -c1-200
tail -n 3 $O/smoke.log; tail -n 8 $O/pytest_all.log | cut -c1-300

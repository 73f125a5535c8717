#!/usr/bin/env bash
# Round-2 GPU session H (one B200): the whole GPU suite after the logical-shard assembly fix (host rendezvous around the
# barriers of the assembly), the structural switch of the A x shape (S3D / S3B / S3 defaults).
set -u
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=300 run smoke python __graft_entry__.py smoke
T=1800 run pytest_all python -m pytest tests -q -m gpu
B200PDLP_TIMING=1 run bench_s3d python bench.py --workload S3D --no-cpu-baseline
B200PDLP_TIMING=1 run bench_s3b python bench.py --workload S3B --no-cpu-baseline
B200PDLP_TIMING=1 run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py
run bench_ref python bench.py --impl reference --steps 20 --warmup 5
run bench_hipdlp python bench.py --solver hipdlp --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 3 $O/smoke.log $O/smoke.err
grep -h "sectors per" $O/*.err | sort | uniq -c
tail -n 30 $O/pytest_all.log | cut -c1-400

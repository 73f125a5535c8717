#!/usr/bin/env bash
# Round-2 GPU session J (one B200): the GPU suite with the new multi-GPU defaults (device-side checks, per-rank device
# formulate + scale) as the logical-shard tests see them; bench lines of record for round 2.
set -u
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=1800 run pytest_all python -m pytest tests -q -m gpu
run bench_s20 python bench.py --steps 20 --warmup 5
run bench_ref_s20 python bench.py --impl reference --steps 20 --warmup 5
run bench_default python bench.py --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 12 $O/pytest_all.log | cut -c1-400

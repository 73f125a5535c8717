#!/usr/bin/env bash
# Round-2 GPU session K (one B200): check sweeps side by side (graph branches), peer loads hoisted in the multi-GPU primal kernel
# (logical shards), whole suite.
set -u
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=1800 run pytest_all python -m pytest tests -q -m gpu
run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py --no-cpu-baseline
B200PDLP_NO_GRAPH=1 T=600 run pytest_nograph python -m pytest tests/test_gpu_solve.py -q -m gpu -k "light or tree or golden"
T=300 run ncu_launches_s20 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $O/launches_s20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 8 $O/pytest_all.log $O/pytest_nograph.log | cut -c1-300

#!/usr/bin/env bash
# Round-2 GPU session Q (one B200): last sanity run on the final library.
set -u
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2> $O/smoke.err; echo "smoke exit $?"; tail -n 2 $O/smoke.log
timeout 300 python -m pytest tests/test_gpu_solve.py -q -m gpu -k "tiled or light or s2_converged" > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 $O/pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_s20.log 2> $O/bench_s20.err; echo "bench exit $?"; cut -c1-200 $O/bench_s20.log

#!/usr/bin/env bash
# Round-2 GPU session B (one B200): full GPU suite after the fixes of session A, logical shards with phase laps (and the
# NO_GRAPH bisect), ncu launch lists (20-step bench: what a check iteration costs) + one --set full capture of the SpMV
# kernels, the software-pipelined SpMV variants.
set -u
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
B200PDLP_TIMING=1 T=300 run shards2 python tests/logical_shards_child.py 2 synthetic threads
B200PDLP_TIMING=1 B200PDLP_NO_GRAPH=1 T=300 run shards2_nograph python tests/logical_shards_child.py 2 synthetic threads
B200PDLP_TIMING=1 B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards2_devcheck python tests/logical_shards_child.py 2 synthetic threads
B200PDLP_TIMING=1 T=300 run shards_c python tests/logical_shards_child.py 2 synthetic c_entry
T=300 run shards4_dense python tests/logical_shards_child.py 4 dense threads
T=1500 run pytest_all python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_logical_shards.py
B200PDLP_TIMING=1 run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py
T=300 run ncu_launches_s20 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_s20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
T=400 run ncu_full ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -s 4 -c 8 -o $O/prof_spmv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
B200PDLP_SPMV_CTAS_PER_SM=4 T=400 run pytest_pipe python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -q -m gpu -k "synthetic or s2 or long_rows or dense or step_kernels or residual"
for k in 3 4 6; do B200PDLP_SPMV_CTAS_PER_SM=$k run bench_pipe$k python bench.py --no-cpu-baseline; done
B200PDLP_SPMV_CTAS_PER_SM=4 run bench_pipe4_s3b python bench.py --no-cpu-baseline --workload S3B
run bench_s5 python bench.py --workload S5 --no-cpu-baseline --parity
grep -h '"metric"' $O/bench_*.log | cut -c1-420
tail -n 25 $O/shards2.err $O/shards2.log $O/shards2_nograph.err $O/shards2_devcheck.err $O/shards_c.err $O/shards4_dense.log | cut -c1-300
tail -n 15 $O/pytest_all.log $O/pytest_pipe.log | cut -c1-300

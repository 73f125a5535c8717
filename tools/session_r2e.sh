#!/usr/bin/env bash
# Round-2 GPU session E (one B200): the light check (dense-check phase), the diagnostics of the three logical-shard failures of
# session C, the pipelined A x with the partial first batch, the multi-diagonal workload S3D.
set -u
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=900 run pytest_fast python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_logical_shards.py -k "not s2_converged"
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards3_adlittle_dev python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=0 T=300 run shards4_dense_host python tests/logical_shards_child.py 4 dense threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards4_dense_dev python tests/logical_shards_child.py 4 dense threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards2_dense_dev python tests/logical_shards_child.py 2 dense threads
run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_LIGHT_CHECK=0 run bench_s20_nolight python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_LIGHT_CHECK=0 B200PDLP_FUSED_CHECK=1 run bench_s20_fused python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py --no-cpu-baseline --parity
B200PDLP_FUSED_CHECK=1 run bench_default_fused python bench.py --no-cpu-baseline
run s2_spread python tools/s2_gpu_spread.py 1e-4 1e-6
for k in 3 4 6; do B200PDLP_SPMV_A_CTAS_PER_SM=$k run bench_a$k python bench.py --no-cpu-baseline; done
run bench_s3d python bench.py --workload S3D --no-cpu-baseline
B200PDLP_SPMV_A_CTAS_PER_SM=4 run bench_s3d_a4 python bench.py --workload S3D --no-cpu-baseline
run bench_s3b python bench.py --workload S3B --no-cpu-baseline
T=300 run ncu_launches_s20 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $O/launches_s20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
T=400 run ncu_full_s3d ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_kernel" -s 20 -c 4 -o $O/prof_s3d python bench.py --workload S3D --steps 60 --warmup 45 --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 30 $O/pytest_fast.log | cut -c1-300
cat $O/shards*.log | cut -c1-1500

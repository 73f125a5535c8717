#!/usr/bin/env bash
# Round-2 GPU session C (one B200): the full GPU suite (logical shards included), check-iteration cost after the session-B
# fixes (launch list), the A'y pipelining sweep, the ncu --set full capture of the pass kernels for profiles/.
set -u
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
T=1800 run pytest_all python -m pytest tests -q -m gpu
B200PDLP_TIMING=1 run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py
for k in 0 2 4; do B200PDLP_SPMV_AT_CTAS_PER_SM=$k run bench_at$k python bench.py --no-cpu-baseline; done
B200PDLP_SPMV_A_CTAS_PER_SM=2 run bench_a2 python bench.py --no-cpu-baseline
T=300 run ncu_launches_s20 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_s20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
T=300 run ncu_launches_pass ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file $O/launches_pass.csv python bench.py --steps 120 --warmup 5 --no-cpu-baseline
T=400 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_kernel|primal_step_kernel|step_rule_kernel" -s 20 -c 8 -o $O/prof_pass python bench.py --steps 60 --warmup 45 --no-cpu-baseline
run bench_s5 python bench.py --workload S5 --no-cpu-baseline --parity
run bench_s2 python bench.py --workload S2 --to-tolerance 1e-4
run bench_ref_s3 python bench.py --impl reference --steps 20 --warmup 5
grep -h '"metric"' $O/bench_*.log | cut -c1-420
tail -n 30 $O/pytest_all.log | cut -c1-300

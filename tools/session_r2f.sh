#!/usr/bin/env bash
# Round-2 GPU session F (one B200): logical-shard diagnostics (B200PDLP_DEBUG_MG), the whole logical-shard matrix with the
# multi-GPU light check, check sweeps with few CTAs, ncu --set full of the check kernels.
set -u
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
export B200PDLP_DEBUG_MG=1
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards3_adlittle_dev python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LIGHT_CHECK=0 T=300 run shards3_adlittle_dev_nolight python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_GRAPH=1 T=300 run shards3_adlittle_dev_nograph python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=0 T=300 run shards3_adlittle_host python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards4_dense_dev python tests/logical_shards_child.py 4 dense threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards3_synth_dev python tests/logical_shards_child.py 3 synthetic threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards2_synth_dev python tests/logical_shards_child.py 2 synthetic threads
unset B200PDLP_DEBUG_MG
T=900 run pytest_light python -m pytest tests/test_gpu_solve.py tests/test_gpu_instances.py -q -m gpu
run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py --no-cpu-baseline
T=300 run ncu_launches_s20 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $O/launches_s20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
T=400 run ncu_full_checks ncu --set full --clock-control none --import-source on -k regex:"restart_sweep|check_cols_sweep|check_rows_sweep|check_decide|check_finish|step_rule" -s 12 -c 10 -o $O/prof_checks python bench.py --steps 20 --warmup 5 --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 12 $O/pytest_light.log | cut -c1-300
for f in $O/shards*.log; do echo "--- $f"; cut -c1-1800 $f; done
for f in $O/shards*.err; do echo "--- $f"; grep "mg-debug" $f | cut -c1-600; done

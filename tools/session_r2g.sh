#!/usr/bin/env bash
# Round-2 GPU session G (one B200): per-segment diagnostics of the logical-shard solution assembly (world 3 / 4), the host
# rendezvous experiment, restart sweep with hoisted loads, B200PDLP_FUSE_K4 A/B.
set -u
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
nvidia-smi -L
export B200PDLP_DEBUG_MG=1
B200PDLP_MG_DEVICE_CHECK=0 T=300 run shards3_adlittle_host python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=0 B200PDLP_LOCAL_RENDEZVOUS=1 T=300 run shards3_adlittle_host_rdv python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LOCAL_RENDEZVOUS=1 T=300 run shards3_adlittle_dev_rdv python tests/logical_shards_child.py 3 adlittle threads
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LOCAL_RENDEZVOUS=1 T=300 run shards4_dense_dev_rdv python tests/logical_shards_child.py 4 dense threads
B200PDLP_MG_DEVICE_CHECK=1 T=300 run shards4_dense_dev python tests/logical_shards_child.py 4 dense threads
unset B200PDLP_DEBUG_MG
T=600 run pytest_s2 python -m pytest tests/test_gpu_solve.py -q -m gpu -k "s2_converged or light"
run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_FUSE_K4=1 run bench_s20_fuse python bench.py --no-cpu-baseline --steps 20 --warmup 5
run bench_default python bench.py --no-cpu-baseline
B200PDLP_FUSE_K4=1 run bench_default_fuse python bench.py --no-cpu-baseline --parity
B200PDLP_FUSE_K4=1 T=600 run pytest_fuse python -m pytest tests/test_gpu_solve.py tests/test_gpu_instances.py -q -m gpu -k "not s2_converged"
T=300 run ncu_launches_fuse env B200PDLP_FUSE_K4=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $O/launches_s20_fuse.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline
grep -h '"metric"' $O/bench_*.log | cut -c1-300
tail -n 8 $O/pytest_s2.log $O/pytest_fuse.log | cut -c1-300
for f in $O/shards*.log; do echo "--- $f"; cut -c1-1500 $f; done
for f in $O/shards*.err; do echo "--- $f"; grep "mg-debug" $f | cut -c1-400; done

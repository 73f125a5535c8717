#!/usr/bin/env bash
# Round-2 GPU session A (one B200): first hardware run of the device-side check iterations and of the device-resident
# prologue, diagnosis of the two tree-mode instance cases that failed in round 1, logical shards, bench variants.
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 4 "$O/$name.log"; }
nvidia-smi -L
T=400 run diag python tools/diag_r2a.py
B200PDLP_HOST_CHECK=1 B200PDLP_DEVICE_PREP=0 T=400 run diag_old python tools/diag_r2a.py
T=900 run pytest_prep python -m pytest tests/test_gpu_device_prep.py -q -m gpu
T=600 run pytest_solve python -m pytest tests/test_gpu_solve.py tests/test_gpu_kernels.py -q -m gpu
T=200 run child_shards2 python tests/logical_shards_child.py 2 synthetic threads
T=200 run child_shards_c python tests/logical_shards_child.py 2 synthetic c_entry
B200PDLP_TIMING=1 run bench_default python bench.py
B200PDLP_TIMING=1 B200PDLP_DEVICE_PREP=0 run bench_hostprep python bench.py --no-cpu-baseline
B200PDLP_TIMING=1 B200PDLP_HOST_CHECK=1 B200PDLP_DEVICE_PREP=0 run bench_old python bench.py --no-cpu-baseline
B200PDLP_TIMING=1 run bench_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_TIMING=1 B200PDLP_HOST_CHECK=1 B200PDLP_DEVICE_PREP=0 run bench_s20_old python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_PDL=1 run bench_pdl1 python bench.py --no-cpu-baseline
B200PDLP_PDL=2 run bench_pdl2 python bench.py --no-cpu-baseline
run tts_ours python bench.py --workload S2 --no-cpu-baseline --to-tolerance 1e-4
run tts_ours_s3 python bench.py --workload S3 --no-cpu-baseline --to-tolerance 1e-4
B200PDLP_SPMV_CTAS_PER_SM=4 T=400 run pytest_pipe python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -q -m gpu -k 'synthetic or s2 or long_rows or dense or step_kernels'
for k in 3 4 6; do B200PDLP_SPMV_CTAS_PER_SM=$k run bench_pipe$k python bench.py --no-cpu-baseline; done
T=120 run dsmem_gather tools/dsmem_gather_bench
T=120 run gather_bench tools/gather_bench
B200PDLP_TIMING=1 run bench_s3b python bench.py --workload S3B --no-cpu-baseline
run bench_hipdlp python bench.py --solver hipdlp --steps 400
T=900 run pytest_rest python -m pytest tests -q -m gpu --deselect tests/test_gpu_device_prep.py --deselect tests/test_gpu_solve.py --deselect tests/test_gpu_kernels.py -rxX
grep -h '"metric"\|"impl"' $O/bench_*.log $O/tts_*.log | cut -c1-700
tail -n 30 $O/diag.log $O/diag_old.log $O/pytest_prep.log $O/pytest_solve.log $O/child_shards2.log $O/child_shards2.err $O/child_shards_c.err | cut -c1-900

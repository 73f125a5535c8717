#!/usr/bin/env bash
# Round-2 GPU session D (2 B200s: gpurun --gpus 2): the multi-GPU path against the oracle, then the opt-in round-2 variants
# (device-side checks + light check, NCCL-free run, device formulate+scale per rank) one at a time and together.
set -u
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 3 "$O/$name.log"; }
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" bench.py --gpus 2 "${@:2}"; }
nvidia-smi -L
T=900 run pytest_multi python -m pytest tests/test_gpu_multi.py -q -m gpu
B200PDLP_MG_DEVICE_CHECK=1 T=900 run pytest_multi_devcheck python -m pytest tests/test_gpu_multi.py -q -m gpu -k p2p
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LIGHT_CHECK=0 T=900 run pytest_multi_devcheck_nolight python -m pytest tests/test_gpu_multi.py -q -m gpu -k p2p
T=600 run bench2_base tr 29701 --no-cpu-baseline
B200PDLP_MG_DEVICE_CHECK=1 T=600 run bench2_devcheck tr 29702 --no-cpu-baseline --parity
B200PDLP_MG_DEVICE_CHECK=1 T=600 run bench2_devcheck_s20 tr 29703 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LIGHT_CHECK=0 T=600 run bench2_devcheck_nolight_s20 tr 29708 --no-cpu-baseline --steps 20 --warmup 5
T=600 run bench2_base_s20 tr 29704 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 T=600 run bench2_nonccl tr 29705 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 B200PDLP_MG_DEVICE_PREP=1 B200PDLP_TIMING=1 T=600 run bench2_all tr 29706 --no-cpu-baseline --parity --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 T=600 run bench2_s5 tr 29707 --workload S5 --no-cpu-baseline --parity
grep -h '"metric"' $O/bench2_*.log | cut -c1-1200
tail -n 20 $O/pytest_multi.log $O/pytest_multi_devcheck.log $O/pytest_multi_devcheck_nolight.log | cut -c1-600

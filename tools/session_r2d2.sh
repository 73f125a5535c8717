#!/usr/bin/env bash
# Round-2 GPU session D' (2 B200s): the bench lines session D lost to a script bug (its helper was called `tr`, and `timeout tr`
# runs /usr/bin/tr) -- host-driven checks vs device-side (+ light) checks, NCCL-free run, device formulate+scale per rank.
set -u
mkdir -p gpurun_out/r2d2
O=gpurun_out/r2d2
mg() { local name=$1 port=$2; shift 2; echo "=== $name: bench.py --gpus 2 $*"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus 2 "$@" > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 1 "$O/$name.log" | cut -c1-300; }
nvidia-smi -L
B200PDLP_MG_DEVICE_CHECK=0 mg bench2_host_s20 29704 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 mg bench2_dev_s20 29703 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_LIGHT_CHECK=0 mg bench2_dev_nolight_s20 29708 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 mg bench2_dev_nonccl_s20 29705 --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 B200PDLP_MG_DEVICE_PREP=1 B200PDLP_TIMING=1 mg bench2_all_s20 29706 --no-cpu-baseline --parity --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=0 mg bench2_host 29701 --no-cpu-baseline
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 B200PDLP_MG_DEVICE_PREP=1 mg bench2_all 29702 --no-cpu-baseline --parity
B200PDLP_MG_DEVICE_CHECK=1 B200PDLP_NO_NCCL=1 B200PDLP_MG_DEVICE_PREP=1 mg bench2_s5 29707 --workload S5 --no-cpu-baseline --parity
grep -h '"metric"' $O/bench2_*.log | cut -c1-1500
grep -h "Error\|error\|Traceback" $O/*.err | head -20

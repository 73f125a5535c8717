#!/usr/bin/env bash
# Round-2 GPU session I (8 B200s: gpurun --gpus 8): the scaling table the driver measures (bench.py --gpus N --steps 20 --warmup 5
# with default settings) plus the 2000-step lines, S5 at 8 GPUs with its parity block.
set -u
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-400}" "$@" ) > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 2 "$O/$name.log" | cut -c1-400; }
tr() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus "$n" "$@"; }
nvidia-smi -L | head -8
run bench8_s20 tr 8 29801 --no-cpu-baseline --steps 20 --warmup 5 --parity
run bench8 tr 8 29802 --no-cpu-baseline
run bench8_s5 tr 8 29803 --workload S5 --no-cpu-baseline --parity
run bench4_s20 tr 4 29804 --no-cpu-baseline --steps 20 --warmup 5
run bench4 tr 4 29805 --no-cpu-baseline
run bench2_s20 tr 2 29806 --no-cpu-baseline --steps 20 --warmup 5
run bench2 tr 2 29807 --no-cpu-baseline
run bench1_s20 python bench.py --no-cpu-baseline --steps 20 --warmup 5
B200PDLP_MG_DEVICE_CHECK=0 run bench8_hostcheck_s20 tr 8 29808 --no-cpu-baseline --steps 20 --warmup 5
grep -h '"metric"' $O/bench*.log | cut -c1-1500

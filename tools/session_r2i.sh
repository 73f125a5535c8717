#!/usr/bin/env bash
# Round-2 GPU session I (8 B200s: gpurun --gpus 8): the scaling table the driver measures (bench.py --gpus N --steps 20 --warmup 5
# with default settings) plus the 2000-step lines, S5 at 8 GPUs with its parity block.
set -u
mkdir -p gpurun_out/r2i2
O=gpurun_out/r2i2
# (the helper must not be called `tr`: `timeout tr ...` runs /usr/bin/tr -- session D lost its bench lines to that)
mg() { local name=$1 n=$2 port=$3; shift 3; echo "=== $name: bench.py --gpus $n $*"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus "$n" "$@" > "$O/$name.log" 2> "$O/$name.err"; echo "    exit $?"; tail -n 1 "$O/$name.log" | cut -c1-400; }
nvidia-smi -L | head -8
mg bench8_s20 8 29801 --no-cpu-baseline --steps 20 --warmup 5 --parity
mg bench8 8 29802 --no-cpu-baseline
mg bench8_s5 8 29803 --workload S5 --no-cpu-baseline --parity
mg bench4_s20 4 29804 --no-cpu-baseline --steps 20 --warmup 5
mg bench4 4 29805 --no-cpu-baseline
mg bench2_s20 2 29806 --no-cpu-baseline --steps 20 --warmup 5
mg bench2 2 29807 --no-cpu-baseline
grep -h '"metric"' $O/bench*.log | cut -c1-1500

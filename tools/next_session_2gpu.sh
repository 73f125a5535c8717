#!/usr/bin/env bash
# Second GPU call of the next round (two B200s):  gpurun --gpus 2 --timeout 1500 -- 'bash tools/next_session_2gpu.sh'
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "    exit $?"; tail -n 3 "gpurun_out/$name.log"; }
T=900 run pytest_multi python -m pytest tests/test_gpu_multi.py tests/test_gpu_logical_shards.py -q -rxX
run bench2 $TR bench.py --gpus 2 --no-cpu-baseline
B200PDLP_NO_NCCL=1 run bench2_no_nccl $TR bench.py --gpus 2 --no-cpu-baseline
B200PDLP_DEVICE_SETUP=1 run bench2_devsetup $TR bench.py --gpus 2 --no-cpu-baseline
grep -h '"metric"' gpurun_out/bench2*.log | cut -c1-400

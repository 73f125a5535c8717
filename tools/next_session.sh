#!/usr/bin/env bash
# First GPU call of the next round (one B200):  gpurun --timeout 3000 -- 'bash tools/next_session.sh'   (about 25-35 GPU-minutes)
# Runs everything that was written after round 1's GPU budget was spent and collects the evidence in gpurun_out/.
# Every step has its own timeout; a failing step does not stop the others.
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "    exit $?"; tail -n 3 "gpurun_out/$name.log"; }

# 1. the GPU suite; -rxX lists every xfail / XPASS with its reason (logical shards, device-side setup, instance sweep)
T=1500 run pytest_gpu python -m pytest tests -q -m gpu -rxX
# 1b. the gated cases once more, directly, so that their details (which case, which array) land in the logs
T=400 run child_devsetup1 python tests/device_scaling_child.py 1
T=400 run child_devsetup2 python tests/device_scaling_child.py 2
T=300 run child_shards2 python tests/logical_shards_child.py 2 synthetic threads
T=300 run child_shards_c python tests/logical_shards_child.py 2 synthetic c_entry
T=700 run child_hipdlp python tests/hipdlp_child.py
# 2. where a short solve's time goes (setup / solve phases / teardown laps on stderr)
B200PDLP_TIMING=1 run bench_default python bench.py
# 3. device-side prologue: scaling only, scaling + sliced-ELL fill
B200PDLP_TIMING=1 B200PDLP_DEVICE_SETUP=1 run bench_devsetup1 python bench.py --no-cpu-baseline
B200PDLP_TIMING=1 B200PDLP_DEVICE_SETUP=2 run bench_devsetup2 python bench.py --no-cpu-baseline
# 4. programmatic dependent launch of the pass kernels
B200PDLP_PDL=1 run bench_pdl1 python bench.py --no-cpu-baseline
B200PDLP_PDL=2 run bench_pdl2 python bench.py --no-cpu-baseline
# 5. wall-clock to solution, both arms, S2
run tts_ours python bench.py --workload S2 --no-cpu-baseline --to-tolerance 1e-4
run tts_reference python bench.py --workload S2 --impl reference --to-tolerance 1e-4
grep -h '"metric"\|"impl"' gpurun_out/bench_*.log gpurun_out/tts_*.log | cut -c1-400

#!/usr/bin/env bash
# Round-2 GPU session O (one B200): the S2 1e-8 objective-parity test against the reference's 51-minute run.
set -u
mkdir -p gpurun_out/r2o
timeout 600 python -m pytest tests/test_gpu_solve.py -q -m gpu -k "tight_tolerance or s2_converged" > gpurun_out/r2o/pytest.log 2> gpurun_out/r2o/pytest.err
echo "exit $?"; tail -n 15 gpurun_out/r2o/pytest.log | cut -c1-400

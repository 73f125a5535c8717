#!/usr/bin/env bash
# Round-2 GPU session P (one B200): the whole GPU suite on the final tree.
set -u
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2p/pytest_all.log 2> gpurun_out/r2p/pytest_all.err
echo "exit $?"; tail -n 12 gpurun_out/r2p/pytest_all.log | cut -c1-400

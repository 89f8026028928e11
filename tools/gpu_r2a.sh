#!/bin/bash
# first GPU check of the warp-stream compose kernel: parity, then A/B timing against the previous persistent kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/r2a_parity.log
B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 4 > gpurun_out/r2a_ws.log 2>&1; echo "ws rc=$?"
tail -6 gpurun_out/r2a_ws.log
B200_COMPOSE_IMPL=coop timeout 300 python tools/profile_run.py --reps 4 > gpurun_out/r2a_coop.log 2>&1; echo "coop rc=$?"
tail -2 gpurun_out/r2a_coop.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dag_order.py tests/test_gpu_nshortest.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2f_parity.log 2>&1; echo "parity rc=$?"; tail -6 gpurun_out/r2f_parity.log
B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2f_ws.log 2>&1; echo "ws rc=$?"
grep "^\[ws\]\|^\[trim\]" gpurun_out/r2f_ws.log | tail -2; tail -1 gpurun_out/r2f_ws.log | cut -c1-330
echo "== spread V=93"; B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 2 --spread --fanout --vocab 93 > gpurun_out/r2f_spread.log 2>&1
grep "^\[ws\]\|^\[trim\]" gpurun_out/r2f_spread.log | tail -2; tail -1 gpurun_out/r2f_spread.log | cut -c1-330

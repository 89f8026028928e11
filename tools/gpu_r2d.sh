#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dag_order.py -x -q > gpurun_out/r2d_dag.log 2>&1; echo "dag rc=$?"; tail -15 gpurun_out/r2d_dag.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nshortest.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/r2d_parity.log
B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2d_ws.log 2>&1; echo "ws rc=$?"
grep "^\[ws\]\|^\[trim\]" gpurun_out/r2d_ws.log | tail -2; tail -1 gpurun_out/r2d_ws.log | cut -c1-330
B200_COOP_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2d_bench.json; grep "dag-order" gpurun_out/r2d_bench.err | tail -2

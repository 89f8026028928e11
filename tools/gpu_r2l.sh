#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_dag_order.py -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
echo "== sssp composed (device order with path heads)"; B200_COOP_TRACE=1 timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | tail -2 | cut -c1-300
echo "== C5 in bench with trace"; B200_BATCH_TRACE=1 timeout 600 python bench.py --workload C5 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep "^\[split\]\|^\[batch\]\|ms_per_batch" | tail -5 | cut -c1-300
echo "== C5 probe with trace"; B200_BATCH_TRACE=1 timeout 600 python tools/c5_probe.py 2>&1 | grep "^\[split\]\|^\[batch\]\|^call" | tail -3 | cut -c1-300

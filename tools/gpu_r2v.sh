#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_sssp.py -x -q -m gpu 2>&1 | tail -8 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nshortest.py tests/test_gpu_full_size.py -x -q -m gpu -k "sssp or shortest or path" 2>&1 | tail -3
echo "== window DAG, full size"
timeout 600 python tools/profile_run.py --no-compose --sssp-window 1000 --reps 3 2>&1 | tail -3 | cut -c1-400
echo "== C4 layered, sweep forced"
B200_RELAX_VISIT_BUDGET=0 timeout 600 python tools/profile_run.py --no-compose --sssp --reps 3 2>&1 | tail -2 | cut -c1-400

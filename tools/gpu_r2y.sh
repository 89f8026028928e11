#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_sweep_sssp.py tests/test_gpu_nshortest.py -x -q -m gpu 2>&1 | tail -3
for sb in 128 64; do
echo "== window DAG, sweep sub-block $sb"
B200_SWEEP_SUB=$sb timeout 600 python tools/profile_run.py --no-compose --sssp-window 1000 --reps 3 2>&1 | tail -1 | cut -c1-330
done
echo "== C4 layered, sweep forced"
B200_RELAX_VISIT_BUDGET=0 timeout 600 python tools/profile_run.py --no-compose --sssp --reps 2 2>&1 | tail -1 | cut -c1-330

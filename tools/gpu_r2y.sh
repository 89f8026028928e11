#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_sweep_sssp.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_dag_order.py -x -q -m gpu -k "sweep or sssp or shortest or path or dag" 2>&1 | tail -3
echo "== window DAG"
timeout 600 python tools/profile_run.py --no-compose --sssp-window 1000 --reps 3 2>&1 | tail -1 | cut -c1-330
echo "== C4 composed-lattice props"
timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | tail -1 | cut -c1-330

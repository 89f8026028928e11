#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_dag_order.py -x -q -m gpu 2>&1 | tail -2
echo "== sssp on a composed-lattice property word, 8 reps"
timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 8 2>&1 | grep -E "ms_order" | cut -c150-330
echo "== host-API compose + sssp (e2e probe)"
timeout 600 python tools/e2e_probe.py 2>&1 | tail -9 | cut -c1-200

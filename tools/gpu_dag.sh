#!/bin/bash
cd "$(dirname "$0")/.."
for p in 1 0; do
echo "== pipelined=$p"
B200_DAG_PIPELINED=$p timeout 900 python -m pytest tests/test_gpu_dag_order.py -x -q -m gpu 2>&1 | tail -1
B200_DAG_PIPELINED=$p B200_COOP_TRACE=1 timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | grep -E "dag-order" | tail -3 | cut -c1-300
B200_DAG_PIPELINED=$p timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 5 2>&1 | grep -E "ms_order" | tail -3 | cut -c150-330
done

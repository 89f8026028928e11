#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dag_order.py -x -q 2>&1 | tail -3
for g in 2; do
echo "== dag order trace, CTAs/SM $g"
B200_DAG_CTAS_PER_SM=$g B200_COOP_TRACE=1 timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | grep -E "dag-order|ms_order" | tail -3 | cut -c1-400
done
echo "== no trace"
timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 6 2>&1 | grep -E "ms_order" | tail -6 | cut -c1-400

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 4 2 1; do
echo "== dag order trace, CTAs/SM $g"
B200_DAG_CTAS_PER_SM=$g B200_COOP_TRACE=1 timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | grep -E "dag-order|ms_order" | tail -2 | cut -c1-400
done
echo "== ncu source of k_dag_tree"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_dag_tree -c 1 -o gpurun_out/r2o_dag_tree -f python tools/profile_run.py --no-compose --sssp-top > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log

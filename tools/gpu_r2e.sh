#!/bin/bash
# ncu captures of the round-2 kernels + calibration of the extra workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_compose_ws -c 1 -o gpurun_out/r2_full_compose_ws -f python tools/profile_run.py --reps 1 > gpurun_out/r2e_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none -k "regex:k_trim_coop|k_ws_move|k_dag_tree|k_dag_orders|k_relax_coop|k_parents" -c 8 -o gpurun_out/r2_full_trim_dag_relax -f python tools/profile_run.py --reps 1 --sssp-top > gpurun_out/r2e_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --callers 1 > gpurun_out/r2e_ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
for v in 90 93 96; do
  echo "== spread V=$v"; timeout 300 python tools/profile_run.py --reps 2 --spread --fanout --vocab $v 2>&1 | tail -1 | cut -c1-400
done
for w in 1000 100; do
  echo "== window dag W=$w (1M states)"; timeout 300 python tools/profile_run.py --no-compose --sssp-window $w --scale 0.2 --reps 2 2>&1 | tail -1
done
echo "== window dag W=1000 (5M states)"; timeout 300 python tools/profile_run.py --no-compose --sssp-window 1000 --reps 2 2>&1 | tail -1

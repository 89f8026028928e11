#!/bin/bash
# final profile refresh: launch list of the bench command + full captures of the kernels that changed late in the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --callers 1 > gpurun_out/r2_final_bench_under_ncu.json 2> gpurun_out/r2_final_bench_under_ncu.err
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_final_launches_bench.csv)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_dag_tree|k_dag_orders|k_relax_coop|k_parents' -c 4 \
  -o gpurun_out/r2_final_sssp_kernels -f python tools/profile_run.py --no-compose --sssp-top > gpurun_out/r2_final_ncu1.log 2>&1
echo "sssp capture rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_relax_sweep' -c 1 \
  -o gpurun_out/r2_final_sweep -f python tools/profile_run.py --no-compose --sssp-window 1000 --scale 0.2 > gpurun_out/r2_final_ncu2.log 2>&1
echo "sweep capture rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_compose_ws|k_trim_coop|k_ws_move' -c 3 \
  -o gpurun_out/r2_final_compose -f python tools/profile_run.py > gpurun_out/r2_final_ncu3.log 2>&1
echo "compose capture rc=$?"
ls -la gpurun_out/r2_final_*

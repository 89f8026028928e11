#!/bin/bash
# memcheck / racecheck of the kernels written late in the round, on their small tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_RELAX_VISIT_BUDGET=0
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep_sssp.py -x -q -m gpu -k "forced and (0 or 2 or 1)" > gpurun_out/san_memcheck_sweep.log 2>&1; echo "memcheck sweep rc=$?"; tail -4 gpurun_out/san_memcheck_sweep.log | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep_sssp.py -x -q -m gpu -k "forced and (0 or 1)" > gpurun_out/san_racecheck_sweep.log 2>&1; echo "racecheck sweep rc=$?"; tail -4 gpurun_out/san_racecheck_sweep.log | cut -c1-200
unset B200_RELAX_VISIT_BUDGET
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dag_order.py -x -q -m gpu -k "small_shapes or deep" > gpurun_out/san_memcheck_dag.log 2>&1; echo "memcheck dag rc=$?"; tail -4 gpurun_out/san_memcheck_dag.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_nshortest.py -x -q -m gpu -k "unique" > gpurun_out/san_memcheck_unique.log 2>&1; echo "memcheck unique rc=$?"; tail -4 gpurun_out/san_memcheck_unique.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kat or batch or sigma" > gpurun_out/san_memcheck_compose.log 2>&1; echo "memcheck compose rc=$?"; tail -4 gpurun_out/san_memcheck_compose.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kat" > gpurun_out/san_racecheck_compose.log 2>&1; echo "racecheck compose rc=$?"; tail -4 gpurun_out/san_racecheck_compose.log | cut -c1-200

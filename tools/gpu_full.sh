#!/bin/bash
# what the driver runs at round end: the whole -m gpu suite, smoke(), and the default bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/full_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/full_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/full_smoke.log
timeout 1800 python bench.py > gpurun_out/full_bench_n1.json 2> gpurun_out/full_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/full_bench_n1.json; tail -3 gpurun_out/full_bench_n1.err | cut -c1-300
if [ "$1" = "ref" ]; then
  timeout 1800 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/full_bench_reference.json 2> gpurun_out/full_bench_reference.err; echo "reference rc=$?"
  cut -c1-600 gpurun_out/full_bench_reference.json
fi

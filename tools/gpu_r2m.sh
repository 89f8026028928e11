#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for env in "" "B200_DAG_NO_HEADS=1"; do
  echo "== sssp leg, $env"; env $env timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-extras --no-cpu-baseline --callers 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); s=d['sssp']; print('compose ms', round(d['ms_per_step'],3), 'sssp ms', round(s['ms_per_step'],3), 'order', round(s['ms_order_device_per_step'],3), 'relax', round(s['ms_relax_and_backtrace_per_step'],3))"
done
echo "== C5 alone"; timeout 600 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_batch'], d['inside_the_call_ms_per_step'])"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nshortest.py -x -q 2>&1 | tail -2

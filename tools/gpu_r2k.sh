#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_parity.py -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2k_tests.log
for c in 2 3; do
  for rep in 1 2; do
  echo "== callers=$c rep $rep"; timeout 600 python bench.py --steps 5 --warmup 3 --no-sssp --no-c5 --no-extras --no-cpu-baseline --callers $c 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), d['e2e'].get('concurrent_callers'))"
  done
done
echo "== bench incl. C5 after the other legs"; timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['c5']['ms_per_batch'], d['c5']['inside_the_call_ms_per_step'], d['e2e'].get('concurrent_callers'))"

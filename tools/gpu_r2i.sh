#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2i_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2i_parity.log
B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2i_ws.log 2>&1
grep "^\[ws\]\|^\[trim\]" gpurun_out/r2i_ws.log | tail -2; tail -1 gpurun_out/r2i_ws.log | cut -c90-330
for c in 1 2 3; do
  echo "== callers=$c (ws)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-sssp --no-c5 --no-extras --no-cpu-baseline --callers $c 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],3), 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value',)}, d['e2e'].get('concurrent_callers'))"
done
echo "== callers=3 (coop)"; B200_COMPOSE_IMPL=coop timeout 600 python bench.py --steps 5 --warmup 3 --no-sssp --no-c5 --no-extras --no-cpu-baseline --callers 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],3), d['e2e'].get('concurrent_callers'))"

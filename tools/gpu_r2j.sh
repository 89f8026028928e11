#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in 8192 24576; do
  echo "== callers=3 (ws) pinned pool $mb MB"; B200_PINNED_POOL_MB=$mb timeout 600 python bench.py --steps 5 --warmup 3 --no-sssp --no-c5 --no-extras --no-cpu-baseline --callers 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['e2e'].get('concurrent_callers'))"
done
echo "== C5 alone"; timeout 600 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_batch'], d['value'])"
echo "== C5 probe with kernel timeline"; B200_COOP_TRACE=1 B200_BATCH_TRACE=1 timeout 600 python tools/c5_probe.py 2>&1 | grep "^\[ws\]\|^\[trim\]\|^\[batch\]\|^call" | tail -5 | cut -c1-330

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "batch or c5 or C5 or packed" 2>&1 | tail -3
echo "== bench --workload C5 with batch trace"
B200_BATCH_TRACE=1 timeout 900 python bench.py --workload C5 --no-cpu-baseline > gpurun_out/r2s_c5.json 2> gpurun_out/r2s_c5.err
grep -E "^\[split\]|^\[batch\]" gpurun_out/r2s_c5.err | tail -6 | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_c5.json').read().strip().split('\n')[-1])
c=d.get('c5', d)
print({k:c[k] for k in ('ms_per_batch','compose_batch_packed_wall_ms_per_call','inside_the_call_ms_per_step') if k in c})
PY
echo "== C5 kernel timeline"
B200_COOP_TRACE=1 timeout 600 python tools/c5_probe.py 2>&1 | grep -E "ws\]|coop\]|trace" | tail -4 | cut -c1-500

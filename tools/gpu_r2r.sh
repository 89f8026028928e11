#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench --workload C5 with batch trace"
B200_BATCH_TRACE=1 timeout 900 python bench.py --workload C5 --no-cpu-baseline > gpurun_out/r2r_c5.json 2> gpurun_out/r2r_c5.err
grep -E "^\[split\]|^\[batch\]" gpurun_out/r2r_c5.err | tail -4 | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_c5.json').read().strip().split('\n')[-1])
c=d.get('c5', d)
print({k:c[k] for k in ('ms_per_batch','compose_batch_packed_wall_ms_per_call','inside_the_call_ms_per_step') if k in c})
PY
echo "== default bench (all legs) with batch trace"
B200_BATCH_TRACE=1 timeout 1200 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2r_all.json 2> gpurun_out/r2r_all.err
grep -E "^\[split\]|^\[batch\]" gpurun_out/r2r_all.err | tail -4 | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_all.json').read().strip().split('\n')[-1])
c=d['c5']
print({k:c[k] for k in ('ms_per_batch','compose_batch_packed_wall_ms_per_call','inside_the_call_ms_per_step') if k in c})
PY

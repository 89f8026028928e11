#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_sweep_sssp.py tests/test_gpu_parity.py -x -q -m gpu -k "sweep or sssp or shortest" 2>&1 | tail -2
echo "== C4 composed-lattice props"; timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 3 2>&1 | tail -1 | cut -c1-330
echo "== C4 sorted"; timeout 600 python tools/profile_run.py --no-compose --sssp --reps 3 2>&1 | tail -1 | cut -c1-330
echo "== window"; timeout 600 python tools/profile_run.py --no-compose --sssp-window 1000 --reps 3 2>&1 | tail -1 | cut -c1-330
timeout 900 python bench.py --no-c5 --no-extras --no-cpu-baseline --no-sssp > gpurun_out/r2z_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('concurrent_callers'))
PY

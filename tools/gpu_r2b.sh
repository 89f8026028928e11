#!/bin/bash
# parity of the default build, then timing of the default build and of every variant library given as argument
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2b_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r2b_parity.log
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=rustfst_b200/librustfst_b200_$v.so
  echo "== variant '${v:-default}'"
  B200_LIB=$lib B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2b_${v:-default}.log 2>&1; echo "rc=$?"
  grep "^\[ws\]" gpurun_out/r2b_${v:-default}.log | tail -1
  tail -1 gpurun_out/r2b_${v:-default}.log | python -c "
import sys,ast
for l in sys.stdin:
    try:
        d=ast.literal_eval(l.strip()); print({k:round(v,3) if isinstance(v,float) else v for k,v in d.items() if k.startswith('ms_') or k=='arcs_emitted'})
    except Exception as e: print('unparsable:', l[:200])"
done

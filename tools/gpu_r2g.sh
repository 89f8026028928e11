#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=rustfst_b200/librustfst_b200_$v.so
  echo "== variant '${v:-default}'"
  B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
  B200_LIB=$lib B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2g_${v:-default}.log 2>&1; echo "rc=$?"
  grep "^\[ws\]\|^\[trim\]" gpurun_out/r2g_${v:-default}.log | tail -2
  tail -1 gpurun_out/r2g_${v:-default}.log | cut -c90-260
done

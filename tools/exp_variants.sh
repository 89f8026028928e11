#!/bin/bash
# A/B of experimental library builds on the C3 compose.  Build a variant first, e.g.
#   B200_LIB_SUFFIX=t512 B200_EXTRA_NVCC_FLAGS="-DB200_COOP_THREADS=512" python rustfst_b200/build.py
# then `bash tools/exp_variants.sh t512` times the default library and librustfst_b200_t512.so back to back
# (rustfst_b200/ffi.py loads the library named by B200_LIB).  Check parity of a variant before timing it:
#   B200_LIB=rustfst_b200/librustfst_b200_t512.so python -m pytest tests/test_gpu_parity.py -q
cd "$(dirname "$0")/.."
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=rustfst_b200/librustfst_b200_$v.so
  echo "== variant '${v:-default}'"
  B200_LIB=$lib python tools/profile_run.py --reps 4 2>&1 | tail -2 | python -c "
import sys,ast
for l in sys.stdin:
    d=ast.literal_eval(l.strip()); print({k:round(v,3) if isinstance(v,float) else v for k,v in d.items() if k.startswith('ms_')})"
done

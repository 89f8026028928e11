#!/bin/bash
# A/B of experimental library builds (see rustfst_b200/build.py: B200_LIB_SUFFIX) on the C3 compose.
cd "$(dirname "$0")/.."
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=rustfst_b200/librustfst_b200_$v.so
  echo "== variant '${v:-default}'"
  B200_LIB=$lib python tools/profile_run.py --reps 4 2>&1 | tail -2 | python -c "
import sys,ast
for l in sys.stdin:
    d=ast.literal_eval(l.strip()); print({k:round(v,3) if isinstance(v,float) else v for k,v in d.items() if k.startswith('ms_')})"
done

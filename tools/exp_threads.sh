#!/bin/bash
# experiment: threads per CTA of the persistent compose kernel (build variants librustfst_b200_t512/_t1024.so)
cd "$(dirname "$0")/.."
run() { echo "== lib=$1 minblocks=$2"; B200_LIB=$1 B200_COOP_MINBLOCKS=$2 python tools/profile_run.py --reps 4 2>&1 | tail -2; }
run "" 3
run "" 2
run rustfst_b200/librustfst_b200_t512.so 1
run rustfst_b200/librustfst_b200_t512.so 2
run rustfst_b200/librustfst_b200_t512.so 3
run rustfst_b200/librustfst_b200_t1024.so 1

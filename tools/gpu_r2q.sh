#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== dag order trace x6"
B200_COOP_TRACE=1 timeout 600 python tools/profile_run.py --no-compose --sssp-top --reps 6 2>&1 | grep -E "dag-order\] ms|ms_order" | cut -c1-40,200-330
bash tools/gpu_full.sh ref

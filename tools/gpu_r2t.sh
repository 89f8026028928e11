#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_NSHORTEST_TRACE=1 timeout 1200 python -m pytest tests/test_gpu_nshortest.py -x -q -m gpu --durations=5 2>&1 | tail -25 | cut -c1-400

#!/bin/bash
cd "$(dirname "$0")/.."
for g in 148 96 74 48 37 24; do
echo "== C5 kernel timeline, grid $g"
B200_WS_GRID=$g B200_COOP_TRACE=1 timeout 600 python tools/c5_probe.py 2>&1 | grep -E "ws\]" | tail -1 | cut -c1-500
done

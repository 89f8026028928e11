#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/callers_probe.py 1 2>&1 | tail -4
timeout 600 python tools/callers_probe.py 2 2>&1 | tail -13

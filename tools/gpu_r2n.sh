#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dag_order.py -x -q 2>&1 | tail -2
echo "== sssp leg (two-phase levels)"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-extras --no-cpu-baseline --callers 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); s=d['sssp']; print('compose ms', round(d['ms_per_step'],3), 'sssp ms', round(s['ms_per_step'],3), 'order', round(s['ms_order_device_per_step'],3), 'relax', round(s['ms_relax_and_backtrace_per_step'],3))"
echo "== C5 probe, no trace"; timeout 600 python tools/c5_probe.py 2>&1 | grep "^call" | tail -3 | cut -c1-300
echo "== C5 probe, prebuilt handle array, no trace"; timeout 600 python - <<'PY' 2>&1 | tail -4
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import rustfst_b200 as R
from rustfst_b200 import synth
t = synth.random_graph_transducer(500_000, 5_000_000, 5000, seed=5)
rng = np.random.default_rng(77)
t["finals"] = np.where(rng.random(500_000) < 0.5, rng.integers(0, 640, size=500_000) / 64.0, np.inf).astype(np.float32)
dt = R.DeviceFst.upload(synth.to_vector_fst(t))
labels = synth.sample_path_labels_batch(t, 200, 8192, seed=100)
accs = R.AcceptorBatch([synth.to_vector_fst(synth.linear_acceptor(labels[i], seed=100 + i)) for i in range(8192)])
pb = None
for i in range(8):
    t0 = time.perf_counter()
    pb, st = R.compose_batch_packed(accs, device_transducer=dt)
    t1 = time.perf_counter()
    print(f"call {i}: wall {1e3*(t1-t0):.2f} ms inside: h2d {st['ms_h2d']:.2f} expand {st['ms_expand']:.2f} connect {st['ms_connect']:.2f} split+d2h {st['ms_d2h']:.2f}")
PY

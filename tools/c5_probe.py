"""C5 (batched compose) on one GPU, step by step: where the time of one b200_compose_batch_packed call goes
(B200_BATCH_TRACE=1 prints the library's own breakdown)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustfst_b200 as R
from rustfst_b200 import synth

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t = synth.random_graph_transducer(500_000, 5_000_000, 5000, seed=5)
rng = np.random.default_rng(77)
t["finals"] = np.where(rng.random(500_000) < 0.5, rng.integers(0, 640, size=500_000) / 64.0, np.inf).astype(np.float32)
ht = synth.to_vector_fst(t)
dt = R.DeviceFst.upload(ht)
labels = synth.sample_path_labels_batch(t, 200, batch, seed=100)
accs = [synth.to_vector_fst(synth.linear_acceptor(labels[i], seed=100 + i)) for i in range(batch)]
for i in range(6):
    t0 = time.perf_counter()
    pb, st = R.compose_batch_packed(accs, device_transducer=dt)
    t1 = time.perf_counter()
    buf = pb.to_numpy()
    t2 = time.perf_counter()
    print(f"call {i}: compose_batch_packed {1e3*(t1-t0):.2f} ms (h2d {st['ms_h2d']:.2f} expand {st['ms_expand']:.2f} "
          f"connect {st['ms_connect']:.2f} split+d2h {st['ms_d2h']:.2f}, waves {st['waves']}), serialise {1e3*(t2-t1):.2f} ms, "
          f"{pb.info()}")

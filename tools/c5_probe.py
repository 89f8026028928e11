import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustfst_b200 as R
from rustfst_b200 import synth
n_t, a_t, B = 500_000, 5_000_000, 8192
t = synth.random_graph_transducer(n_t, a_t, 5000, seed=5)
rng = np.random.default_rng(77)
t["finals"] = np.where(rng.random(n_t) < 0.5, rng.integers(0, 640, size=n_t) / 64.0, np.inf).astype(np.float32)
ht = synth.to_vector_fst(t)
t0 = time.perf_counter()
accs = [synth.to_vector_fst(synth.linear_acceptor(synth.sample_path_labels(t, 200, seed=100 + i), seed=100 + i)) for i in range(B)]
print("build acceptors", time.perf_counter() - t0)
for i in range(4):
    t0 = time.perf_counter()
    res, st = R.compose_batch(accs, ht)
    t1 = time.perf_counter()
    print(f"batch {i}: wall {1e3*(t1-t0):.1f} ms  h2d {st['ms_h2d']:.1f} expand {st['ms_expand']:.1f} (kernel {st['ms_emit_kernel']:.1f}) connect {st['ms_connect']:.1f} d2h {st['ms_d2h']:.1f} waves {st['waves']} arcs_out {st['arcs_out']} states_exp {st['states_expanded']}")
    t2 = time.perf_counter(); del res; print(f"   free {1e3*(time.perf_counter()-t2):.1f} ms")

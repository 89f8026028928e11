import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfst_b200 as R
from rustfst_b200 import synth
a1 = synth.layered_acceptor(1_000_000, 10_000_000, 32, 3, 50)
a2 = synth.bigram_transducer(1_000_000, 10_000_000, 32, 4, 50, out_vocab=20000)
h1, h2 = synth.to_vector_fst(a1), synth.to_vector_fst(a2)
for i in range(5):
    t0 = time.perf_counter()
    res, st = R.compose_with_stats(h1, h2)
    t1 = time.perf_counter()
    print(f"call {i}: wall {1e3*(t1-t0):.1f} ms  h2d {st['ms_h2d']:.1f}  expand {st['ms_expand']:.1f}  connect {st['ms_connect']:.1f}  d2h {st['ms_d2h']:.1f}")
    t2 = time.perf_counter(); del res; print(f"   free {1e3*(time.perf_counter()-t2):.1f} ms")
g = synth.layered_acceptor(5_000_000, 50_000_000, 1000, 6, 50)
hg = synth.to_vector_fst(g)
for i in range(3):
    t0 = time.perf_counter(); sp, st = R.shortestpath_with_stats(hg); t1 = time.perf_counter()
    print(f"sssp call {i}: wall {1e3*(t1-t0):.1f} ms h2d {st['ms_h2d']:.1f} device {st['ms_device']:.1f} relax {st['ms_relax_kernel']:.2f} plan {st['ms_queue_plan_host']:.2f}")

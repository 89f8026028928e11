"""Two (or N) host threads calling fst_compose on the same host operands: per-call breakdown under concurrency."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfst_b200 as R
from rustfst_b200 import synth
n_callers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a1 = synth.layered_acceptor(1_000_000, 10_000_000, 32, 3, 50)
a2 = synth.bigram_transducer(1_000_000, 10_000_000, 32, 4, 50, out_vocab=20000)
h1, h2 = synth.to_vector_fst(a1), synth.to_vector_fst(a2)
log = [[] for _ in range(n_callers)]
def worker(k, n):
    for _ in range(n):
        t0 = time.perf_counter()
        r, s = R.compose_with_stats(h1, h2)
        t1 = time.perf_counter()
        del r
        t2 = time.perf_counter()
        log[k].append((1e3 * (t1 - t0), s["ms_h2d"], s["ms_expand"] + s["ms_connect"], s["ms_d2h"], 1e3 * (t2 - t1)))
def run(n):
    th = [threading.Thread(target=worker, args=(k, n)) for k in range(n_callers)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    return 1e3 * (time.perf_counter() - t0)
run(3)
for l in log: l.clear()
wall = run(6)
print(f"{n_callers} callers x 6 calls: wall {wall:.1f} ms -> {wall / (6 * n_callers):.2f} ms per compose")
for k, l in enumerate(log):
    for c in l:
        print(f"  caller {k}: call {c[0]:.1f} ms (h2d {c[1]:.1f}, kernels {c[2]:.1f}, d2h {c[3]:.1f}), free {c[4]:.1f}")

"""n > 1 shortest paths on the C4 lattice: wall time of fst_shortest_path_with_config on a host handle and where it goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfst_b200 as R
from rustfst_b200 import synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = synth.workload("C4", scale=scale)
hg = synth.to_vector_fst(g)
for n in (1, 2, 10, 100):
    cfg = R.ShortestPathConfig(nshortest=n)
    for i in range(3):
        t0 = time.perf_counter(); sp, st = R.shortestpath_with_stats(hg, cfg); t1 = time.perf_counter()
    print(f"n={n}: wall {1e3*(t1-t0):.1f} ms  h2d {st['ms_h2d']:.1f}  device/total {st['ms_device']:.1f}  relax {st['ms_relax_kernel']:.2f}  "
          f"path {st['path']}  result states {sp.num_states()}", flush=True)

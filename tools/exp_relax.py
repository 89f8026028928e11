"""Shape sweep of the persistent relaxation kernel on the C4 lattice (knobs are read per call by sssp.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfst_b200 as R
from rustfst_b200 import synth
g = synth.workload("C4")
dg = R.DeviceFst.upload(synth.to_vector_fst(g))
ref = None
for threads, per_sm, lanes, pre, ilp in [(1024, 0, 8, 0, 1), (1024, 0, 8, 0, 2), (1024, 0, 8, 0, 4), (512, 0, 8, 0, 2), (512, 0, 8, 0, 4),
                                         (256, 0, 8, 0, 2), (1024, 0, 16, 0, 1), (1024, 0, 16, 0, 2), (1024, 0, 4, 0, 2), (1024, 0, 8, 1, 2),
                                         (256, 0, 8, 1, 1)]:
    os.environ["B200_RELAX_ILP"] = str(ilp)
    os.environ["B200_RELAX_THREADS"] = str(threads)
    os.environ["B200_RELAX_LANES"] = str(lanes)
    os.environ["B200_RELAX_PRETEST"] = str(pre)
    if per_sm: os.environ["B200_RELAX_CTAS_PER_SM"] = str(per_sm)
    else: os.environ.pop("B200_RELAX_CTAS_PER_SM", None)
    best = None
    for _ in range(4):
        sp, st = R.device_shortest_path(dg)
        best = st if best is None or st["ms_relax_kernel"] < best["ms_relax_kernel"] else best
    b = sp.to_bytes()
    if ref is None: ref = b
    print(f"threads {threads:4d} ctas/sm {per_sm or 'max'} lanes {lanes} pretest {pre} ilp {ilp}: relax {best['ms_relax_kernel']:.3f} ms, call {best['ms_device']:.3f} ms, "
          f"waves {best['waves']}, same result {b == ref}", flush=True)

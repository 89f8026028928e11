"""Minimal driver for ncu captures: one device-resident C3 compose (+ optionally one C4 shortest path).

    ncu ... python tools/profile_run.py [--scale S] [--sssp] [--reps R]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import rustfst_b200 as R  # noqa: E402
from rustfst_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--levels", type=int, default=50)
ap.add_argument("--sssp", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--fanout", action="store_true")
ap.add_argument("--vocab", type=int, default=32)
ap.add_argument("--no-compose", action="store_true")
ap.add_argument("--spread", action="store_true", help="transducer targets depend on (label, source): searched side HBM-resident")
ap.add_argument("--sssp-top", action="store_true", help="C4 with the property word compose leaves (device DFS order)")
ap.add_argument("--sssp-window", type=int, default=0, help="SSSP on a window DAG (skip-level arcs) with this window")
args = ap.parse_args()
if args.spread and args.vocab == 32 and not args.fanout:  # the bench's compose_spread shape: with 32 labels the product explodes
    args.vocab, args.fanout = 93, True

n, a = int(1_000_000 * args.scale), int(10_000_000 * args.scale)
a1 = synth.layered_acceptor(n, a, args.vocab, 3, args.levels, start_fanout=args.fanout)
a2 = synth.bigram_transducer(n, a, args.vocab, 4, args.levels, out_vocab=20000, start_fanout=args.fanout, spread=args.spread)
d1, d2 = R.DeviceFst.upload(synth.to_vector_fst(a1)), R.DeviceFst.upload(synth.to_vector_fst(a2))
out = None
for _ in range(0 if args.no_compose else args.reps):
    del out  # release the previous result first: the stream-ordered pool then reuses its blocks
    out, st = R.device_compose(d1, d2)
    print({k: st[k] for k in ("states_expanded", "arcs_emitted", "waves", "kernel_launches", "ms_expand", "ms_connect",
                              "ms_emit_kernel", "ms_phase_match", "ms_phase_emit", "ms_phase_rank", "ms_phase_resolve")})
if args.sssp or args.sssp_top or args.sssp_window:
    from rustfst_b200 import props as PR
    if args.sssp_window:
        g = synth.window_dag(int(5_000_000 * args.scale), int(50_000_000 * args.scale), 1000, 6, window=args.sssp_window)
    else:
        g = synth.layered_acceptor(int(5_000_000 * args.scale), int(50_000_000 * args.scale), 1000, 6, 50)
    if args.sssp_top:
        g = dict(g, props=g["props"] & ~(PR.TOP_SORTED | PR.NOT_TOP_SORTED))
    dg = R.DeviceFst.upload(synth.to_vector_fst(g))
    for _ in range(args.reps):
        sp, sst = R.device_shortest_path(dg)
        print({k: sst[k] for k in ("arcs_relaxed", "waves", "kernel_launches", "ms_device", "ms_relax_kernel", "path",
                                   "ms_order_device", "order_on_device", "queue_kind")})

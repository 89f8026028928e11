#!/usr/bin/env python
"""bench.py — the driver-facing benchmark of the B200 compose + shortest-path engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3|C2] [--scale S]

One "step" = one pass of the hot path over one synthetic workload:
  * compose  : BASELINE.json configs[2] "C3" — 1M-state/10M-arc acceptor o 1M-state/10M-arc transducer
               (TropicalWeight, AutoFilter, connect=true) — the configuration north_star's target is quoted on;
  * sssp     : BASELINE.json configs[3] "C4" — shortest_path(n=1) on a 5M-state/50M-arc acyclic lattice.
The JSON line's `metric`/`value` is the composed-arcs/s of the compose leg (whole job, inputs resident in HBM);
the SSSP leg is reported in the `sssp` object of the same line.  N > 1 runs one replica per GPU with different
seeds (weak scaling, no data-path collective: a single compose does not shard — DESIGN.md §5); NCCL is only used
for the barrier / max-over-ranks reduction of the timings.

`--impl reference` times the CPU oracle (oracle/liboracle.so, the restated reference algorithm; the Rust
reference itself cannot be built here) on the box's host cores — it is single-threaded like rustfst.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every 10 ms (the timed region of the
    default run is ~50 ms), falling back to one nvidia-smi query per 200 ms when pynvml is unavailable."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.source = None
        self._stop_evt = threading.Event()

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.gpu)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        bits = {"hw_slowdown": getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        self.source = "nvml"
        while not self._stop_evt.is_set():
            self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
            r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self._stop_evt.wait(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": self.source, "window": "compose warm-up + timed region"}


def load_synth():
    """rustfst_b200/synth.py loaded by file path: the generators are numpy-only, and the reference arm must not map the
    product library into its process (importing the package would dlopen it)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_synth", os.path.join(ROOT, "rustfst_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def gen_compose_workload(name, scale, rank, synth=None):
    """Operands of the compose leg.  Every replica composes machines of the SAME shape (generator seeds 3 / 4: the same
    states, labels and targets, hence the same product size on every rank, so per-GPU work is fixed as N grows); the
    arc and final weights are re-drawn per rank, so the data differ.  (Round 1 seeded the whole generator per rank; the
    start states of ranks 1, 3 and 6 then shared no label and those replicas composed an empty product.)"""
    synth = synth or load_synth()
    if name == "C2":
        n, a, lv, v, ov = int(100_000 * scale), int(1_000_000 * scale), 25, 32, 2000
    else:
        n, a, lv, v, ov = int(1_000_000 * scale), int(10_000_000 * scale), 50, 32, 20000
    a1 = synth.layered_acceptor(n, a, v, 3, lv)
    a2 = synth.bigram_transducer(n, a, v, 4, lv, out_vocab=ov)
    if rank:
        rng = np.random.default_rng(1000 + rank)
        for d in (a1, a2):
            d["arcs"]["weight"] = rng.integers(0, 640, size=len(d["arcs"])).astype(np.float32) / np.float32(64.0)
            fin = np.isfinite(d["finals"])
            d["finals"][fin] = rng.integers(0, 640, size=int(fin.sum())).astype(np.float32) / np.float32(64.0)
    return a1, a2


def gen_sssp_workload(scale, rank, synth=None):
    synth = synth or load_synth()
    return synth.layered_acceptor(int(5_000_000 * scale), int(50_000_000 * scale), 1000, 6 + 100 * rank, 50)


def compose_config(workload, scale, world, a1, a2):
    """`config` of the JSON line — identical for the b200 arm and the reference arm (the driver compares them)."""
    return {"workload": f"{workload}: layered acyclic acceptor ({a1['num_states']} states, {len(a1['arcs'])} arcs, "
                        f"olabel-sorted) o bigram-structured transducer ({a2['num_states']} states, "
                        f"{len(a2['arcs'])} arcs, ilabel-sorted), TropicalWeight, AutoFilter, connect=true",
            "scale": scale, "replicas": world, "parallelism": f"replicas x{world} (no data-path collective)",
            "l2": "operands + table + output (>= 0.8 GB) exceed the 126 MB L2; no flush needed"}


def csr_bytes(d):
    return int(d["offsets"].nbytes + d["arcs"].nbytes + d["finals"].nbytes)


def run_reference(args, rank, world):
    """CPU arm: the oracle port of rustfst's compose on the host cores (1 thread: the reference is single-threaded),
    on the SAME full-size operands as the b200 arm (rank 0's), every warm-up and every step a full compose + connect
    (about 10 s each).  Nothing of the product is imported: the generators are loaded by path, the oracle through
    tests/oracle_lib.py."""
    if rank != 0:
        return
    from tests import oracle_lib as O
    workload = "C3" if args.workload == "C5" else args.workload
    a1, a2 = gen_compose_workload(workload, args.scale, 0)
    oa = O.OFst.from_csr(a1["offsets"].astype(np.uint64), a1["arcs"], a1["finals"], a1["start"], a1["props"])
    ob = O.OFst.from_csr(a2["offsets"].astype(np.uint64), a2["arcs"], a2["finals"], a2["start"], a2["props"])
    arcs = 0
    for _ in range(args.warmup):
        O.compose(oa, ob)
    t = 0.0
    for _ in range(args.steps):
        _, st = O.compose(oa, ob, want_stats=True)
        t += st["seconds"]
        arcs += st["arcs_emitted"]
    value = arcs / t
    sample = (f"the full {workload} workload ({a1['num_states']} x {a2['num_states']} states, {len(a1['arcs'])} + "
              f"{len(a2['arcs'])} arcs), one complete compose+connect per step ({arcs // max(1, args.steps)} arcs emitted), "
              f"{args.warmup} warm-up passes, oracle port of rustfst's algorithm (C++ -O3), 1 thread: the reference is "
              "single-threaded")
    line = {
        "impl": "reference", "metric": "composed_arcs_per_sec", "value": value, "unit": "arcs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": compose_config(workload, args.scale, args.gpus, a1, a2),
        "cpu_baseline": {"value": value, "unit": "arcs/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "arcs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_c5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE.json configs[4]: `--batch` linear acceptors (200 arcs each, label strings sampled as walks of T)
    against one shared 500K-state/5M-arc transducer.  Acceptors are block-sharded over the ranks, T is replicated,
    every rank runs ONE device BFS for its whole shard (b200_compose_batch), and the result FSTs are gathered on
    rank 0 over NCCL (strong scaling: the batch is fixed)."""
    import torch
    import rustfst_b200 as R
    from rustfst_b200 import synth
    from rustfst_b200.parallel import gather_blobs, shard_range
    n_t, a_t = int(500_000 * args.scale), int(5_000_000 * args.scale)
    t = synth.random_graph_transducer(n_t, a_t, 5000, seed=5)
    rng = np.random.default_rng(77)
    t["finals"] = np.where(rng.random(n_t) < 0.5, rng.integers(0, 640, size=n_t) / 64.0, np.inf).astype(np.float32)
    ht = synth.to_vector_fst(t)
    lo, hi = shard_range(args.batch, rank, world)
    accs = [synth.to_vector_fst(synth.linear_acceptor(synth.sample_path_labels(t, 200, seed=100 + i), seed=100 + i))
            for i in range(lo, hi)]

    def step(gather):
        results, st = R.compose_batch(accs, ht)
        blobs = None
        if gather and dist is not None:
            packed = []
            for r in results:
                o, a, f, _ = r.to_csr()
                packed.append(o.tobytes() + a.tobytes() + f.tobytes())
            blobs = gather_blobs(packed, dist, device=torch.device("cuda", local_rank))
        return results, st, blobs

    for _ in range(args.warmup):
        step(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    arcs, launches, waves = 0, 0, 0
    for _ in range(args.steps):
        results, st, blobs = step(True)
        arcs += st["arcs_out"]; launches += st["kernel_launches"]; waves += st["waves"]
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    total_arcs = sum_over_ranks(float(arcs))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib as O
        ot = O.OFst.from_csr(t["offsets"].astype(np.uint64), t["arcs"], t["finals"], t["start"], t["props"])
        n_sample = min(args.batch, 2048)
        oaccs = []
        for i in range(n_sample):
            d = synth.linear_acceptor(synth.sample_path_labels(t, 200, seed=100 + i), seed=100 + i)
            oaccs.append(O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"]))
        secs, oarcs = 0.0, 0
        for oa in oaccs:
            r, ost = O.compose(oa, ot, want_stats=True)
            secs += ost["seconds"]; oarcs += r.num_trs
        cpu = {"value": oarcs / secs, "unit": "arcs/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"the first {n_sample} acceptors of the batch composed one by one with the same transducer "
                         f"({oarcs} result arcs in {secs:.2f} s), oracle port, 1 thread"}
    if rank == 0:
        line = {"metric": "composed_arcs_per_sec", "value": total_arcs / (ms * 1e-3), "unit": "arcs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C5: {args.batch} linear acceptors (200 arcs) o {n_t}-state/{len(t['arcs'])}-arc "
                                       "transducer, sharded by acceptor, results gathered on rank 0 over NCCL",
                           "parallelism": f"acceptor shards x{world}, transducer replicated",
                           "waves_per_step": waves // max(1, args.steps)},
                "e2e": {"value": total_arcs / (ms * 1e-3), "unit": "arcs/s", "api": "b200_compose_batch on host handles",
                        "h2d_bytes_per_step": int(csr_bytes(t) + (hi - lo) * (201 * 8 + 200 * 16 + 4)),
                        "d2h_bytes_per_step": int(arcs // max(1, args.steps) * 16)},
                "cpu_baseline": cpu, "gpu_launches": int(launches)}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C2", "C5"])
    ap.add_argument("--batch", type=int, default=8192, help="C5: number of linear acceptors")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (tests only; full size = 1.0)")
    ap.add_argument("--callers", type=int, default=3, help="host threads of the supplementary concurrent e2e figure (1 = skip)")
    ap.add_argument("--no-sssp", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import rustfst_b200 as R
    from rustfst_b200.ffi import check_ffi_error, lib
    from rustfst_b200 import synth

    if not torch.cuda.is_available() or R.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    check_ffi_error(lib.b200_set_device(local_rank), "b200_set_device")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        check_ffi_error(lib.b200_device_synchronize(), "sync")

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak_gbs, peak_src = measured_peak_gbs()

    if args.workload == "C5":
        run_c5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks)
        if dist is not None:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ compose leg
    a1, a2 = gen_compose_workload(args.workload, args.scale, rank, synth)
    h1, h2 = synth.to_vector_fst(a1), synth.to_vector_fst(a2)
    d1, d2 = R.DeviceFst.upload(h1), R.DeviceFst.upload(h2)  # inputs resident in HBM before the timed region
    sampler = ClockSampler(local_rank)  # samples every 200 ms from the warm-up on: the timed region is ~50 ms
    sampler.start()
    for _ in range(args.warmup):
        out, st = R.device_compose(d1, d2)
        del out
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    tot = {"arcs_emitted": 0, "arcs_iterated": 0, "states_expanded": 0, "kernel_launches": 0, "emit_launches": 0,
           "ms_emit_kernel": 0.0, "ms_expand": 0.0, "ms_connect": 0.0, "waves": 0, "ms_phase_match": 0.0,
           "ms_phase_emit": 0.0, "ms_phase_rank": 0.0, "ms_phase_resolve": 0.0}
    for _ in range(args.steps):
        out, st = R.device_compose(d1, d2)  # blocks until the result is complete in HBM
        for k in tot:
            tot[k] += st[k]
        del out
    ev1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    ms_region = max_over_ranks(dev_ms)
    arcs_all = sum_over_ranks(float(tot["arcs_emitted"]))
    value = arcs_all / (ms_region * 1e-3)
    steps = args.steps
    bytes_compose = 16.0 * tot["arcs_iterated"] + 48.0 * tot["arcs_emitted"] + 32.0 * tot["states_expanded"]
    # Dominant kernel = the persistent BFS kernel k_compose_coop (one launch per compose): algorithmic bytes of the
    # whole expansion (16*A_it + 48*A_out + 32*S) over its CUDA-event duration.  The arc-scan phase of
    # that kernel (phase B: gather matched arc, table probe, write output arc = 48 B/arc) is timed inside the kernel
    # with %globaltimer and reported separately.
    kern_gbs = bytes_compose / (tot["ms_emit_kernel"] * 1e-3) / 1e9 if tot["ms_emit_kernel"] > 0 else 0.0
    scan_gbs = 48.0 * tot["arcs_emitted"] / (tot["ms_phase_emit"] * 1e-3) / 1e9 if tot["ms_phase_emit"] > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "compose_traffic.json")
    if os.path.exists(tpath) and args.workload == "C3" and args.scale == 1.0:  # the capture is of the full-size C3 kernel
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": "k_compose_coop (persistent cooperative kernel: the whole BFS expansion)",
        "achieved": kern_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": kern_gbs / peak_gbs, "peak_source": peak_src,
        "traffic": traffic,
        "algorithmic_bytes_per_launch": bytes_compose / max(1, tot["emit_launches"]),
        "bytes_model": "16*A_it + 48*A_out + 32*S",
        "avg_launch_ms": tot["ms_emit_kernel"] / max(1, tot["emit_launches"]),
        "launches": tot["emit_launches"],
        "arc_scan_phase": {"bytes_model": "48*A_out", "ms_per_step": tot["ms_phase_emit"] / steps,
                           "achieved_GBps": scan_gbs, "frac": (scan_gbs / peak_gbs) if scan_gbs else None,
                           "timer": "%globaltimer inside the kernel"},
        "phase_ms_per_step": {k[9:]: tot[k] / steps for k in ("ms_phase_match", "ms_phase_emit", "ms_phase_rank",
                                                               "ms_phase_resolve")},
        "whole_compose": {"device_ms_per_step": (tot["ms_expand"] + tot["ms_connect"]) / steps,
                          "expand_ms_per_step": tot["ms_expand"] / steps, "connect_ms_per_step": tot["ms_connect"] / steps,
                          "achieved_GBps": bytes_compose / ((tot["ms_expand"] + tot["ms_connect"]) * 1e-3) / 1e9,
                          "frac": bytes_compose / ((tot["ms_expand"] + tot["ms_connect"]) * 1e-3) / 1e9 / peak_gbs},
    }

    # ------------------------------------------------------------------ end-to-end through the C-ABI (host buffers)
    for _ in range(args.warmup):  # first calls populate the page-locked host pool that result handles come from
        res, st = R.compose_with_stats(h1, h2)
        del res
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_arcs, d2h = 0, 0
    for _ in range(steps):
        res, st = R.compose_with_stats(h1, h2)  # fst_compose path: H2D of both operands, kernels, D2H of the result
        e2e_arcs += st["arcs_emitted"]
        d2h = 4 * (res.num_states() + 1) + 16 * res.num_trs_total() + 4 * res.num_states()
        del res
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": sum_over_ranks(float(e2e_arcs)) / (e2e_ms * 1e-3), "unit": "arcs/s",
           "h2d_bytes_per_step": csr_bytes(a1) + csr_bytes(a2), "d2h_bytes_per_step": int(d2h),
           "api": "fst_compose (b200_compose_with_stats) on host VectorFst handles"}
    # Same calls issued by several host threads at once (the C-ABI is re-entrant, every call owns a stream): the PCIe
    # link is full duplex, so one caller's upload overlaps another's kernels and download.  Supplementary figure; the
    # headline `value` above is the single-caller number.
    if args.callers > 1:
        import threading
        done = [0] * args.callers

        def worker(k):
            for _ in range(steps):
                r, s = R.compose_with_stats(h1, h2)
                done[k] += s["arcs_emitted"]
                del r
        barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(k,)) for k in range(args.callers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        check_ffi_error(lib.b200_device_synchronize(), "sync")
        dt = time.perf_counter() - t0
        e2e["concurrent_callers"] = {"callers": args.callers, "value": float(sum(done)) / dt, "unit": "arcs/s",
                                     "ms_per_compose": dt * 1e3 / (steps * args.callers), "timer": "host wall clock"}
    del d1, d2

    # ------------------------------------------------------------------ SSSP leg (C4)
    sssp = None
    if not args.no_sssp:
        g = gen_sssp_workload(args.scale, rank, synth)
        hg = synth.to_vector_fst(g)
        dg = R.DeviceFst.upload(hg)
        for _ in range(args.warmup):
            R.device_shortest_path(dg)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        stot = {"arcs_relaxed": 0, "states_settled": 0, "ms_relax_kernel": 0.0, "ms_device": 0.0, "relax_launches": 0,
                "kernel_launches": 0}
        path_kind = None
        for _ in range(steps):
            sp, sst = R.device_shortest_path(dg)
            for k in stot:
                stot[k] += sst[k]
            path_kind = sst["path"]
        s1.record()
        barrier()
        sssp_ms = max_over_ranks(s0.elapsed_time(s1))
        edges = sum_over_ranks(float(stot["arcs_relaxed"]))
        n_states = g["num_states"]
        sssp_bytes = 16.0 * stot["arcs_relaxed"] + 20.0 * n_states * steps
        relax_gbs = (16.0 * stot["arcs_relaxed"] + 12.0 * stot["states_settled"]) / (stot["ms_relax_kernel"] * 1e-3) / 1e9
        sssp = {
            "metric": "sssp_edges_per_sec", "value": edges / (sssp_ms * 1e-3), "unit": "edges/s",
            "ms_per_step": sssp_ms / steps, "workload": f"C4: shortest_path(n=1), {n_states}-state/{len(g['arcs'])}-arc "
            "layered acyclic lattice, TOP_SORTED known (StateOrderQueue), dyadic weights",
            "device_path": "parallel relaxation + certificate" if path_kind == 0 else "serial replay",
            "roofline": {"bound": "hbm", "kernel": "k_relax", "achieved": relax_gbs, "peak": peak_gbs, "unit": "GB/s",
                         "frac": relax_gbs / peak_gbs, "traffic": None,
                         "whole_call": {"bytes_model": "16*E + 20*N", "achieved_GBps": sssp_bytes / (stot["ms_device"] * 1e-3) / 1e9,
                                        "frac": sssp_bytes / (stot["ms_device"] * 1e-3) / 1e9 / peak_gbs}},
            "gpu_launches": stot["kernel_launches"],
        }
        # end to end through fst_shortest_path on the host handle (H2D of the lattice inside)
        barrier()
        t0 = time.perf_counter()
        e2e_edges = 0
        for _ in range(steps):
            _, sst = R.shortestpath_with_stats(hg)
            e2e_edges += sst["arcs_relaxed"]
        barrier()
        sssp["e2e"] = {"value": e2e_edges / (time.perf_counter() - t0), "unit": "edges/s",
                       "h2d_bytes_per_step": csr_bytes(g), "d2h_bytes_per_step": 16 * 64}
        # SURVEY.md §8d asks for the lattice "with props as compose would leave them" as well: ACYCLIC known, TOP_SORTED
        # unknown -> AutoQueue picks TopOrderQueue, whose order comes from the reference's sequential DFS (run on the
        # host, exactly as the reference does; its time is reported next to the device time)
        from rustfst_b200 import props as PR
        hg_top = synth.to_vector_fst(dict(g, props=g["props"] & ~(PR.TOP_SORTED | PR.NOT_TOP_SORTED)))
        dg_top = R.DeviceFst.upload(hg_top)
        R.device_shortest_path(dg_top)
        barrier()
        t0 = time.perf_counter()
        top_steps = min(steps, 2)
        for _ in range(top_steps):
            _, tst = R.device_shortest_path(dg_top)
        barrier()
        sssp["top_order_variant"] = {
            "workload": "same lattice, ACYCLIC known / TOP_SORTED unknown (TopOrderQueue)", "queue_kind": tst["queue_kind"],
            "ms_per_call_wall": (time.perf_counter() - t0) * 1e3 / top_steps, "ms_device": tst["ms_device"],
            "ms_queue_plan_host_dfs": tst["ms_queue_plan_host"], "device_path": tst["path"]}
        del dg_top, hg_top
        # n-best on the same device-resident lattice (fst_shortest_path_with_config, nshortest = 10, unique = false):
        # forward distances + reversed machine on the device, heap search over rows fetched from HBM, device trim
        cfg10 = R.ShortestPathConfig(nshortest=10)
        R.device_shortest_path(dg, config=cfg10)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            nb, nst = R.device_shortest_path(dg, config=cfg10)
        barrier()
        sssp["nbest"] = {"nshortest": 10, "unique": False, "ms_per_call": (time.perf_counter() - t0) * 1e3 / steps,
                         "result_states": nb.num_states(), "distance_device_path": nst["path"],
                         "gpu_launches_per_call": nst["kernel_launches"], "timer": "host wall clock, lattice resident in HBM"}
        del dg

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib as O
        oa = O.OFst.from_csr(a1["offsets"].astype(np.uint64), a1["arcs"], a1["finals"], a1["start"], a1["props"])
        ob = O.OFst.from_csr(a2["offsets"].astype(np.uint64), a2["arcs"], a2["finals"], a2["start"], a2["props"])
        _, ost = O.compose(oa, ob, want_stats=True)
        cpu = {"value": ost["arcs_emitted"] / ost["seconds"], "unit": "arcs/s", "cores": 1, "kind": "port",
               "host_cores": os.cpu_count(),
               "sample": f"one full {args.workload} compose+connect of the same inputs "
                         f"({ost['arcs_emitted']} arcs in {ost['seconds']:.2f} s), oracle port of rustfst, 1 thread"}
        if sssp is not None:
            og = O.OFst.from_csr(g["offsets"].astype(np.uint64), g["arcs"], g["finals"], g["start"], g["props"])
            _, sst = O.shortest_path(og, want_stats=True)
            cpu["sssp"] = {"value": sst["arcs_relaxed"] / sst["seconds"], "unit": "edges/s",
                           "sample": f"one full C4 shortest_path ({sst['arcs_relaxed']} edges in {sst['seconds']:.2f} s)"}
            t0 = time.perf_counter()
            O.shortest_path(og, nshortest=10)
            cpu["sssp"]["nbest_ms"] = (time.perf_counter() - t0) * 1e3
            cpu["sssp"]["nbest_sample"] = "one full C4 shortest_path(nshortest=10), oracle port, 1 thread"

    if rank == 0:
        line = {
            "metric": "composed_arcs_per_sec", "value": value, "unit": "arcs/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_region / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": compose_config(args.workload, args.scale, world, a1, a2),
            "workload_stats": {"states_expanded_per_step": tot["states_expanded"] // steps,
                               "arcs_emitted_per_step": tot["arcs_emitted"] // steps,
                               "waves_per_step": tot["waves"] // steps},
            "wall_ms_per_step": wall_ms / steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "sssp": sssp,
            "gpu_launches": int(tot["kernel_launches"]), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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


def run_c5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks, synth, standalone=True):
    """BASELINE.json configs[4]: `--batch` linear acceptors (200 arcs each, label strings sampled as walks of T)
    against one shared 500K-state/5M-arc transducer.  STRONG scaling: the batch is fixed, the acceptors are block-
    sharded over the ranks, the transducer is replicated and stays resident in HBM (uploaded once, outside the timed
    region, like any shared model), every rank runs ONE device BFS for its whole shard (b200_compose_batch_packed:
    union of the shard's acceptors uploaded inside the timed region, results split per acceptor on the device and
    brought back as one block) and sends its block to rank 0 over NCCL (point-to-point; rank 0 ends up holding every
    result).  Returns the c5 object of the JSON line (rank 0) or None."""
    import torch
    import rustfst_b200 as R
    from rustfst_b200.parallel import gather_buffers, shard_range
    n_t, a_t = int(500_000 * args.scale), int(5_000_000 * args.scale)
    t = synth.random_graph_transducer(n_t, a_t, 5000, seed=5)
    rng = np.random.default_rng(77)
    t["finals"] = np.where(rng.random(n_t) < 0.5, rng.integers(0, 640, size=n_t) / 64.0, np.inf).astype(np.float32)
    ht = synth.to_vector_fst(t)
    dt = R.DeviceFst.upload(ht)
    lo, hi = shard_range(args.batch, rank, world)
    labels = synth.sample_path_labels_batch(t, 200, args.batch, seed=100)  # the same strings on every rank
    acc_dicts = [synth.linear_acceptor(labels[i], seed=100 + i) for i in range(lo, hi)]
    accs = R.AcceptorBatch([synth.to_vector_fst(d) for d in acc_dicts])  # the C-ABI's array of handles, built once
    dev = torch.device("cuda", local_rank)

    call_ms, ser_ms, xfer_ms = [], [], []
    stage = [None]  # page-locked staging buffer for the block that goes to rank 0, reused from step to step

    def step():
        tc = time.perf_counter()
        pb, st = R.compose_batch_packed(accs, device_transducer=dt)
        call_ms.append(1e3 * (time.perf_counter() - tc))
        blocks = None
        if dist is not None:
            t1 = time.perf_counter()
            nbytes = pb.info()["bytes"]
            if stage[0] is None or stage[0].numel() < nbytes:
                stage[0] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
            pb.to_numpy(out=stage[0].numpy())
            t2 = time.perf_counter()
            blocks = gather_buffers(stage[0][:nbytes], dist, device=dev)
            torch.cuda.synchronize()  # the staging buffer is reused by the next step: its copy and the send are done
            ser_ms.append(1e3 * (t2 - t1)); xfer_ms.append(1e3 * (time.perf_counter() - t2))
        return pb, st, blocks

    pb = None
    for _ in range(args.warmup):
        pb, _, _ = step()  # held like in the timed loop: the previous block is alive while the next one is produced,
        #                    so both sets of page-locked result buffers exist before the clock starts
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    arcs, launches, waves, got = 0, 0, 0, 0
    inside = {"ms_h2d": 0.0, "ms_expand": 0.0, "ms_connect": 0.0, "ms_d2h": 0.0}
    for _ in range(args.steps):
        pb, st, blocks = step()
        arcs += st["arcs_out"]; launches += st["kernel_launches"]; waves += st["waves"]
        for k in inside:
            inside[k] += st[k]
        if blocks is not None:
            torch.cuda.synchronize()
            got = sum(int(b.numel()) for b in blocks)
    e1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))  # the step is host-driven: take the larger clock
    total_arcs = sum_over_ranks(float(arcs))
    info = pb.info()
    check = None
    if rank == 0 and blocks is not None:  # rank 0 really holds everything: rebuild the blocks and count
        n_res = sum(len(R.PackedBatch.from_buffer(b.cpu().numpy())) for b in blocks)
        check = {"results_on_rank0": n_res, "bytes_on_rank0": got}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib as O
        ot = O.OFst.from_csr(t["offsets"].astype(np.uint64), t["arcs"], t["finals"], t["start"], t["props"])
        n_sample = min(args.batch, 1024)
        secs, oarcs = 0.0, 0
        for d in acc_dicts[:n_sample]:
            oa = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
            r, ost = O.compose(oa, ot, want_stats=True)
            secs += ost["seconds"]; oarcs += r.num_trs
        cpu = {"value": oarcs / secs, "unit": "arcs/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"the first {n_sample} acceptors of the batch composed one by one with the same transducer "
                         f"({oarcs} result arcs in {secs:.2f} s), oracle port, 1 thread"}
    del dt
    if rank != 0:
        return None
    return {"metric": "composed_arcs_per_sec (result arcs of the whole batch / step time)",
            "value": total_arcs / (ms * 1e-3), "unit": "arcs/s", "n_gpus": world, "ms_per_batch": ms / args.steps,
            "scaling": "strong", "steps": args.steps, "warmup": args.warmup,
            "workload": f"C5: {args.batch} linear acceptors (200 arcs) o {n_t}-state/{len(t['arcs'])}-arc transducer, "
                        f"sharded by acceptor x{world}, transducer resident in HBM on every rank, result blocks sent "
                        "to rank 0 over NCCL",
            "per_step": {"h2d_bytes_per_rank": int((hi - lo) * (201 * 8 + 200 * 16 + 4)),
                         "d2h_bytes_per_rank": int(info["bytes"]), "waves": waves // max(1, args.steps),
                         "result_states_rank0_shard": info["num_states"], "result_arcs_rank0_shard": info["num_trs"]},
            "compose_batch_packed_wall_ms_per_call": float(np.mean(call_ms[-args.steps:])),
            "inside_the_call_ms_per_step": {"union_upload": inside["ms_h2d"] / args.steps, "expand": inside["ms_expand"] / args.steps,
                                            "connect": inside["ms_connect"] / args.steps,
                                            "split_and_download": inside["ms_d2h"] / args.steps},
            "gather": check,
            "gather_ms_per_step_rank0": None if not ser_ms else {"serialize_into_pinned": float(np.mean(ser_ms[-args.steps:])),
                                                                 "h2d_sizes_send_recv_sync": float(np.mean(xfer_ms[-args.steps:]))},
            "cpu_baseline": cpu, "gpu_launches": int(launches),
            "timer": "max(CUDA events, host wall clock) around K host-driven steps, max over ranks"}


def run_extras(args, R, synth, torch, peak_gbs, barrier):
    """Two workloads next to the headline ones (N = 1 only), chosen to be UNfavourable where the headline ones are kind:
      * compose_spread — the same operand sizes as C3, but the transducer's arc target depends on (label, source), so
        every state of a level is reached and the arc lists the matcher searches are spread over the whole 160 MB
        machine (HBM-resident) instead of 32 slots per level; both start states fan out (a 20 000-arc hub), labels
        are drawn from 93 symbols so that the product neither dies nor explodes (~1 successor per product state).
      * sssp_window — SURVEY.md 8d's acyclic acceptor with arcs to targets up to 1000 ids ahead: arcs skip levels, the
        longest path has tens of thousands of hops, and a label-correcting relaxation re-relaxes states many times
        (the visit budget trips and the in-order sweep of sssp.cu takes over)."""
    from tests import oracle_lib as O
    out = {}
    steps = max(1, min(args.steps, 5))
    n, a = int(1_000_000 * args.scale), int(10_000_000 * args.scale)
    a1 = synth.layered_acceptor(n, a, 93, 3, 50, start_fanout=True)
    a2 = synth.bigram_transducer(n, a, 93, 4, 50, out_vocab=20000, start_fanout=True, spread=True)
    d1, d2 = R.DeviceFst.upload(synth.to_vector_fst(a1)), R.DeviceFst.upload(synth.to_vector_fst(a2))
    for _ in range(3):
        r, st = R.device_compose(d1, d2)
        del r
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot = {"arcs_emitted": 0, "arcs_iterated": 0, "states_expanded": 0, "ms_emit_kernel": 0.0, "ms_expand": 0.0, "ms_connect": 0.0}
    for _ in range(steps):
        r, st = R.device_compose(d1, d2)
        for k in tot:
            tot[k] += st[k]
        del r
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    bytes_c = 16.0 * tot["arcs_iterated"] + 48.0 * tot["arcs_emitted"] + 32.0 * tot["states_expanded"]
    out["compose_spread"] = {
        "workload": f"layered acceptor ({a1['num_states']} states, {len(a1['arcs'])} arcs) o spread transducer "
                    f"({a2['num_states']} states, {len(a2['arcs'])} arcs), 93 labels, start states fan out",
        "value": tot["arcs_emitted"] / (ms * 1e-3), "unit": "arcs/s", "ms_per_step": ms / steps,
        "states_expanded_per_step": tot["states_expanded"] // steps, "arcs_emitted_per_step": tot["arcs_emitted"] // steps,
        "arcs_iterated_per_step": tot["arcs_iterated"] // steps,
        "expand_ms_per_step": tot["ms_expand"] / steps, "connect_ms_per_step": tot["ms_connect"] / steps,
        "roofline": {"bound": "hbm", "kernel": "k_compose_ws", "bytes_model": "16*A_it + 48*A_out + 32*S",
                     "achieved": bytes_c / (tot["ms_emit_kernel"] * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": bytes_c / (tot["ms_emit_kernel"] * 1e-3) / 1e9 / peak_gbs,
                     "avg_launch_ms": tot["ms_emit_kernel"] / steps}}
    if not args.no_cpu_baseline:
        oa = O.OFst.from_csr(a1["offsets"].astype(np.uint64), a1["arcs"], a1["finals"], a1["start"], a1["props"])
        ob = O.OFst.from_csr(a2["offsets"].astype(np.uint64), a2["arcs"], a2["finals"], a2["start"], a2["props"])
        _, ost = O.compose(oa, ob, want_stats=True)
        out["compose_spread"]["cpu_baseline"] = {"value": ost["arcs_emitted"] / ost["seconds"], "unit": "arcs/s", "cores": 1,
                                                 "kind": "port", "sample": f"one full compose+connect ({ost['seconds']:.2f} s)"}
        del oa, ob
    del d1, d2
    g = synth.window_dag(int(5_000_000 * args.scale), int(50_000_000 * args.scale), 1000, 6, window=1000)
    dg = R.DeviceFst.upload(synth.to_vector_fst(g))
    for _ in range(2):
        R.device_shortest_path(dg)
    barrier()
    t0 = time.perf_counter()
    edges, relaxed, waves = 0, 0, 0
    wsteps = max(1, min(steps, 3))
    for _ in range(wsteps):
        _, sst = R.device_shortest_path(dg)
        relaxed += sst["arcs_relaxed"]; waves = sst["waves"]
    barrier()
    ms = 1e3 * (time.perf_counter() - t0)
    n_edges = len(g["arcs"])
    out["sssp_window"] = {
        "workload": f"window DAG: {g['num_states']} states, {n_edges} arcs, targets up to 1000 ids ahead (TOP_SORTED known)",
        "value": n_edges * wsteps / (ms * 1e-3), "unit": "edges/s (distinct edges of the machine / call time)",
        "ms_per_step": ms / wsteps, "relaxation_waves": waves, "arcs_relaxed_per_call": relaxed // wsteps,
        "re_relaxation_factor": relaxed / wsteps / max(1, n_edges), "device_path": sst["path"],
        "in_order_sweep": int(sst.get("sweep", 0)),
        "note": "label-correcting waves revisit a state whenever a shorter route arrives (456 visits per state here, 387 ms); "
                "they are cut off after 4 visits per state and a top-sorted machine goes to the in-order sweep (one CTA, "
                "distances of the next 8192 ids in a shared-memory ring); relaxation_waves / arcs_relaxed include the "
                "abandoned waves"}
    if not args.no_cpu_baseline:
        og = O.OFst.from_csr(g["offsets"].astype(np.uint64), g["arcs"], g["finals"], g["start"], g["props"])
        _, cst = O.shortest_path(og, want_stats=True)
        out["sssp_window"]["cpu_baseline"] = {"value": cst["arcs_relaxed"] / cst["seconds"], "unit": "edges/s", "cores": 1,
                                              "kind": "port", "sample": f"one full shortest_path ({cst['seconds']:.2f} s)"}
    del dg
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C2", "C5"])
    ap.add_argument("--batch", type=int, default=8192, help="C5: number of linear acceptors")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (tests only; full size = 1.0)")
    ap.add_argument("--callers", type=int, default=2, help="host threads of the supplementary concurrent e2e figure (1 = skip); "
                    "the persistent kernels own the whole GPU (one at a time), so callers overlap their transfers with each "
                    "other's kernels and beyond two or three callers calls only queue")
    ap.add_argument("--no-sssp", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the two extra workloads (spread compose, window-DAG SSSP)")
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
        c5 = run_c5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks, synth)
        if rank == 0:
            print(json.dumps(c5), flush=True)
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
    # Dominant kernel = the persistent BFS kernel k_compose_ws (one launch per compose): algorithmic bytes of the
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
        "bound": "hbm", "kernel": "k_compose_ws (persistent warp-stream kernel: the whole BFS expansion in one launch)",
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
    e2e_parts = {"ms_h2d": 0.0, "ms_expand": 0.0, "ms_connect": 0.0, "ms_d2h": 0.0}
    for _ in range(steps):
        res, st = R.compose_with_stats(h1, h2)  # fst_compose path: H2D of both operands, kernels, D2H of the result
        e2e_arcs += st["arcs_emitted"]
        for k in e2e_parts:
            e2e_parts[k] += st[k]
        d2h = 4 * (res.num_states() + 1) + 16 * res.num_trs_total() + 4 * res.num_states()
        del res
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": sum_over_ranks(float(e2e_arcs)) / (e2e_ms * 1e-3), "unit": "arcs/s",
           "h2d_bytes_per_step": csr_bytes(a1) + csr_bytes(a2), "d2h_bytes_per_step": int(d2h),
           "api": "fst_compose (b200_compose_with_stats) on host VectorFst handles",
           "ms_per_step": e2e_ms / steps,
           "inside_the_call_ms_per_step": {"h2d": e2e_parts["ms_h2d"] / steps, "kernels": (e2e_parts["ms_expand"] + e2e_parts["ms_connect"]) / steps,
                                           "d2h": e2e_parts["ms_d2h"] / steps},
           "note": "the three parts depend on each other inside one call (operands -> kernels -> result), so a single caller is "
                   "bound by link time + kernel time; concurrent callers overlap one call's transfers with another's kernels"}
    # Same calls issued by several host threads at once (the C-ABI is re-entrant, every call owns a stream): the PCIe
    # link is full duplex, so one caller's upload overlaps another's kernels and download.  Supplementary figure; the
    # headline `value` above is the single-caller number.
    if args.callers > 1:
        import threading
        done = [0] * args.callers

        n_calls = max(steps, 8)
        gate = threading.Barrier(args.callers + 1)

        def worker(k):
            for _ in range(2):  # warm-up WITH the same concurrency: every caller needs its own page-locked result buffers
                r, s = R.compose_with_stats(h1, h2)
                del r
            gate.wait()
            time.sleep(0.009 * k)  # callers that start in lockstep upload together, queue for the kernels together and
            #                        download together: nothing overlaps.  Half a call of offset de-phases them.
            for _ in range(n_calls):
                r, s = R.compose_with_stats(h1, h2)
                done[k] += s["arcs_emitted"]
                del r

        th = [threading.Thread(target=worker, args=(k,)) for k in range(args.callers)]
        for t in th:
            t.start()
        gate.wait()  # every caller has finished its warm-up calls
        t0 = time.perf_counter()
        for t in th:
            t.join()
        check_ffi_error(lib.b200_device_synchronize(), "sync")
        dt = time.perf_counter() - t0
        e2e["concurrent_callers"] = {"callers": args.callers, "value": float(sum(done)) / dt, "unit": "arcs/s",
                                     "ms_per_compose": dt * 1e3 / (n_calls * args.callers), "calls_per_caller": n_calls,
                                     "timer": "host wall clock"}
    del d1, d2

    # ------------------------------------------------------------------ SSSP leg (C4)
    # Headline = the lattice with the property word compose leaves behind (ACYCLIC known, TOP_SORTED unknown): AutoQueue
    # picks the TopOrderQueue, whose order is the reference's DFS order — computed on the device (dag_order.cu) inside
    # every timed call.  The same lattice with TOP_SORTED known (StateOrderQueue, no order needed) is reported next to it.
    sssp = None
    if not args.no_sssp:
        from rustfst_b200 import props as PR
        g_sorted = gen_sssp_workload(args.scale, rank, synth)
        g = dict(g_sorted, props=g_sorted["props"] & ~(PR.TOP_SORTED | PR.NOT_TOP_SORTED))
        hg = synth.to_vector_fst(g)
        dg = R.DeviceFst.upload(hg)

        def sssp_run(dev, n_steps):
            for _ in range(args.warmup):
                R.device_shortest_path(dev)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            tw = time.perf_counter()
            tot_ = {"arcs_relaxed": 0, "states_settled": 0, "ms_relax_kernel": 0.0, "ms_device": 0.0, "relax_launches": 0,
                    "kernel_launches": 0, "ms_order_device": 0.0, "ms_queue_plan_host": 0.0}
            last = None
            for _ in range(n_steps):
                _, last = R.device_shortest_path(dev)
                for k in tot_:
                    tot_[k] += last[k]
            s1.record()
            barrier()
            wall = 1e3 * (time.perf_counter() - tw)
            return max_over_ranks(max(s0.elapsed_time(s1), wall)), tot_, last

        sssp_ms, stot, last = sssp_run(dg, steps)
        edges = sum_over_ranks(float(stot["arcs_relaxed"]))
        n_states = g["num_states"]
        sssp_bytes = 16.0 * stot["arcs_relaxed"] + 20.0 * n_states * steps
        relax_gbs = (16.0 * stot["arcs_relaxed"] + 12.0 * stot["states_settled"]) / (stot["ms_relax_kernel"] * 1e-3) / 1e9
        sssp = {
            "metric": "sssp_edges_per_sec", "value": edges / (sssp_ms * 1e-3), "unit": "edges/s",
            "ms_per_step": sssp_ms / steps,
            "workload": f"C4 as a composed lattice: shortest_path(n=1), {n_states}-state/{len(g['arcs'])}-arc layered acyclic "
                        "lattice, property word as compose leaves it (ACYCLIC known, TOP_SORTED unknown -> TopOrderQueue), "
                        "dyadic weights; lattice resident in HBM",
            "queue_kind": last["queue_kind"], "order_on_device": last["order_on_device"],
            "ms_order_device_per_step": stot["ms_order_device"] / steps,
            "ms_queue_plan_host_per_step": stot["ms_queue_plan_host"] / steps,
            "ms_relax_and_backtrace_per_step": stot["ms_device"] / steps,
            "device_path": {0: "parallel relaxation + certificate", 1: "serial replay", 2: "order-faithful parallel fold"}.get(last["path"]),
            "roofline": {"bound": "hbm", "kernel": "k_relax", "achieved": relax_gbs, "peak": peak_gbs, "unit": "GB/s",
                         "frac": relax_gbs / peak_gbs, "traffic": None,
                         "whole_call": {"bytes_model": "16*E + 20*N", "achieved_GBps": sssp_bytes / (sssp_ms * 1e-3) / 1e9,
                                        "frac": sssp_bytes / (sssp_ms * 1e-3) / 1e9 / peak_gbs}},
            "gpu_launches": stot["kernel_launches"],
            "timer": "max(CUDA events, host wall clock) around K blocking calls, max over ranks",
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
        # the same lattice with TOP_SORTED known: StateOrderQueue, no order to compute
        hg_s = synth.to_vector_fst(g_sorted)
        dg_s = R.DeviceFst.upload(hg_s)
        ms_s, st_s, last_s = sssp_run(dg_s, steps)
        sssp["top_sorted_variant"] = {
            "workload": "same lattice, TOP_SORTED known (StateOrderQueue)", "queue_kind": last_s["queue_kind"],
            "value": sum_over_ranks(float(st_s["arcs_relaxed"])) / (ms_s * 1e-3), "unit": "edges/s",
            "ms_per_step": ms_s / steps, "device_path": last_s["path"]}
        del dg_s, hg_s
        # the host DFS the device order replaces (B200_HOST_DFS=1 forces it), one call
        os.environ["B200_HOST_DFS"] = "1"
        try:
            t0 = time.perf_counter()
            _, hst = R.device_shortest_path(dg)
            sssp["host_dfs_variant"] = {"ms_per_call_wall": (time.perf_counter() - t0) * 1e3,
                                        "ms_queue_plan_host_dfs": hst["ms_queue_plan_host"], "ms_device": hst["ms_device"]}
        finally:
            del os.environ["B200_HOST_DFS"]
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
                           "sample": f"one full C4 shortest_path on the same lattice and property word, DFS order included "
                                     f"({sst['arcs_relaxed']} edges in {sst['seconds']:.2f} s)"}
            t0 = time.perf_counter()
            O.shortest_path(og, nshortest=10)
            cpu["sssp"]["nbest_ms"] = (time.perf_counter() - t0) * 1e3
            cpu["sssp"]["nbest_sample"] = "one full C4 shortest_path(nshortest=10), oracle port, 1 thread"

    # ------------------------------------------------------------------ batched compose leg (C5), sharded over the ranks
    c5 = None
    if not args.no_c5:
        c5 = run_c5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks, synth)

    extras = None
    if world == 1 and not args.no_extras and args.workload == "C3":
        extras = run_extras(args, R, synth, torch, peak_gbs, barrier)

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
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "sssp": sssp, "c5": c5, "extra_workloads": extras,
            "gpu_launches": int(tot["kernel_launches"]), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

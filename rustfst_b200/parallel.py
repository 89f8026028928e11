"""Multi-GPU plumbing for the batched-compose mode (BASELINE.json configs[4]): one process per GPU.

The path shards by independent units: acceptor_i o T is an independent composition, so acceptors are block-
partitioned across ranks, the shared transducer is replicated (read-only) on every GPU and NO collective runs
during compute.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to gather the result FSTs
(serialised in the OpenFst binary format) back to rank 0.
"""
from typing import Callable, List, Optional, Sequence

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition: ranks [0, n % world) get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_blobs(blobs: Sequence[bytes], dist, device=None, dst: int = 0) -> Optional[List[bytes]]:
    """Gather variable-length byte strings from every rank to `dst`, preserving (rank, local index) order.

    Two collectives: all_gather of the per-rank byte counts, then all_gather of the padded payloads (on
    NVLink5/NVSwitch every peer is one hop away, so a flat all_gather is the right shape; no topology tuning)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = np.array([len(b) for b in blobs], dtype=np.int64)
    payload = np.frombuffer(b"".join(blobs), dtype=np.uint8) if len(blobs) else np.zeros(0, dtype=np.uint8)
    meta = torch.tensor([len(sizes), payload.size], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_n = int(max(m[0].item() for m in metas))
    max_b = int(max(m[1].item() for m in metas))
    sz = torch.zeros(max(1, max_n), dtype=torch.int64, device=device)
    sz[:len(sizes)] = torch.from_numpy(sizes).to(sz.device)
    pl = torch.zeros(max(1, max_b), dtype=torch.uint8, device=device)
    pl[:payload.size] = torch.from_numpy(payload.copy()).to(pl.device)
    szs = [torch.zeros_like(sz) for _ in range(world)]
    pls = [torch.zeros_like(pl) for _ in range(world)]
    dist.all_gather(szs, sz)
    dist.all_gather(pls, pl)
    if rank != dst:
        return None
    out: List[bytes] = []
    for r in range(world):
        n = int(metas[r][0].item())
        s = szs[r][:n].cpu().numpy()
        raw = pls[r][:int(metas[r][1].item())].cpu().numpy().tobytes()
        off = 0
        for k in s:
            out.append(raw[off:off + int(k)])
            off += int(k)
    return out


def gather_buffers(local, dist, device=None, dst: int = 0):
    """One variable-length uint8 buffer per rank -> list of buffers on `dst` (rank order), None elsewhere.  Only the
    destination receives payload: an all_gather of the byte counts (8 bytes per rank), then point-to-point send/recv
    of exactly those bytes (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    t = torch.as_tensor(local, dtype=torch.uint8)
    if device is not None:
        t = t.to(device, non_blocking=True)
    size = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    if rank != dst:
        if t.numel():
            dist.send(t, dst)
        return None
    out = []
    for r in range(world):
        n = int(sizes[r].item())
        if r == dst:
            out.append(t)
        else:
            buf = torch.empty(n, dtype=torch.uint8, device=t.device)
            if n:
                dist.recv(buf, r)
            out.append(buf)
    return out


def batched_compose_sharded(acceptor_blobs: Sequence[bytes], transducer_blob: bytes,
                            compose_fn: Callable[[List[bytes], bytes], List[bytes]], dist=None, device=None):
    """Shard `acceptor_blobs` over the ranks of `dist`, run `compose_fn(local_acceptors, transducer)` locally and
    gather the serialised results on rank 0 in input order.  `compose_fn` is the device path on GPUs
    (rustfst_b200.compose_batch over deserialised handles); the CPU tests inject a stand-in."""
    if dist is None or not dist.is_initialized():
        return compose_fn(list(acceptor_blobs), transducer_blob)
    lo, hi = shard_range(len(acceptor_blobs), dist.get_rank(), dist.get_world_size())
    local = compose_fn(list(acceptor_blobs[lo:hi]), transducer_blob)
    return gather_blobs(local, dist, device=device)

"""world_size-2 gloo test (CPU) of the N>1 plumbing of the batched mode: block sharding of the acceptors and the
gather of variable-length serialised results back to rank 0 in input order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rustfst_b200.parallel import batched_compose_sharded, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8192):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                got.extend(range(lo, hi))
            assert got == list(range(n))


def _fake_compose(local, transducer):
    # stand-in for the device path: deterministic, variable-length, depends on both inputs
    return [a[::-1] + transducer[: (len(a) % 5)] for a in local]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blobs = [bytes([i % 251]) * (1 + (i * 7) % 23) for i in range(n_items)]
    out = batched_compose_sharded(blobs, b"TRANSDUCER", _fake_compose, dist)
    if rank == 0:
        q.put(out == _fake_compose(blobs, b"TRANSDUCER"))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [0, 5, 64])
def test_gather_world_size_2_gloo(n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok

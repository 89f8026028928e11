"""PackedBatch (the block a rank sends to rank 0 in the sharded batched-compose mode) on the CPU: the byte layout
round-trips, every index of a block that comes from another process is validated, and the point-to-point gather
delivers each rank's block to rank 0 only (world_size-2 gloo)."""
import os
import socket
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rustfst_b200 as R
from rustfst_b200.fst import TR_DTYPE
from rustfst_b200.parallel import gather_buffers

MAGIC = 0x4B43415030303242  # "B200PACK"


def _block(results):
    """results: list of (offsets, arcs, finals, start or -1, props) with local state ids -> bytes in the block layout."""
    n = len(results)
    state_off, arc_off, starts, props = [0], [0], [], []
    offs, fins, arcs = [0], [], []
    for o, a, f, st, p in results:
        a0 = arc_off[-1]
        offs.pop()
        offs.extend(int(a0 + x) for x in o)
        fins.extend(f)
        arcs.append(np.asarray(a, dtype=TR_DTYPE))
        state_off.append(state_off[-1] + len(f)); arc_off.append(a0 + len(a))
        starts.append(st); props.append(p)
    arcs = np.concatenate(arcs) if arcs else np.zeros(0, dtype=TR_DTYPE)
    return (struct.pack("<4Q", MAGIC, n, len(fins), len(arcs)) + np.array(state_off, "<u4").tobytes() +
            np.array(arc_off, "<u4").tobytes() + np.array(starts, "<i4").tobytes() + np.array(props, "<u8").tobytes() +
            np.array(offs, "<u4").tobytes() + np.array(fins, "<f4").tobytes() + arcs.tobytes())


def _sample():
    a0 = np.array([(1, 2, 0.5, 1), (3, 4, 1.5, 1)], dtype=TR_DTYPE)
    a2 = np.array([(7, 7, 0.25, 2), (8, 8, 0.0, 2), (9, 9, 2.0, 0)], dtype=TR_DTYPE)
    return [([0, 2, 2], a0, [np.inf, 0.75], 0, 0x10000),
            ([0], np.zeros(0, dtype=TR_DTYPE), [], -1, 0),
            ([0, 1, 2, 3], a2, [np.inf, np.inf, 1.0], 1, 0x30000)]


def test_block_round_trip_and_results():
    blob = _block(_sample())
    pb = R.PackedBatch.from_buffer(blob)
    assert pb.info() == {"n": 3, "num_states": 5, "num_trs": 5, "bytes": len(blob)}
    assert pb.to_bytes() == blob
    for i, (o, a, f, st, p) in enumerate(_sample()):
        r = pb.result(i)
        ro, ra, rf, rs = r.to_csr()
        assert list(ro) == list(o) and np.array_equal(ra, a) and np.array_equal(rf, np.array(f, dtype=np.float32))
        assert rs == (st if st >= 0 else None) and r.properties == p
    with pytest.raises(ValueError):
        pb.result(3)
    # serialising into a caller-owned (e.g. page-locked, reused) buffer: the used prefix comes back
    big = np.full(len(blob) + 100, 0xAB, dtype=np.uint8)
    view = pb.to_numpy(out=big)
    assert view.nbytes == len(blob) and view.tobytes() == blob and np.shares_memory(view, big)
    assert (big[len(blob):] == 0xAB).all()
    with pytest.raises(ValueError, match="at least"):
        pb.to_numpy(out=np.zeros(len(blob) - 1, dtype=np.uint8))
    with pytest.raises(ValueError):
        pb.to_numpy(out=np.zeros(len(blob), dtype=np.uint16))


def test_large_block_round_trips_through_the_threaded_copy():
    """Arrays of more than 4 MB are serialised by several host threads (batch.cu put())."""
    n_states, n_arcs = 300_000, 600_000
    arcs = np.zeros(n_arcs, dtype=TR_DTYPE)
    rng = np.random.default_rng(5)
    arcs["ilabel"] = rng.integers(1, 1000, n_arcs); arcs["olabel"] = arcs["ilabel"]
    arcs["weight"] = rng.integers(0, 64, n_arcs) / 8.0
    arcs["nextstate"] = rng.integers(0, n_states, n_arcs)
    offs = np.arange(0, n_arcs + 1, 2, dtype=np.uint32)
    blob = _block([(list(offs), arcs, list(np.zeros(n_states, dtype=np.float32)), 0, 0)])
    pb = R.PackedBatch.from_buffer(blob)
    assert pb.info()["num_trs"] == n_arcs and pb.to_bytes() == blob


@pytest.mark.parametrize("damage", ["truncate", "magic", "nextstate", "start", "offsets", "totals"])
def test_damaged_blocks_are_rejected(damage):
    res = _sample()
    if damage == "nextstate":
        res[2][1]["nextstate"][0] = 3           # result 2 has 3 states
    if damage == "start":
        res[0] = (res[0][0], res[0][1], res[0][2], 2, res[0][4])
    if damage == "offsets":
        res[2] = ([0, 2, 1, 3], res[2][1], res[2][2], res[2][3], res[2][4])
    blob = bytearray(_block(res))
    if damage == "truncate":
        blob = blob[:-7]
    if damage == "magic":
        blob[0] ^= 0xFF
    if damage == "totals":
        blob[24] += 1                            # number of arcs in the header
    with pytest.raises(ValueError):
        R.PackedBatch.from_buffer(bytes(blob))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _block(_sample()[: 2 + rank])        # different sizes per rank
    out = gather_buffers(torch.frombuffer(bytearray(mine), dtype=torch.uint8), dist)
    if rank == 0:
        ok = len(out) == world
        for r, b in enumerate(out):
            pb = R.PackedBatch.from_buffer(b.numpy().tobytes())
            ok = ok and len(pb) == 2 + r and pb.to_bytes() == _block(_sample()[: 2 + r])
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_blocks_are_gathered_on_rank_0_only_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok

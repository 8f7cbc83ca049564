"""Multi-rank host logic on CPU: two gloo ranks shard one batch with no data-path collective and
gather the end-of-run statistics record (the only communication of the N>1 path)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_outputs(rec):
    """Deterministic stand-in for per-QP solver outputs (the GPU is not needed for the plumbing)."""
    B = rec.B
    status = np.where(rec.t0 * 1000 % 7 == 0, -2, 1).astype(np.int32)
    iters = (rec.t0 * 1000 % 5).astype(np.int32)
    stats = np.zeros((B, 8))
    stats[:, 0] = np.abs(rec.x0[:, 0]) * 1e-9
    stats[:, 1] = np.abs(rec.x0[:, 1]) * 1e-9
    stats[:, 3] = 108 + 3 * (rec.t0 * 1000 % 5)
    stats[:, 6] = rec.t0 * 1000 % 3
    stats[:, 7] = rec.t0 * 1000 % 4
    return status, iters, stats


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    from convex_mpc_b200 import records, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert sharding.env_rank_world() == (rank, rank, world)
    rec = records.random_records(B, seed=7)
    lo, hi = sharding.shard_range(B, rank, world)
    mine = rec.shard(rank, world)
    assert mine.B == hi - lo and np.array_equal(mine.x0, rec.x0[lo:hi])
    st, it, stats = _fake_outputs(mine)
    local = sharding.local_stats(st, it, stats, elapsed_ms=10.0 + rank, flops=float(mine.B))
    allr = sharding.gather_stats(local)
    q.put((rank, lo, hi, allr))
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather():
    sys.path.insert(0, ROOT)
    from convex_mpc_b200 import records, sharding
    B, world = 1001, 2                       # ragged: 501 + 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1:3] == (0, 501) and got[1][1:3] == (501, 1001)
    assert np.array_equal(got[0][3], got[1][3]) and got[0][3].shape == (2, len(sharding.STAT_FIELDS))
    summ = sharding.reduce_stats(got[0][3])
    rec = records.random_records(B, seed=7)
    st, it, stats = _fake_outputs(rec)
    whole = sharding.reduce_stats(sharding.local_stats(st, it, stats, 11.0, float(B))[None, :])
    for k in sharding.STAT_FIELDS:
        assert abs(summ[k] - whole[k]) <= 1e-12 * max(1.0, abs(whole[k])), k


def test_shard_range_edges():
    sys.path.insert(0, ROOT)
    from convex_mpc_b200 import sharding
    for B in (0, 1, 7, 8, 65536, 262144):
        for world in (1, 2, 4, 8):
            cover = [sharding.shard_range(B, r, world) for r in range(world)]
            assert cover[0][0] == 0 and cover[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
            assert max(hi - lo for lo, hi in cover) <= -(-B // world)

"""The N>1 host path on CPU: two gloo ranks shard a shot range, all-reduce the counters, and farm jobs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tensorqec.jl_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(1001, rank, world)
    # stand-in for the per-rank pipeline: counts derived from the global shot indices of this rank's range
    idx = np.arange(lo, hi)
    local = np.array([(idx % 7 == 0).sum(), (idx % 11 == 0).sum(), ((idx % 7 == 0) | (idx % 11 == 0)).sum(), hi - lo])
    tot = sharding.allreduce_counts(local)
    sq = sharding.multiprocess_run(lambda x: x ** 2, range(1, 6))
    q.put((rank, (lo, hi), tot.tolist(), sq))
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(1001)
    want = [int((idx % 7 == 0).sum()), int((idx % 11 == 0).sum()), int(((idx % 7 == 0) | (idx % 11 == 0)).sum()), 1001]
    assert out[0][1] == (0, 501) and out[1][1] == (501, 1001)
    for _, _, tot, sq in out:
        assert tot == want
        assert sq == [1, 4, 9, 16, 25]                       # test/multiprocessing.jl:3-6


def test_shard_range_covers_everything():
    from tensorqec.jl_b200.sharding import multiprocess_run, shard_range
    for n in (0, 1, 7, 10 ** 7):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    assert multiprocess_run(lambda x: x ** 2, range(1, 6)) == [1, 4, 9, 16, 25]

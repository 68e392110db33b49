"""world_size-2 `gloo` test (CPU) of the N>1 host logic: shard ranges tile the instance stream exactly and the final
statistics reduction sums / maxes across ranks.  No data-path collective exists."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from wbc_quadruped_dob_b200 import scenarios as S
from wbc_quadruped_dob_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    sc = S.make(hi - lo, mode_mix=(0.25, 0.375, 0.375), pushes=True, seed=2, start=lo)
    st = sharding.local_stats(hi - lo, status=np.zeros(hi - lo, dtype=np.int32),
                              qp_info=np.full((8, hi - lo), rank + 1), qp_flops=np.full(hi - lo, 2.0), ms=10.0 * (rank + 1))
    tot = sharding.gather_stats(st)
    q.put((rank, lo, hi, float(sc["q"].sum()), int(sc["mode"].sum()), tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_gather():
    n_total, world = 9001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, qs0, ms0, t0), (r1, lo1, hi1, qs1, ms1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4500, 4500, 9001)
    whole = S.make(n_total, mode_mix=(0.25, 0.375, 0.375), pushes=True, seed=2)
    assert np.isclose(qs0 + qs1, whole["q"].sum(), rtol=1e-12) and ms0 + ms1 == int(whole["mode"].sum())
    assert t0 == t1
    assert t0["instances"] == n_total and t0["sum_flops"] == 2.0 * n_total
    assert t0["sum_ncholesky"] == 4500 * 1 + 4501 * 2 and t0["max_ms"] == 20.0 and t0["max_kkt_dim"] == 2.0


def test_shard_ranges_tile():
    for n in (0, 1, 7, 4096, 1048576):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))

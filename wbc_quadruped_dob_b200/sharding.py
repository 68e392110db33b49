"""Batch sharding across ranks: instances are independent -- including the observer
recurrence, whose state is per instance -- so rank r owns the contiguous range [r*N/G, (r+1)*N/G) and there is no
data-path collective.  The only exchange is one small statistics reduction at the end of a run."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous range of instance indices owned by `rank`."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


STAT_KEYS = ("instances", "solver_failures", "sum_ncholesky", "sum_outer_its", "sum_flops", "max_ms", "max_kkt_dim")
_MAX_KEYS = ("max_ms", "max_kkt_dim")


def local_stats(n, status=None, qp_info=None, qp_flops=None, ms=0.0):
    s = dict.fromkeys(STAT_KEYS, 0.0)
    s["instances"] = float(n)
    s["max_ms"] = float(ms)
    if status is not None:
        s["solver_failures"] = float(np.count_nonzero(np.asarray(status) != 0))
    if qp_info is not None:
        qi = np.asarray(qp_info)
        s["sum_ncholesky"] = float(qi[0].sum())
        s["sum_outer_its"] = float(qi[1].sum())
        s["max_kkt_dim"] = float(qi[4].max()) if qi.shape[1] else 0.0
    if qp_flops is not None:
        s["sum_flops"] = float(np.asarray(qp_flops).sum())
    return s


def gather_stats(stats, device=None):
    """All-reduce the statistics vector over the default process group (NCCL on GPUs, gloo on CPU): sums for the
    counters, max for the max_* keys.  With no initialised group this is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    sums = torch.tensor([stats[k] for k in STAT_KEYS if k not in _MAX_KEYS], dtype=torch.float64, device=device)
    maxs = torch.tensor([stats[k] for k in _MAX_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out = {}
    it_s, it_m = iter(sums.tolist()), iter(maxs.tolist())
    for k in STAT_KEYS:
        out[k] = next(it_m) if k in _MAX_KEYS else next(it_s)
    return out

"""Upper bound of phase-coherent execution: the same batch size with (a) the workload's varied instances and (b) every
instance a copy of one of them (all warps walk the same code at the same time).  Compares ns per solve at equal work."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
from wbc_quadruped_dob_b200 import api, scenarios as S
n = 16384
cfg = dict(S.CONFIGS["standing_4096"]); cfg.pop("n")
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
def run(scn, label):
    b.set_observer_state(scn["obs_yd"], scn["obs_yw"])
    ts = []; cyc = None
    for it in range(6):
        out = b.cycle(scn, want=("qp_flops",))
        f, s_ = b.last_timing(); ts.append(s_)
        b.set_observer_state(scn["obs_yd"], scn["obs_yw"])
    cyc = b.last_solve_cycles(n).astype(np.float64)
    print("%-28s solve %.3f ms  -> %.0f solves/s ; mean per-instance latency %.3f ms ; flops/solve %.3g" % (label, np.median(ts[2:]), n / np.median(ts[2:]) * 1e3, cyc.mean() / 1.965e6, out["qp_flops"].mean()))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "varied"): run(sc, "varied instances")
for pick in ((0, 1, 2) if which == "all" else ((1,) if which == "same" else ())):
    same = {k: (np.ascontiguousarray(np.repeat(v[..., pick:pick + 1], n, axis=-1)) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    run(same, "all copies of instance %d" % pick)

"""Where does a small batch's step go?  Per-instance solve latencies (clock64, the figure the dispatch order is built from) in the
steady state (longest-first order in use) against the solver kernel's duration, for the resident-warp count in WBC_SOLVE_CTAS_PER_SM."""
import os, sys, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
name = sys.argv[1] if len(sys.argv) > 1 else "standing_4096"
cfg = dict(S.CONFIGS[name]); n = cfg.pop("n"); n = min(n, int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
ks = []
for it in range(6):
    out = b.cycle(sc)
    ks.append(b.last_timing()[1])
occ, smem, grid = b.solver_shape()
ms = b.last_solve_cycles(n).astype(np.float64) / 1.965e6
print("%s n %d: grid %d warps (%s), solver kernel ms per cycle %s" % (name, n, grid, b.last_solver_kernel, " ".join("%.3f" % k for k in ks)))
print("  steady state: sum of latencies / warps = %.3f ms, longest %.3f ms, mean %.3f, p99 %.3f; kernel %.3f ms -> %.1f %% above the larger bound"
      % (ms.sum() / grid, ms.max(), ms.mean(), np.percentile(ms, 99), ks[-1], 100.0 * (ks[-1] / max(ms.sum() / grid, ms.max()) - 1.0)))

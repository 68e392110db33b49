"""Express lanes: which solves are the longest, and did they run on an express warp (rank by flop count below the head)?"""
import os, sys, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
cfg = dict(S.CONFIGS["standing_4096"]); n = cfg.pop("n")
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
for it in range(6):
    out = b.cycle(sc)
ms = b.last_solve_cycles(n).astype(np.float64) / 1.965e6
fl = out["qp_flops"]
rank = np.empty(n, dtype=int); rank[np.argsort(-fl, kind="stable")] = np.arange(n)
print("kernel %.3f ms; mean latency %.3f" % (b.last_timing()[1], ms.mean()))
for lo, hi in ((0, 64), (64, 128), (128, 256), (256, 512), (512, 1024), (1024, 4096)):
    m = (rank >= lo) & (rank < hi)
    print("  flop rank %4d..%4d: latency mean %.3f max %.3f ms, flops mean %.3g, ms per Mflop %.3f" % (lo, hi, ms[m].mean(), ms[m].max(), fl[m].mean(), (ms[m] / fl[m]).mean() * 1e6))
top = np.argsort(-ms)[:12]
for i in top: print("  long: %.3f ms flops %.3g rank %d nchol %d outer %d qqp %d nicwork %d kktdim %d flags %d reused %d  (ms per Mflop %.3f)" % ((ms[i], fl[i], rank[i]) + tuple(int(out["qp_info"][k][i]) for k in (0, 1, 2, 3, 4, 5, 6)) + (ms[i] / fl[i] * 1e6,)))

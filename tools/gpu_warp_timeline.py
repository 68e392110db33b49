"""Per-warp timeline of the one-warp-per-solve kernel (WBC_STAGE_PROF=1): when do the warps start, how long are they busy, when do
they run dry -- where a small batch's step goes.  Usage: gpu_warp_timeline.py [workload] [n]"""
import os, sys, numpy as np
os.environ["WBC_STAGE_PROF"] = "1"
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
name = sys.argv[1] if len(sys.argv) > 1 else "standing_4096"
cfg = dict(S.CONFIGS[name]); n = cfg.pop("n"); n = min(n, int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
for it in range(6): out = b.cycle(sc)
k_ms = b.last_timing()[1]
p = b.stage_profile().astype(np.float64)
p = p[p[:, 1] > 0]                                  # warps that solved something (express SMs send most of theirs home)
t0 = p[:, 2].min()
entry, first, last, exit_ = [(p[:, k] - t0) * 1e-6 for k in (2, 3, 4, 5)]
busy = p[:, 0] / 1.965e6
print("%s n %d: kernel %.3f ms, %d working warps (%d express), solves per warp mean %.2f max %d" % (name, n, k_ms, len(p), int(p[:, 7].sum()), p[:, 1].mean(), int(p[:, 1].max())))
print("  kernel entry (ms after the first warp): p50 %.3f p99 %.3f max %.3f ; first solve starts: p50 %.3f max %.3f" % (np.median(entry), np.percentile(entry, 99), entry.max(), np.median(first), first.max()))
print("  last solve ends: min %.3f p10 %.3f p50 %.3f p90 %.3f max %.3f ; busy per warp: mean %.3f min %.3f max %.3f" % (last.min(), np.percentile(last, 10), np.median(last), np.percentile(last, 90), last.max(), busy.mean(), busy.min(), busy.max()))
for lab, m in (("express", p[:, 7] == 1), ("regular", p[:, 7] == 0)):
    if m.any(): print("  %s warps: %d, busy mean %.3f ms, last end p50 %.3f max %.3f, longest solve max %.3f ms, solves per warp %.2f" % (lab, m.sum(), busy[m].mean(), np.median(last[m]), last[m].max(), p[m, 6].max() / 1.965e6, p[m, 1].mean()))
late = np.argsort(-last)[:6]
for i in late: print("  late warp: ends %.3f ms, %d solves, busy %.3f, longest %.3f, express %d, SM %d" % (last[i], int(p[i, 1]), busy[i], p[i, 6] / 1.965e6, int(p[i, 7]), int(p[i, 8])))

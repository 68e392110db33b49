"""Distribution of per-instance solver work on the GPU (flops proxy, flags, counts) for a workload."""
import sys, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
name = sys.argv[1] if len(sys.argv) > 1 else "standing_4096"
cfg = dict(S.CONFIGS[name]); n = cfg.pop("n"); n = min(n, 65536)
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
out = b.cycle(sc)
qi, fl = out["qp_info"], out["qp_flops"]
print(name, "n", n, "flops mean %.3g max %.3g p99 %.3g  max/mean %.2f" % (fl.mean(), fl.max(), np.percentile(fl, 99), fl.max() / fl.mean()))
names = ["nchol", "outer", "qqp_calls", "nicwork", "kktdim", "flags"]
for k, nm in enumerate(names):
    v = qi[k]
    print("  %-9s mean %.2f max %d  hist(top) %s" % (nm, v.mean(), v.max(), dict(zip(*np.unique(v, return_counts=True))) if nm in ("flags", "outer", "qqp_calls") else ""))
cyc = b.last_solve_cycles(n).astype(np.float64)
ms = cyc / 1.965e6
print("  solve latency ms: mean %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f   corr(flops, cycles) %.3f" % (ms.mean(), np.median(ms), np.percentile(ms, 90), np.percentile(ms, 99), ms.max(), np.corrcoef(fl, cyc)[0, 1]))
f0, s0 = b.last_timing()
print("  kernel ms front %.3f solve %.3f ; sum of latencies / 1184 warps = %.3f ms" % (f0, s0, ms.sum() / 1184))
top = np.argsort(-cyc)[:8]
for i in top: print("  heavy inst %d ms %.3f flops %.3g nchol %d outer %d qqp %d nicwork %d flags %d" % (i, ms[i], fl[i], qi[0][i], qi[1][i], qi[2][i], qi[3][i], qi[5][i]))
fb = (qi[5] & 8) != 0
if fb.any():
    print("  literal-KKT fallback instances: %d (%.2f%%), latency ms mean %.3f max %.3f (others mean %.3f)" % (fb.sum(), 100.0 * fb.mean(), ms[fb].mean(), ms[fb].max(), ms[~fb].mean()))
sp = (qi[5] & 32) != 0
if sp.any():
    print("  spill-mode instances: %d, latency ms mean %.3f max %.3f" % (sp.sum(), ms[sp].mean(), ms[sp].max()))
print("  Newton factorisations reused: %.2f of %.2f per solve; multiplier-cache hit in %.1f%% of solves" % (qi[6].mean(), qi[0].mean(), 100.0 * np.mean((qi[5] & 64) != 0)))

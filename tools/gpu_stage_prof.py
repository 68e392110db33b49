"""Where do the warps of the staged solver kernel spend their cycles?  (WBC_STAGE_PROF=1: per-warp counters of the last launch)"""
import os, sys, numpy as np
os.environ["WBC_STAGE_PROF"] = "1"
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
name = sys.argv[1] if len(sys.argv) > 1 else "standing_4096"
cfg = dict(S.CONFIGS[name]); n = cfg.pop("n"); n = min(n, 65536)
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
for it in range(3):
    out = b.cycle(sc)
f_ms, s_ms = b.last_timing()
p = b.stage_profile().astype(np.float64)
tot = p[:, 9]
print(name, "n", n, "solve kernel %.3f ms, warps %d, mean kernel cycles per warp %.3g (%.3f ms at 1.965 GHz)" % (s_ms, len(p), tot.mean(), tot.mean() / 1.965e6))
for role, nm in ((0, "QQP-role"), (1, "POST-role"), (2, "SETUP-role")):
    m = p[:, 10] == role
    if not m.any(): continue
    q = p[m]; t = q[:, 9].sum()
    print("  %-18s warps %4d | share of their cycles: QQP %.1f%%  POST %.1f%%  SETUP %.1f%%  looking for a task %.1f%%  fences+queue %.1f%% | tasks per warp: Q %.1f U %.1f S %.1f, stolen %.1f"
          % (nm, m.sum(), 100 * q[:, 0].sum() / t, 100 * q[:, 1].sum() / t, 100 * q[:, 2].sum() / t, 100 * q[:, 6].sum() / t, 100 * q[:, 7].sum() / t,
             q[:, 3].mean(), q[:, 4].mean(), q[:, 5].mean(), q[:, 8].mean()))
cnt = p[:, 3:6].sum(axis=0); cyc = p[:, 0:3].sum(axis=0)
print("  mean task duration (us): QQP %.1f  POST %.1f  SETUP %.1f ; tasks per solve: Q %.2f U %.2f S %.2f" % tuple(list(cyc / np.maximum(cnt, 1) / 1965.0) + list(cnt / n)))
print("  sum of task cycles per solve: %.3f ms (QQP %.3f, POST %.3f, SETUP %.3f)" % (cyc.sum() / n / 1.965e6, cyc[0] / n / 1.965e6, cyc[1] / n / 1.965e6, cyc[2] / n / 1.965e6))

"""Small mixed batch (stance + both swing modes, terrain, pushes) + trajectory sampling + plant step, for compute-sanitizer."""
import sys, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
sc = S.make(n, mode_mix=(0.3, 0.35, 0.35), pushes=True, terrain=True, seed=77)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
for it in range(2):
    out = b.cycle(sc)
tr = S.make_trajectory(sc, nseg=3, seed=3)
b.set_trajectory(tr)
b.sample_trajectory(n, t=tr["t"])
out2 = b.cycle({k: v for k, v in sc.items() if k not in api.TRAJ_FIELDS}, sampled_traj=True)
print("status ok:", bool((out["status"] == 0).all()), "flags", np.unique(out["qp_info"][5]), "finite", bool(np.isfinite(out2["tau"]).all()))
b.close()

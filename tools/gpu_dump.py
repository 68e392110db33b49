"""Bit-exactness harness for refactors that must not change a result: `dump` writes the outputs of three fixed batches to
gpurun_out/ (merged back by gpurun); `compare PATH` re-runs them and compares bit for bit with a dump kept in-tree."""
import sys, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
mode, path = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/exact_ref.npz")
res = {}
for name, n in (("standing_4096", 4096), ("trot_65536", 6000), ("mixed_terrain_1m", 6000)):
    cfg = dict(S.CONFIGS[name]); cfg.pop("n")
    sc = S.make(n, start=0, **cfg)
    b = api.WbcBatch(max_batch=n, device=0)
    b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
    out = b.cycle(sc)
    out2 = b.cycle(sc)          # second cycle: observer state advanced, longest-first order in use
    for k in ("tau", "w", "x", "qp_obj"):
        res[name + "/" + k] = out[k]; res[name + "/2/" + k] = out2[k]
    res[name + "/nchol"] = out["qp_info"][0]
    b.close()
if mode == "dump":
    np.savez_compressed(path, **res); print("dumped", path)
else:
    ref = np.load(path); bad = [k for k in res if not np.array_equal(res[k], ref[k])]
    print("BIT-IDENTICAL" if not bad else "DIFFERENT: %s" % bad[:6])
    for k in bad:
        a, r = np.asarray(res[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
        d = (a != r) & ~(np.isnan(a) & np.isnan(r))
        inst = d.any(axis=0) if d.ndim > 1 else d
        print("  %-28s instances differing %d of %d, max abs diff %.3e" % (k, int(inst.sum()), inst.size, float(np.nanmax(np.abs(a - r)))))

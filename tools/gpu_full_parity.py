"""Every instance of BASELINE configs 3 and 4 (one GPU's shard) against the live oracle (reference ALGLIB): how many torques deviate by
more than 1e-7 / 1e-6 relative, worst deviation, histogram of Cholesky-count differences.  (The test suite checks 16 384 of each.)"""
import sys, time, numpy as np
sys.path.insert(0, ".")
from oracle import oracle_py as O
from tests import util
from wbc_quadruped_dob_b200 import api, scenarios as S
for cfg, n in (("trot_65536", 65536), ("mixed_terrain_1m", 131072)):
    sc = S.make_config(cfg, n=n)
    b = api.WbcBatch(max_batch=n, device=0)
    b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
    got = b.cycle(sc)
    t0 = time.time()
    ref, _ = O.run_cycle_batch(sc, nthreads=32)
    ok = (ref["status"] == 0) & (got["status"] == 0)
    et = util.rel_rows(got["tau"].T, ref["tau"])[ok]
    ew = np.abs(got["w"].T - ref["w"]).max()
    eo = (np.abs(got["qp_obj"] - ref["qp_obj"]) / np.maximum(1e-30, np.abs(ref["qp_obj"])))[ok]
    d = got["qp_info"][0].astype(int) - ref["ncholesky"].astype(int)
    vals, cnts = np.unique(d, return_counts=True)
    print("%s, all %d instances (oracle %.1f s): status mismatch %d; torque rel err worst %.2e, > 1e-7: %d, > 1e-6: %d; objective rel err worst %.2e; observer abs err %.2e; ncholesky equal on %.3f %%, differences %s"
          % (cfg, n, time.time() - t0, int((ref["status"] != got["status"]).sum()), et.max(), int((et > 1e-7).sum()), int((et > 1e-6).sum()), eo.max(), ew,
             100.0 * np.mean(d == 0), dict(zip(vals.tolist(), cnts.tolist()))))
    b.close()

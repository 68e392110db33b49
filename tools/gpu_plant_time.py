import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
for n in (4096, 65536):
    sc = S.make(n, mode_mix=(0.4, 0.3, 0.3), pushes=False, seed=5)
    dev = torch.device("cuda", 0)
    din = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
    tau = torch.zeros(12, n, dtype=torch.float64, device=dev); push = torch.zeros(6, n, dtype=torch.float64, device=dev)
    b = api.WbcBatch(max_batch=n)
    for sub in (1, 5):
        for it in range(3): b.plant_dynamics_step(din, tau, push, n=n, ld=n, substeps=sub, gamma=100.0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(20): b.plant_dynamics_step(din, tau, push, n=n, ld=n, substeps=sub, gamma=100.0)
        torch.cuda.synchronize(); print("n %d substeps %d: %.3f ms per plant step" % (n, sub, (time.perf_counter() - t0) / 20 * 1e3))
    b.close()

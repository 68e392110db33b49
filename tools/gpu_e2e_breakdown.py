"""Where the end-to-end (host buffers) step time goes: python mirror, C-ABI call, kernels."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api, scenarios as S
import ctypes as C
name = sys.argv[1] if len(sys.argv) > 1 else "standing_4096"
cfg = dict(S.CONFIGS[name]); n = cfg.pop("n")
sc = S.make(n, start=0, **cfg)
b = api.WbcBatch(max_batch=n, device=0)
b.set_observer_state(sc["obs_yd"], sc["obs_yw"])
for label, host in (("pageable", sc), ("pinned, per-field", {k: (b.pinned_copy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}),
                    ("pinned slab", b.pinned_inputs(sc))):
    out = {"tau": b.pinned((12, n)), "w": b.pinned((6, n))} if label != "pageable" else None
    for _ in range(3): b.cycle(host, want=(), out=out)
    t = []; g = []
    for _ in range(20):
        t0 = time.perf_counter(); b.cycle(host, want=(), out=out); t.append(time.perf_counter() - t0)
        f, s_ = b.last_timing(); g.append(f + s_)
    print("%-18s n=%d  call %.3f ms (min %.3f)  kernels %.3f ms  overhead %.3f ms" % (label, n, 1e3 * np.median(t), 1e3 * min(t), np.median(g), 1e3 * np.median(t) - np.median(g)))
# python-only cost of building the argument structs
keep = []
t0 = time.perf_counter()
for _ in range(100): b._inputs_struct(sc, n, keep)
print("python _inputs_struct: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / 100))

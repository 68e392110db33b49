#!/usr/bin/env python3
"""Generate the committed golden fixtures (run in the BUILD container, where /root/reference exists).

  cycle_*.npz   inputs (SoA) + outputs of the CPU oracle: oracle/wbc_oracle.c (our restatement of
                main.cpp's cycle) with the QP solved by the REFERENCE's own vendored ALGLIB compiled from
                /root/reference (oracle/_ref/libref_alglib_qp.so).
  qp_*.npz      dense (Q, c, L) problems assembled by the oracle and the reference ALGLIB's x for each,
                with the reference settings of lopt.cpp:91-106 -- outputs of the reference itself.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as op                      # noqa: E402
from wbc_quadruped_dob_b200 import scenarios as S       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cycle_fixture(name, sc, chain_observer=False):
    n = sc["mode"].shape[0]
    if chain_observer:
        # config 1: ONE robot, instance index = cycle index; observer state chained cycle to cycle
        outs = []
        yd, yw = np.zeros(6), np.zeros(6)
        sc = {k: v.copy() for k, v in sc.items()}
        for i in range(n):
            sc["obs_yd"][:, i], sc["obs_yw"][:, i] = yd, yw
            one = {k: (v[..., i:i + 1] if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
            res, _ = op.run_cycle_batch(one)
            outs.append(res)
            yd, yw = res["yd"][0], res["yw"][0]
        res = {k: np.concatenate([o[k] for o in outs], axis=0) for k in outs[0]}
    else:
        res, _ = op.run_cycle_batch(sc, nthreads=8)
    assert (res["status"] == 0).all()
    out = {"in_" + k: v for k, v in sc.items() if v is not None}
    out.update({"out_" + k: v for k, v in res.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, n, "instances; mean ncholesky", res["ncholesky"].mean())


def qp_fixture(name, sc, idx):
    Qs, cs, Ls, xs, nch = [], [], [], [], []
    neq = None
    for i in idx:
        _, qp = op.assemble_only(sc, i, w=[0.3, -0.2, 0.5, 0.01, -0.02, 0.03])
        Q = np.array(qp.Q).reshape(30, 30)
        c = np.array(qp.c)
        L = np.array(qp.L)[:qp.nrows * 31].reshape(qp.nrows, 31)
        x, nc, rc = op.ref_qp_solve(Q, c, L, qp.neq)
        assert rc == 0
        neq = qp.neq
        Qs.append(Q); cs.append(c); Ls.append(L); xs.append(x); nch.append(nc)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), Q=np.array(Qs), c=np.array(cs), L=np.array(Ls), x=np.array(xs),
                        ncholesky=np.array(nch), neq=neq)
    print(name, len(idx), "QPs")


def main():
    op.build(ref=True)
    cycle_fixture("cycle_standing", S.make(96, mode_mix=(1.0, 0.0, 0.0), seed=1))
    cycle_fixture("cycle_trot_pushes", S.make(96, mode_mix=(0.25, 0.375, 0.375), pushes=True, seed=2))
    cycle_fixture("cycle_mixed_terrain", S.make(64, mode_mix=(0.4, 0.3, 0.3), pushes=True, terrain=True, seed=3))
    cycle_fixture("cycle_trot_replay", S.trot_replay(cycles_per_phase=12), chain_observer=True)
    sc = S.make(64, mode_mix=(0.5, 0.25, 0.25), pushes=True, seed=5)
    qp_fixture("qp_stance", sc, [i for i in range(64) if sc["mode"][i] == 0][:16])
    qp_fixture("qp_swing", sc, [i for i in range(64) if sc["mode"][i] != 0][:16])


if __name__ == "__main__":
    main()

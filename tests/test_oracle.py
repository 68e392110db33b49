"""The CPU oracle itself: physics identities for the (unpinned) rigid-body restatement, the compiled reference
ALGLIB against a tightly converged solve, and the committed golden fixtures."""
import numpy as np
import pytest

from tests import util
from wbc_quadruped_dob_b200 import scenarios as S

MASS = 21.261


@pytest.fixture(scope="module")
def sc():
    return S.make(24, mode_mix=(0.4, 0.3, 0.3), pushes=True, seed=21)


def test_mass_matrix_identities(oracle, sc):
    for i in range(6):
        d = oracle.update_only(sc, i)
        M = np.array(d.M).reshape(18, 18)
        assert np.abs(M - M.T).max() < 1e-12
        assert np.linalg.eigvalsh(M).min() > 0
        assert np.allclose(M[:3, :3], MASS * np.eye(3), atol=1e-12)          # total mass 21.261 (dogbot_model.h:91)
        # centre-of-mass Jacobian = M[0:3,:]/m: check against com_vel
        nu = np.concatenate([sc["base_vel"][:, i], sc["dq"][:, i]])
        assert np.allclose(M[:3] @ nu / MASS, np.array(d.com_vel), atol=1e-12)
        # the whole point of T: MassMatrixCOM is block diagonal with m*I on top
        Mc = np.array(d.Mcom).reshape(18, 18)
        assert np.allclose(Mc[:3, :3], MASS * np.eye(3), atol=1e-10)
        assert np.abs(Mc[:6, 6:]).max() < 1e-10 and np.abs(Mc[:3, 3:6]).max() < 1e-10


def test_com_and_feet_match_independent_kinematics(oracle, sc):
    com, feet = S.forward_kinematics(sc["base_pos"], sc["base_rot"], sc["q"])
    for i in range(6):
        d = oracle.update_only(sc, i)
        assert np.allclose(np.array(d.com), com[i], atol=1e-12)
        assert np.allclose(np.array(d.foot_pos).reshape(4, 3), feet[i], atol=1e-12)


def test_bias_equals_gravity_at_rest(oracle, sc):
    rest = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    rest["base_vel"][:] = 0.0
    rest["dq"][:] = 0.0
    for i in range(4):
        d = oracle.update_only(rest, i)
        assert np.allclose(np.array(d.h), np.array(d.g), atol=1e-12)
        assert np.allclose(np.array(d.g)[:3], [0, 0, MASS * 9.8], atol=1e-10)   # gravity (0,0,-9.8), main.cpp:855
        assert np.abs(np.array(d.Jdqd_lin)).max() < 1e-14


def _integrate(sc, i, dt):
    """Advance instance i by dt along its velocity (base: v, omega in world axes; joints: dq)."""
    one = {k: (v[..., i:i + 1].copy() if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    w = one["base_vel"][3:6, 0]
    th = np.linalg.norm(w) * dt
    R = one["base_rot"][:, 0].reshape(3, 3)
    if th != 0:
        a = w / np.linalg.norm(w)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K) @ R
    one["base_rot"][:, 0] = R.reshape(9)
    one["base_pos"][:, 0] += dt * one["base_vel"][0:3, 0]
    one["q"][:, 0] += dt * one["dq"][:, 0]
    return one


def test_jacobian_and_bias_acceleration_by_finite_differences(oracle, sc):
    """J nu = foot velocity, and Jdqd = d/dt(J) nu (central difference along the motion, nu held constant)."""
    eps = 1e-6
    for i in range(4):
        d = oracle.update_only(sc, i)
        J = np.array(d.Jac_lin).reshape(12, 18)
        nu = np.concatenate([sc["base_vel"][:, i], sc["dq"][:, i]])
        assert np.allclose(J @ nu, np.array(d.foot_vel), atol=1e-12)
        dp = oracle.update_only(_integrate(sc, i, eps), 0)
        dm = oracle.update_only(_integrate(sc, i, -eps), 0)
        vdot = (np.array(dp.foot_vel) - np.array(dm.foot_vel)) / (2 * eps)       # nu_dot = 0 along this path
        assert np.allclose(vdot, np.array(d.Jdqd_lin), atol=2e-7)
        pdot = (np.array(dp.foot_pos) - np.array(dm.foot_pos)) / (2 * eps)
        assert np.allclose(pdot, np.array(d.foot_vel), atol=1e-8)


def test_bias_force_by_energy_balance(oracle, sc):
    """With nu_dot = 0: d/dt(M nu) - dT/dq = C nu, and h = C nu + g.  Checked on the power identity
    nu' (h - g) = d/dt(1/2 nu' M nu) along the constant-nu motion (C - 1/2 dM/dt is skew in MIXED coordinates
    only up to the base-frame transport term, so use the scalar identity nu' C nu = 1/2 nu' Mdot nu)."""
    eps = 1e-6
    for i in range(4):
        d = oracle.update_only(sc, i)
        nu = np.concatenate([sc["base_vel"][:, i], sc["dq"][:, i]])
        Mp = np.array(oracle.update_only(_integrate(sc, i, eps), 0).M).reshape(18, 18)
        Mm = np.array(oracle.update_only(_integrate(sc, i, -eps), 0).M).reshape(18, 18)
        Tdot = 0.5 * nu @ ((Mp - Mm) / (2 * eps)) @ nu
        power = nu @ (np.array(d.h) - np.array(d.g))
        assert abs(power - Tdot) < 1e-6 * max(1.0, abs(Tdot))


def test_reference_alglib_converges_on_stance(oracle, have_ref, sc):
    """stance-shaped problems agree with a tightly converged solve to <= 2e-8."""
    for i in [k for k in range(24) if sc["mode"][k] == 0][:6]:
        _, qp = oracle.assemble_only(sc, i)
        Q = np.array(qp.Q).reshape(30, 30)
        c = np.array(qp.c)
        L = np.array(qp.L)[:qp.nrows * 31].reshape(qp.nrows, 31)
        x, nch, rc = oracle.ref_qp_solve(Q, c, L, qp.neq)
        xe, _, rce = oracle.ref_qp_solve(Q, c, L, qp.neq, exact=True)
        assert rc == 0 and rce == 0 and nch > 0
        assert np.abs(x - xe).max() <= 1e-6 * max(1.0, np.abs(xe).max())


@pytest.mark.parametrize("name", ["cycle_standing", "cycle_trot_pushes", "cycle_mixed_terrain"])
def test_oracle_reproduces_golden(oracle, have_ref, name):
    sc, gold = util.load_golden(name)
    res, _ = oracle.run_cycle_batch(sc, nthreads=4)
    for k in ("tau", "w", "x", "yd", "yw", "qp_obj"):
        assert np.array_equal(res[k], gold[k]), k
    assert np.array_equal(res["ncholesky"], gold["ncholesky"])


@pytest.mark.parametrize("name", ["qp_stance", "qp_swing"])
def test_reference_alglib_reproduces_golden_qp(oracle, have_ref, name):
    z = np.load(util.GOLDEN + "/" + name + ".npz")
    for k in range(z["Q"].shape[0]):
        x, nch, rc = oracle.ref_qp_solve(z["Q"][k], z["c"][k], z["L"][k], int(z["neq"]))
        assert rc == 0 and nch == z["ncholesky"][k]
        assert np.array_equal(x, z["x"][k])

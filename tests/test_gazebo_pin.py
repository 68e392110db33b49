"""The rigid-body stage against the one artefact the reference ships that records its robot in motion: the Gazebo log decoded
into tests/golden/gazebo_states.npz (generator: tests/golden/make_gazebo_fixture.py, tools/decode_gazebo_log.py).

Nothing here goes through tools/gen_model.py, which generates the model tables of BOTH the kernels and the oracle: the kinematic
tree used to turn logged link poses into (base pose, joint angles, rates) is the SDF that Gazebo itself built from dogbot.urdf
and embedded in the log.  The oracle's (CPU test) and the kernels' (GPU test) forward kinematics must then land on the logged
link frames: lower-leg orientation, foot position and velocity, whole-body centre of mass and its velocity -- to the log's print
precision (1e-5 in poses, 1e-4 in twists).  This pins joint axes and their signs, hip / pitch / knee offsets, the foot offset
(urdf:320-325), link masses and centres of mass.  (It does not make the dynamics "reference-pinned": iDynTree is still absent.)"""
import os

import numpy as np
import pytest

from tests import util
from wbc_quadruped_dob_b200 import scenarios as S

FIX = os.path.join(util.GOLDEN, "gazebo_states.npz")
STACKED_LEG = [1, 0, 2, 3]          # stacked foot order BR, BL, FL, FR (main.cpp:674-686) -> leg index in BL, BR, FL, FR


def _rot(rpy):
    return S.rpy_to_rot(np.asarray(rpy, dtype=np.float64).reshape(3, -1))          # [n,3,3], R = Rz Ry Rx as Gazebo prints poses


def _signed_angle(Rrel, axis):
    """Angle of the rotation Rrel about the unit vector `axis` (Rrel = Rot(axis, q))."""
    v = 0.5 * np.stack([Rrel[:, 2, 1] - Rrel[:, 1, 2], Rrel[:, 0, 2] - Rrel[:, 2, 0], Rrel[:, 1, 0] - Rrel[:, 0, 1]], axis=1)
    s = v @ axis
    c = 0.5 * (np.trace(Rrel, axis1=1, axis2=2) - 1.0)
    # the rotation should be about that axis only; the off-axis part is print rounding -- or ODE's hinge constraint giving way
    # under the first impacts, in which case the state is not one a rigid tree can reproduce to 1e-5
    return np.arctan2(s, c), np.abs(v - np.outer(s, axis)).max(axis=1)


def gazebo_scenario():
    z = np.load(FIX)
    names = [str(x) for x in z["link_names"]]
    li = {nm: k for k, nm in enumerate(names)}
    pose, vel = z["pose"], z["vel"]                     # [n, link, 6]
    n = pose.shape[0]
    R = {nm: _rot(pose[:, li[nm], 3:6].T) for nm in names}
    p = {nm: pose[:, li[nm], 0:3] for nm in names}
    v = {nm: vel[:, li[nm], 0:3] for nm in names}
    w = {nm: vel[:, li[nm], 3:6] for nm in names}
    assert not z["link_pose0"][:, 3:].any()             # every link frame is aligned with the model frame at the zero configuration
    jn = [str(x) for x in z["joint_names"]]
    jpar, jch, jax = [str(x) for x in z["joint_parent"]], [str(x) for x in z["joint_child"]], z["joint_axis"]
    assert (z["joint_axis_in_parent_model_frame"] == 1).all()
    legs = [str(x) for x in z["legs"]]
    q, dq = np.zeros((12, n)), np.zeros((12, n))
    hinge_err = np.zeros(n)
    for leg, nm in enumerate(legs):
        for jname, dof in ((nm + "_roll_joint", leg), (nm + "_pitch_joint", 4 + 2 * leg), (nm + "_knee_joint", 5 + 2 * leg)):
            k = jn.index(jname)
            a = jax[k] / np.linalg.norm(jax[k])
            Rrel = np.einsum("nji,njk->nik", R[jpar[k]], R[jch[k]])
            q[dof], err = _signed_angle(Rrel, a)
            hinge_err = np.maximum(hinge_err, err)
            aw = np.einsum("nij,j->ni", R[jpar[k]], a)
            dq[dof] = np.einsum("ni,ni->n", w[jch[k]] - w[jpar[k]], aw)
    sc = {"base_pos": np.ascontiguousarray(p["base_link"].T), "base_rot": np.ascontiguousarray(R["base_link"].reshape(n, 9).T),
          "base_rpy": np.ascontiguousarray(pose[:, li["base_link"], 3:6].T),
          "base_vel": np.ascontiguousarray(np.hstack([v["base_link"], w["base_link"]]).T), "q": q, "dq": dq,
          "mode": np.zeros(n, dtype=np.int32), "foot_force": np.zeros((12, n)), "terrain": None,
          "obs_yd": np.zeros((6, n)), "obs_yw": np.zeros((6, n))}
    for k in ("com_des_pos", "com_des_vel", "com_des_acc", "sw_des_pos", "sw_des_vel", "sw_des_acc"):
        sc[k] = np.zeros((6, n))
    # what the log says about the quantities the controller computes
    M = z["link_mass"].sum()
    com = sum(z["link_mass"][li[nm]] * (p[nm] + np.einsum("nij,j->ni", R[nm], z["link_com"][li[nm], 0:3])) for nm in names) / M
    comv = sum(z["link_mass"][li[nm]] * (v[nm] + np.cross(w[nm], np.einsum("nij,j->ni", R[nm], z["link_com"][li[nm], 0:3]))) for nm in names) / M
    foot_p, foot_v, foot_R = np.zeros((n, 4, 3)), np.zeros((n, 4, 3)), np.zeros((n, 4, 3, 3))
    for sf, leg in enumerate(STACKED_LEG):
        nm = legs[leg] + "_lowerleg"
        off = np.einsum("nij,j->ni", R[nm], z["foot_offset"][leg])
        foot_p[:, sf] = p[nm] + off
        foot_v[:, sf] = v[nm] + np.cross(w[nm], off)
        foot_R[:, sf] = R[nm]
    keep = hinge_err < 3e-5                              # states in which every hinge is a hinge to print precision
    assert hinge_err.max() < 1e-3
    sc = {k: (np.ascontiguousarray(v_[..., keep]) if isinstance(v_, np.ndarray) else v_) for k, v_ in sc.items()}
    return sc, dict(com=com[keep], com_vel=comv[keep], foot_pos=foot_p[keep], foot_vel=foot_v[keep], foot_R=foot_R[keep], mass=M, q=q[:, keep])


def test_fixture_is_a_moving_robot():
    sc, ref = gazebo_scenario()
    n = sc["mode"].shape[0]
    assert n >= 500 and abs(ref["mass"] - S.TOTAL_MASS) < 1e-9
    assert np.ptp(sc["q"], axis=1).min() > 0.05                 # every joint moves
    assert np.ptp(sc["base_pos"][2]) > 0.1                      # the robot is dropped and stands up
    # the hand-typed numpy kinematics of scenarios.py against the log as well
    com, feet = S.forward_kinematics(sc["base_pos"], sc["base_rot"], sc["q"])
    assert np.abs(com - ref["com"]).max() < 1e-4 and np.abs(feet - ref["foot_pos"]).max() < 1e-4


def test_oracle_kinematics_reproduce_the_logged_link_frames(oracle):
    sc, ref = gazebo_scenario()
    n = sc["mode"].shape[0]
    worst = dict(com=0.0, com_vel=0.0, foot_pos=0.0, foot_vel=0.0, foot_R=0.0)
    for i in range(n):
        d = oracle.update_only(sc, i)
        worst["com"] = max(worst["com"], np.abs(np.array(d.com) - ref["com"][i]).max())
        worst["com_vel"] = max(worst["com_vel"], np.abs(np.array(d.com_vel) - ref["com_vel"][i]).max())
        worst["foot_pos"] = max(worst["foot_pos"], np.abs(np.array(d.foot_pos).reshape(4, 3) - ref["foot_pos"][i]).max())
        worst["foot_vel"] = max(worst["foot_vel"], np.abs(np.array(d.foot_vel).reshape(4, 3) - ref["foot_vel"][i]).max())
        worst["foot_R"] = max(worst["foot_R"], np.abs(np.array(d.foot_R).reshape(4, 3, 3) - ref["foot_R"][i]).max())
    print("oracle vs Gazebo log over %d states: %s" % (n, {k: "%.2e" % v for k, v in worst.items()}))
    assert worst["foot_pos"] < 1e-4 and worst["com"] < 1e-4 and worst["foot_R"] < 1e-4
    assert worst["foot_vel"] < 5e-3 and worst["com_vel"] < 1e-3      # twists are printed to 1e-4 and the joint rates are differences of them


@pytest.mark.gpu
def test_kernel_kinematics_reproduce_the_logged_link_frames(gpu_batch):
    sc, ref = gazebo_scenario()
    n = sc["mode"].shape[0]
    dbg = gpu_batch.debug_update(sc)
    fp = dbg["foot_pos"].T.reshape(n, 4, 3)
    fv = dbg["foot_vel"].T.reshape(n, 4, 3)
    e = dict(com=np.abs(dbg["com"].T - ref["com"]).max(), com_vel=np.abs(dbg["com_vel"].T - ref["com_vel"]).max(),
             foot_pos=np.abs(fp - ref["foot_pos"]).max(), foot_vel=np.abs(fv - ref["foot_vel"]).max())
    print("kernel vs Gazebo log over %d states: %s" % (n, {k: "%.2e" % v for k, v in e.items()}))
    assert e["foot_pos"] < 1e-4 and e["com"] < 1e-4 and e["foot_vel"] < 5e-3 and e["com_vel"] < 1e-3

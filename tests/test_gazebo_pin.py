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
    # kinetic energy, linear momentum and angular momentum about the base origin of the LOGGED motion: every link's logged twist
    # with the SDF's mass, centre of mass and inertia tensor -- no model of ours involved
    T = np.zeros(n); Pl = np.zeros((n, 3)); La = np.zeros((n, 3)); Pg = np.zeros(n)
    for nm in names:
        k = li[nm]
        m = z["link_mass"][k]
        rc = np.einsum("nij,j->ni", R[nm], z["link_com"][k, 0:3])
        vc = v[nm] + np.cross(w[nm], rc)
        ixx, iyy, izz, ixy, ixz, iyz = z["link_inertia"][k]
        Il = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        Iw = np.einsum("nij,jk,nlk->nil", R[nm], Il, R[nm])
        Iww = np.einsum("nij,nj->ni", Iw, w[nm])
        T += 0.5 * m * np.einsum("ni,ni->n", vc, vc) + 0.5 * np.einsum("ni,ni->n", w[nm], Iww)
        Pl += m * vc
        La += Iww + np.cross(p[nm] + rc - p["base_link"], m * vc)
        Pg += m * (-9.8) * vc[:, 2]                          # power of gravity on this link (gravity (0, 0, -9.8), main.cpp:855)
    keep = hinge_err < 3e-5                              # states in which every hinge is a hinge to print precision
    assert hinge_err.max() < 1e-3
    sc = {k: (np.ascontiguousarray(v_[..., keep]) if isinstance(v_, np.ndarray) else v_) for k, v_ in sc.items()}
    return sc, dict(com=com[keep], com_vel=comv[keep], foot_pos=foot_p[keep], foot_vel=foot_v[keep], foot_R=foot_R[keep], mass=M, q=q[:, keep],
                    kinetic=T[keep], lin_mom=Pl[keep], ang_mom=La[keep], gravity_power=Pg[keep])


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


def _model_tables():
    """Masses, centres of mass and inertia tensors of the kernels' lumped bodies, parsed from csrc/dogbot_model.h (what the CUDA
    code compiles in), in leg order BL, BR, FL, FR x (hip, upperleg, lowerleg + foot)."""
    import re
    txt = open(os.path.join(util.ROOT, "wbc_quadruped_dob_b200", "csrc", "dogbot_model.h")).read()

    def table(name, shape):
        m = re.search(name + r"[^=]*=\s*\{(.*?)\};", txt, re.S)
        body = re.sub(r"//[^\n]*", "", m.group(1))
        vals = [float(v) for v in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]
        return np.array(vals).reshape(shape)
    return (table("kLinkMass", (4, 3)), table("kLinkCom", (4, 3, 3)), table("kLinkInertia", (4, 3, 6)),
            float(re.search(r"kBaseMass\s*=\s*([0-9.eE+-]+)", txt).group(1)), table("kBaseInertia", (3,)))


def test_inertial_parameters_match_the_sdf_gazebo_built_from_the_urdf():
    """Masses, centres of mass and inertia tensors compiled into the kernels (dogbot_model.h, generated by tools/gen_model.py from
    dogbot.urdf) against the inertial blocks of the SDF that Gazebo built from the same URDF and embedded in its log -- a second,
    independent reading of the file.  Base, hips and upper legs agree to the SDF's print precision.  The lower legs carry the 1 g
    foot link (fixed joint, urdf:320-325): Gazebo's converter of that vintage lumped it AT THE LOWER LEG'S ORIGIN, the model
    tables lump it at its joint offset (as iDynTree does for fixed joints): the lumped centre of mass differs by exactly that,
    1.06 mm in z, and is reproduced here from both conventions."""
    z = np.load(FIX)
    names = [str(x) for x in z["link_names"]]
    mass, com, inertia, base_mass, base_inertia = _model_tables()
    legs = [str(x) for x in z["legs"]]
    k = names.index("base_link")
    assert abs(z["link_mass"][k] - base_mass) < 1e-9
    assert np.abs(z["link_inertia"][k, :3] - base_inertia).max() < 1e-6 and not z["link_inertia"][k, 3:].any()
    assert not z["link_com"][:, 3:].any()               # every inertial frame is aligned with its link frame
    for leg, nm in enumerate(legs):
        for j, part in enumerate(("hip", "upperleg", "lowerleg")):
            k = names.index(nm + "_" + part)
            assert abs(z["link_mass"][k] - mass[leg, j]) < 1e-9, (nm, part)
            if part != "lowerleg":
                assert np.abs(z["link_com"][k, :3] - com[leg, j]).max() < 1e-6, (nm, part)
                assert np.abs(z["link_inertia"][k] - inertia[leg, j]).max() < 1e-8, (nm, part)
            else:
                # un-lump: lower leg alone = URDF (0.302 kg at (0, -0.029, -0.1439)); the foot is 1 g
                m_foot = 0.001
                m_leg = mass[leg, j] - m_foot
                c_leg = (mass[leg, j] * com[leg, j] - m_foot * z["foot_offset"][leg]) / m_leg          # the tables' convention
                c_sdf = z["link_mass"][k] * z["link_com"][k, :3] / m_leg                                # foot at the origin
                assert np.abs(c_leg - c_sdf).max() < 2e-6, (nm, c_leg, c_sdf)
                assert np.abs(z["link_com"][k, :3] - com[leg, j]).max() < 1.2e-3                        # the two lumpings, 1 mm apart
                assert np.abs(z["link_inertia"][k] - inertia[leg, j]).max() < 2e-5


def _energy_and_momentum_errors(M_of, sc, ref):
    """nu' M nu / 2 and the base rows of M nu (linear momentum; angular momentum about the base origin -- MIXED representation)
    against the kinetic energy and momenta of the logged link twists."""
    n = sc["mode"].shape[0]
    eT, eP, eL, Tmax = 0.0, 0.0, 0.0, 0.0
    for i in range(n):
        M = M_of(i)
        nu = np.concatenate([sc["base_vel"][:, i], sc["dq"][:, i]])
        Mnu = M @ nu
        T = 0.5 * nu @ Mnu
        eT = max(eT, abs(T - ref["kinetic"][i]) / max(ref["kinetic"][i], 1e-3))
        eP = max(eP, np.abs(Mnu[0:3] - ref["lin_mom"][i]).max())
        eL = max(eL, np.abs(Mnu[3:6] - ref["ang_mom"][i]).max())
        Tmax = max(Tmax, T)
    return eT, eP, eL, Tmax


def test_oracle_mass_matrix_reproduces_the_kinetic_energy_and_momentum_of_the_logged_motion(oracle):
    """The dynamics are not reference-pinned (no iDynTree), but the mass matrix has to agree with the one artefact that records the
    robot moving: nu' M nu / 2 is the kinetic energy, and the base rows of M nu the linear momentum and the angular momentum about
    the base origin, of the thirteen logged link twists weighted with the SDF's inertial data.  To the log's print precision."""
    sc, ref = gazebo_scenario()
    eT, eP, eL, Tmax = _energy_and_momentum_errors(lambda i: np.array(oracle.update_only(sc, i).M).reshape(18, 18), sc, ref)
    print("oracle M vs the logged motion: kinetic energy rel err %.2e (largest %.2f J), linear momentum %.2e kg m/s, angular %.2e kg m2/s" % (eT, Tmax, eP, eL))
    assert Tmax > 1.0                                    # the drop carries real energy
    assert eT < 1e-2 and eP < 5e-3 and eL < 2e-3


@pytest.mark.gpu
def test_kernel_mass_matrix_reproduces_the_kinetic_energy_and_momentum_of_the_logged_motion(gpu_batch):
    sc, ref = gazebo_scenario()
    dbg = gpu_batch.debug_update(sc)
    eT, eP, eL, Tmax = _energy_and_momentum_errors(lambda i: dbg["M"][:, i].reshape(18, 18), sc, ref)
    print("kernel M vs the logged motion: kinetic energy rel err %.2e (largest %.2f J), linear momentum %.2e kg m/s, angular %.2e kg m2/s" % (eT, Tmax, eP, eL))
    assert eT < 1e-2 and eP < 5e-3 and eL < 2e-3


def _gravity_power_error(g_of, sc, ref):
    """-g(q)' nu is the power gravity puts into the robot; from the log it is the sum of m grav . v_c over the thirteen links."""
    n = sc["mode"].shape[0]
    worst, scale = 0.0, 0.0
    for i in range(n):
        nu = np.concatenate([sc["base_vel"][:, i], sc["dq"][:, i]])
        worst = max(worst, abs(-(g_of(i) @ nu) - ref["gravity_power"][i]))
        scale = max(scale, abs(ref["gravity_power"][i]))
    return worst, scale


def test_oracle_gravity_forces_reproduce_the_power_of_gravity_in_the_logged_motion(oracle):
    sc, ref = gazebo_scenario()
    worst, scale = _gravity_power_error(lambda i: np.array(oracle.update_only(sc, i).g), sc, ref)
    print("oracle g vs the logged motion: power of gravity abs err %.2e W (largest %.1f W)" % (worst, scale))
    assert scale > 50.0 and worst < 2e-3 * scale


@pytest.mark.gpu
def test_kernel_gravity_forces_reproduce_the_power_of_gravity_in_the_logged_motion(gpu_batch):
    sc, ref = gazebo_scenario()
    dbg = gpu_batch.debug_update(sc)
    worst, scale = _gravity_power_error(lambda i: dbg["g"][:, i], sc, ref)
    print("kernel g vs the logged motion: power of gravity abs err %.2e W (largest %.1f W)" % (worst, scale))
    assert worst < 2e-3 * scale

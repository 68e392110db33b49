"""The C++ host layer (include/wbc_dogctrl.hpp): the mirrors of the reference's OPT (lopt.h:5-36) and
DOGCTRL::update (main.cpp:63) compile with plain g++, link against libwbc_b200.so, refuse to run without a GPU
(no CPU fallback), and on a B200 reproduce the golden trot replay and the reference ALGLIB's QP solutions."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import util
from wbc_quadruped_dob_b200 import build as B

SRC = os.path.join(util.ROOT, "tests", "cpp", "dogctrl_host.cpp")
EXE = os.path.join(util.ROOT, "tests", "cpp", "dogctrl_host.bin")


@pytest.fixture(scope="module")
def exe():
    lib = B.build()
    deps = [SRC, os.path.join(util.ROOT, "include", "wbc_dogctrl.hpp"), os.path.join(util.ROOT, "include", "wbc_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        libdir = os.path.dirname(lib)
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(util.ROOT, "include"), SRC, "-o", EXE,
                        "-L" + libdir, "-lwbc_b200", "-Wl,-rpath," + libdir], check=True)
    return EXE


def test_host_layer_compiles_and_refuses_without_gpu(exe):
    import torch
    out = subprocess.run([exe, "probe"], capture_output=True, text=True, check=True).stdout.strip()
    assert out == ("ok" if torch.cuda.is_available() else "nodev")


@pytest.mark.gpu
def test_dogctrl_trot_replay_matches_golden(exe, tmp_path):
    """BASELINE config 1 through DogCtrl::update / cycle_stance / cycle_swing, observer chained inside the object."""
    sc, gold = util.load_golden("cycle_trot_replay")
    n = sc["mode"].shape[0]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("i", n))
        for i in range(n):
            H = np.eye(4)
            H[:3, :3] = sc["base_rot"][:, i].reshape(3, 3)
            H[:3, 3] = sc["base_pos"][:, i]
            rec = [H.ravel(), sc["q"][:, i], sc["dq"][:, i], sc["base_vel"][:, i], np.array([0.0, 0.0, -9.8]), sc["base_rpy"][:, i],
                   sc["com_des_pos"][:, i], sc["com_des_vel"][:, i], sc["com_des_acc"][:, i], sc["foot_force"][:, i],
                   sc["sw_des_pos"][:, i], sc["sw_des_vel"][:, i], sc["sw_des_acc"][:, i], np.array([float(sc["mode"][i])])]
            f.write(np.concatenate(rec).astype(np.float64).tobytes())
    subprocess.run([exe, "cycle", fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64).reshape(n, 50)
    assert (out[:, 49] == 0).all()
    assert util.rel_rows(out[:, :12], gold["tau"]).max() <= util.TOL_TAU
    assert np.abs(out[:, 12:18] - gold["w"]).max() <= util.TOL_OBS * max(1.0, np.abs(gold["w"]).max())
    eo = np.abs(out[:, 48] - gold["qp_obj"]) / np.maximum(1e-30, np.abs(gold["qp_obj"]))
    assert eo.max() <= util.TOL_OBJ


def _write_replay(fin, sc, idx):
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(idx)))
        for i in idx:
            H = np.eye(4)
            H[:3, :3] = sc["base_rot"][:, i].reshape(3, 3)
            H[:3, 3] = sc["base_pos"][:, i]
            rec = [H.ravel(), sc["q"][:, i], sc["dq"][:, i], sc["base_vel"][:, i], np.array([0.0, 0.0, -9.8]), sc["base_rpy"][:, i],
                   sc["com_des_pos"][:, i], sc["com_des_vel"][:, i], sc["com_des_acc"][:, i], sc["foot_force"][:, i],
                   sc["sw_des_pos"][:, i], sc["sw_des_vel"][:, i], sc["sw_des_acc"][:, i], np.array([float(sc["mode"][i])])]
            f.write(np.concatenate(rec).astype(np.float64).tobytes())


@pytest.mark.gpu
def test_dogctrl_plan_on_the_device_matches_host_fed_samples(exe, tmp_path):
    """DogCtrl::set_trajectory / sample_trajectory / cycle_*(true): a cycle fed from the samples that stay
    on the GPU gives the torques of the same cycle fed those samples from the host, bit for bit, in all three modes."""
    sc, _ = util.load_golden("cycle_trot_replay")
    idx = [0, 5, 12, 20, 30, 36, 47]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_replay(fin, sc, idx)
    subprocess.run([exe, "traj", fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64).reshape(len(idx), 24)
    assert np.isfinite(out).all() and np.abs(out[:, :12]).max() > 0.1
    assert np.array_equal(out[:, :12], out[:, 12:])
    assert {int(sc["mode"][i]) for i in idx} == {0, 1, 2}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["qp_stance", "qp_swing"])
def test_opt_mirror_matches_reference_alglib(exe, tmp_path, name):
    z = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    for i in range(0, z["Q"].shape[0], 5):
        fin, fout = str(tmp_path / "q.bin"), str(tmp_path / "x.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("i", z["L"].shape[1]))
            f.write(np.ascontiguousarray(z["Q"][i]).tobytes() + np.ascontiguousarray(z["c"][i]).tobytes() + np.ascontiguousarray(z["L"][i]).tobytes())
        subprocess.run([exe, "opt", fin, fout], check=True)
        x = np.fromfile(fout, dtype=np.float64)
        assert np.abs(x - z["x"][i]).max() <= 1e-7 * max(1.0, np.abs(z["x"][i]).max())

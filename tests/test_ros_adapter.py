"""the ROS-facing adapter's message handling (include/wbc_ros_adapter.hpp) against literal restatements of the
reference's callbacks (dogbot_controller/src/client/main.cpp:388-456, 768-779, 794-834).  No ROS and no GPU needed: the node glue
itself (under WBC_WITH_ROS) cannot be compiled in this image and is not tested."""
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from tests import util

SRC = os.path.join(util.ROOT, "tests", "cpp", "ros_adapter_host.cpp")
EXE = os.path.join(util.ROOT, "tests", "cpp", "ros_adapter_host.bin")

# kinDynComp.getDescriptionOfDegreeOfFreedom(i) for the DoF order of main.cpp:612-613 (dogbot.urdf:180-933)
DOF_NAMES = ["back_left_roll_joint", "back_right_roll_joint", "front_left_roll_joint", "front_right_roll_joint",
             "back_left_pitch_joint", "back_left_knee_joint", "back_right_pitch_joint", "back_right_knee_joint",
             "front_left_pitch_joint", "front_left_knee_joint", "front_right_pitch_joint", "front_right_knee_joint"]


@pytest.fixture(scope="module")
def exe():
    deps = [SRC, os.path.join(util.ROOT, "include", "wbc_ros_adapter.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(util.ROOT, "include"), SRC, "-o", EXE], check=True)
    return EXE


def _reference_callbacks(msg_names):
    """jointStateCallback (main.cpp:388-414) and publish_cmd (main.cpp:768-779), literally, on the same synthetic message."""
    id2index, index2id = {}, {}
    for i in range(12):
        index = 0
        while index < len(msg_names):
            if msg_names[index] == DOF_NAMES[i]:
                id2index[i] = index
                index2id[index] = i
                break
            index += 1
    position = [100.0 + k for k in range(len(msg_names))]
    velocity = [200.0 + k for k in range(len(msg_names))]
    q = [position[id2index[i]] for i in range(12)]
    dq = [velocity[id2index[i]] for i in range(12)]
    tau = [300.0 + i for i in range(12)]
    data = []
    for i in range(11, -1, -1):
        data.append(tau[index2id[i]])
    return q, dq, data


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_joint_map_and_command_order_match_the_reference_callbacks(exe, tmp_path, seed):
    rng = np.random.default_rng(seed)
    names = list(DOF_NAMES) if seed == 0 else [DOF_NAMES[k] for k in rng.permutation(12)]
    if seed == 0:
        names = sorted(names)                       # gazebo_ros_control publishes joint_states in alphabetical order
    p = tmp_path / "names.txt"
    p.write_text("\n".join(names) + "\n")
    out = subprocess.run([exe, "joints", str(p)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {l.split()[0]: [float(v) for v in l.split()[1:]] for l in out if l}
    q, dq, data = _reference_callbacks(names)
    assert got["q"] == q and got["dq"] == dq and got["cmd"] == data


def test_joint_map_refuses_a_message_without_every_dof(exe, tmp_path):
    p = tmp_path / "names.txt"
    p.write_text("\n".join(DOF_NAMES[:11]) + "\n")
    out = subprocess.run([exe, "joints", str(p)], capture_output=True, text=True, check=True).stdout.strip()
    assert out == "incomplete"


def test_model_state_gives_tf_rotation_and_fixed_axis_rpy(exe):
    rng = np.random.default_rng(7)
    for k in range(40):
        quat = rng.standard_normal(4) * (0.2 + 3.0 * rng.random())          # unnormalised on purpose (main.cpp:433)
        if k == 0:
            quat = np.array([0.0, np.sin(np.pi / 4), 0.0, np.cos(np.pi / 4)])  # pitch = +90 deg: tf's gimbal branch
        pos, lin, ang = rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(3)
        args = [repr(float(v)) for v in np.concatenate([pos, quat, lin, ang])]
        vals = np.array([float(v) for v in subprocess.run([exe, "pose"] + args, capture_output=True, text=True, check=True).stdout.split()])
        H, base_pos, base_vel = vals[:16].reshape(4, 4), vals[16:22], vals[22:28]
        rot = Rotation.from_quat(quat / np.linalg.norm(quat))
        assert np.abs(H[:3, :3] - rot.as_matrix()).max() < 1e-14
        assert np.array_equal(H[:3, 3], pos) and np.array_equal(H[3], [0, 0, 0, 1])
        assert np.array_equal(base_pos[:3], pos) and np.array_equal(base_vel, np.concatenate([lin, ang]))
        roll, pitch, yaw = base_pos[3:]
        rebuilt = Rotation.from_euler("xyz", [roll, pitch, yaw]).as_matrix()   # extrinsic xyz = Rz(yaw) Ry(pitch) Rx(roll)
        assert np.abs(rebuilt - rot.as_matrix()).max() < (1e-7 if k == 0 else 1e-12)
        assert abs(pitch) <= np.pi / 2 + 1e-12


def test_contact_sample_keeps_the_last_force_when_the_message_is_empty(exe):
    out = subprocess.run([exe, "contact"], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert out == ["0 0 0 0", "1 1 2 30", "0 1 2 30", "1 -1 0.5 25"]


# ---- the node glue (WBC_WITH_ROS) against the stand-in ROS headers of tests/cpp/mock_ros
NODE_SRC = os.path.join(util.ROOT, "tests", "cpp", "ros_node_host.cpp")
NODE_EXE = os.path.join(util.ROOT, "tests", "cpp", "ros_node_host.bin")


@pytest.fixture(scope="module")
def node_exe():
    from wbc_quadruped_dob_b200 import build as B
    lib = B.build()
    inc = os.path.join(util.ROOT, "include")
    deps = [NODE_SRC, os.path.join(inc, "wbc_ros_adapter.hpp"), os.path.join(inc, "wbc_dogctrl.hpp"), os.path.join(inc, "wbc_b200.h")]
    if not os.path.exists(NODE_EXE) or any(os.path.getmtime(d) > os.path.getmtime(NODE_EXE) for d in deps):
        libdir = os.path.dirname(lib)
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + inc, "-I" + os.path.join(util.ROOT, "tests", "cpp", "mock_ros"),
                        NODE_SRC, "-o", NODE_EXE, "-L" + libdir, "-lwbc_b200", "-Wl,-rpath," + libdir], check=True)
    return NODE_EXE


def test_node_glue_compiles_against_the_stand_in_ros_headers_and_refuses_without_gpu(node_exe):
    import torch
    out = subprocess.run([node_exe, "probe"], capture_output=True, text=True, check=True).stdout.strip()
    assert out == ("ok" if torch.cuda.is_available() else "nodev")


@pytest.mark.gpu
def test_node_step_publishes_the_controllers_torques_in_publish_cmd_order(node_exe):
    """Messages in on the reference's topics (joint_states in Gazebo's alphabetical order, model_states with an unnormalised
    quaternion, four contact sensors), one stance step, and the published command is the torque vector of a DogCtrl driven
    directly with the same numbers, in publish_cmd's reverse message order (main.cpp:768-779); estimation_ee carries w."""
    out = subprocess.run([node_exe, "run"], capture_output=True, text=True, check=True).stdout.split()
    vals = dict(zip(out[0::2], out[1::2]))
    assert float(vals["cmd_err"]) == 0.0 and float(vals["est_err"]) == 0.0, vals
    assert int(vals["published"]) == 1 and int(vals["status"]) == 0 and float(vals["tau_max"]) > 0.1

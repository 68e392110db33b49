#!/usr/bin/env python3
"""Generate tests/golden/gazebo_states.npz from the one artefact the reference ships that records its robot in motion: the Gazebo
log /root/reference/DogBotV4/log/2020-05-13T09_33_15.915042/gzserver/state.log (decoded by tools/decode_gazebo_log.py).

The fixture holds (a) the DogBot model exactly as Gazebo/sdformat built it from dogbot.urdf -- the world SDF embedded in the log:
link poses at the zero configuration, lumped inertials, joint parents/children/axes -- and (b) 600 logged states: world pose and
twist of all 13 links.  It is independent of tools/gen_model.py (which generates the model tables of BOTH the kernels and the
oracle): tests/test_gazebo_pin.py rebuilds base pose, joint angles and rates from the logged link poses with this SDF tree and
checks that the oracle's and the kernels' forward kinematics land on the logged link frames.  The foot frame is not a link of
its own in the log (sdformat lumps fixed joints), so the foot joint's origin is read from dogbot.urdf here (urdf:320-325).

Run where /root/reference exists:  python tests/golden/make_gazebo_fixture.py"""
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import decode_gazebo_log as gz          # noqa: E402

LOG = "/root/reference/DogBotV4/log/2020-05-13T09_33_15.915042/gzserver/state.log"
URDF = "/root/reference/DogBotV4/ROS/src/dogbot_description/urdf/dogbot.urdf"
LEGS = ["back_left", "back_right", "front_left", "front_right"]      # canonical leg order BL, BR, FL, FR


def main():
    world = next(gz.chunks(LOG))
    sdf = ET.fromstring(world[world.index("<sdf"):])
    model = [m for m in sdf.iter("model") if m.get("name") == "dogbot"][0]
    links = {}
    for l in model.findall("link"):
        pose = np.array(l.find("pose").text.split(), dtype=np.float64)
        ine = l.find("inertial")
        ip = np.array(ine.find("pose").text.split(), dtype=np.float64)
        I = ine.find("inertia")
        links[l.get("name")] = dict(pose=pose, mass=float(ine.find("mass").text), com=ip,
                                    inertia=np.array([float(I.find(k).text) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")]))
    joints = {}
    for j in model.findall("joint"):
        ax = j.find("axis")
        joints[j.get("name")] = dict(parent=j.find("parent").text, child=j.find("child").text, type=j.get("type"),
                                     axis=np.array(ax.find("xyz").text.split(), dtype=np.float64),
                                     parent_frame=int(ax.find("use_parent_model_frame").text))
    names = sorted(links)
    urdf = ET.parse(URDF).getroot()
    foot = {}
    for j in urdf.findall("joint"):
        if j.get("name").endswith("_foot_joint"):
            foot[j.find("parent").get("link")] = np.array(j.find("origin").get("xyz").split(), dtype=np.float64)
            assert j.find("origin").get("rpy").split() == ["0", "0", "0"]
    # states: every 6th of the 3632 logged ones -> 606
    times, poses, vels = [], [], []
    k = 0
    for ci, ch in enumerate(gz.chunks(LOG)):
        if ci == 0:
            continue
        for t, st in gz.parse_states(ch):
            if k % 6 == 0 and set(st) >= set(names):
                times.append(t)
                poses.append([st[nm][0] for nm in names])
                vels.append([st[nm][1] for nm in names])
            k += 1
    jn = sorted(joints)
    out = os.path.join(ROOT, "tests", "golden", "gazebo_states.npz")
    np.savez_compressed(out, time=np.array(times), link_names=np.array(names), pose=np.array(poses), vel=np.array(vels),
                        link_pose0=np.array([links[n]["pose"] for n in names]), link_mass=np.array([links[n]["mass"] for n in names]),
                        link_com=np.array([links[n]["com"] for n in names]), link_inertia=np.array([links[n]["inertia"] for n in names]),
                        joint_names=np.array(jn), joint_parent=np.array([joints[n]["parent"] for n in jn]),
                        joint_child=np.array([joints[n]["child"] for n in jn]), joint_axis=np.array([joints[n]["axis"] for n in jn]),
                        joint_axis_in_parent_model_frame=np.array([joints[n]["parent_frame"] for n in jn]),
                        foot_offset=np.array([foot[l + "_lowerleg"] for l in LEGS]), legs=np.array(LEGS))
    print("wrote %s: %d states of %d, %d links, %d joints, total mass %.4f" % (out, len(times), k, len(names), len(jn), sum(links[n]["mass"] for n in names)))


if __name__ == "__main__":
    main()

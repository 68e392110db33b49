#!/usr/bin/env python3
"""Decode a Gazebo 7 `state.log` (the only artefact the reference ships that records its robot in motion:
/root/reference/DogBotV4/log/*/gzserver/state.log): <chunk encoding='zlib'> = base64 -> zlib -> XML.  The first chunk is the
world SDF, the following ones are <sdf><state> snapshots with every link's world pose (x y z roll pitch yaw), velocity,
acceleration and wrench.

    python tools/decode_gazebo_log.py LOG [--out fixtures.npz] [--every K] [--max N]

Writes, for the model `dogbot`, per sample: sim time, and for every link the world pose and twist.  TEST INFRASTRUCTURE ONLY
(the generated fixture under tests/golden/ pins the oracle's and the kernels' kinematic tree against the reference's own
simulation, independently of tools/gen_model.py)."""
import argparse
import base64
import re
import sys
import zlib
import xml.etree.ElementTree as ET

import numpy as np


def chunks(path):
    txt = open(path, "r", errors="ignore").read()
    for m in re.finditer(r"<chunk encoding='(\w+)'>\s*<!\[CDATA\[(.*?)\]\]>", txt, flags=re.S):
        enc, data = m.group(1), m.group(2)
        raw = base64.b64decode(data)
        yield zlib.decompress(raw).decode("utf-8", errors="ignore") if enc == "zlib" else raw.decode("utf-8", errors="ignore")


def parse_states(xml_text):
    """Yield (sim_time, {link: (pose6, vel6)}) for every <state> in a chunk."""
    # a chunk may hold several <sdf version=...><state>...</state></sdf> documents back to back
    for m in re.finditer(r"<state world_name=.*?</state>", xml_text, flags=re.S):
        st = ET.fromstring(m.group(0))
        tnode = st.find("sim_time")
        sec, nsec = tnode.text.split()
        t = int(sec) + 1e-9 * int(nsec)
        links = {}
        for model in st.findall("model"):
            if model.get("name") != "dogbot":
                continue
            for link in model.findall("link"):
                pose = np.array(link.find("pose").text.split(), dtype=np.float64)
                vel = np.array(link.find("velocity").text.split(), dtype=np.float64)
                links[link.get("name")] = (pose, vel)
        if links:
            yield t, links


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("log")
    ap.add_argument("--out", default=None)
    ap.add_argument("--every", type=int, default=1)
    ap.add_argument("--max", type=int, default=0)
    ap.add_argument("--dump-world", action="store_true", help="print the first chunk (world SDF) and exit")
    a = ap.parse_args()
    times, poses, vels, names = [], [], [], None
    k = 0
    for ci, ch in enumerate(chunks(a.log)):
        if ci == 0 and a.dump_world:
            print(ch)
            return 0
        for t, links in parse_states(ch):
            if names is None:
                names = sorted(links)
            if k % a.every == 0 and set(links) >= set(names):
                times.append(t)
                poses.append([links[nm][0] for nm in names])
                vels.append([links[nm][1] for nm in names])
            k += 1
            if a.max and len(times) >= a.max:
                break
        if a.max and len(times) >= a.max:
            break
    if names is None:
        print("no dogbot states found", file=sys.stderr)
        return 1
    times, poses, vels = np.array(times), np.array(poses), np.array(vels)
    print("states %d (of %d), links %d: %s" % (len(times), k, len(names), names))
    print("time span %.3f .. %.3f s" % (times[0], times[-1]))
    if a.out:
        np.savez_compressed(a.out, time=times, link_names=np.array(names), pose=poses, vel=vels)
        print("wrote", a.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())

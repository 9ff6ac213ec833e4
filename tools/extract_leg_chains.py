"""Extracts the 3-joint leg chains of a quadruped URDF into the constants the CUDA library needs.

Runs in the build container only (reads /root/reference); its output is committed as data in
robot-gym_b200/robot_gym/model/robots/descriptions.py and as tests/golden/leg_chains.json.

PyBullet conventions reproduced here:
  * the base frame (getBasePositionAndOrientation, robot.py:367-383) is the base link's INERTIAL
    frame, so joint origins of the base's children are shifted by the base inertial origin;
  * URDF rpy is fixed-axis XYZ: R = Rz(yaw) Ry(pitch) Rx(roll).

Usage: python tools/extract_leg_chains.py <urdf> <motor_name> x12  -> JSON on stdout
"""
import json
import math
import sys
import xml.etree.ElementTree as ET


def rpy_matrix(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp, cp * sr, cp * cr]


def parse(urdf_path, motor_names):
    root = ET.parse(urdf_path).getroot()
    joints = {j.get("name"): j for j in root.findall("joint")}
    by_parent = {}
    for j in root.findall("joint"):
        by_parent.setdefault(j.find("parent").get("link"), []).append(j)
    links = {l.get("name"): l for l in root.findall("link")}

    def origin(j):
        o = j.find("origin")
        xyz = [float(v) for v in (o.get("xyz") if o is not None and o.get("xyz") else "0 0 0").split()]
        rpy = [float(v) for v in (o.get("rpy") if o is not None and o.get("rpy") else "0 0 0").split()]
        return xyz, rpy_matrix(*rpy)

    def axis(j):
        a = j.find("axis")
        return [float(v) for v in (a.get("xyz") if a is not None else "1 0 0").split()]

    legs = []
    for leg in range(4):
        names = motor_names[3 * leg:3 * leg + 3]
        chain = {"p": [], "r": [], "axis": [], "joint_names": names}
        for n, name in enumerate(names):
            j = joints[name]
            xyz, rot = origin(j)
            if n == 0:
                base = links[j.find("parent").get("link")]
                inert = base.find("inertial")
                io = inert.find("origin") if inert is not None else None
                ixyz = [float(v) for v in (io.get("xyz") if io is not None and io.get("xyz") else "0 0 0").split()]
                irpy = [float(v) for v in (io.get("rpy") if io is not None and io.get("rpy") else "0 0 0").split()]
                assert all(abs(v) < 1e-12 for v in irpy), "rotated base inertial frame not handled"
                xyz = [a - b for a, b in zip(xyz, ixyz)]
            else:
                assert j.find("parent").get("link") == joints[names[n - 1]].find("child").get("link")
            chain["p"].append(xyz)
            chain["r"].append(rot)
            chain["axis"].append(axis(j))
        lower_link = joints[names[2]].find("child").get("link")
        toe_joints = [j for j in by_parent.get(lower_link, []) if j.get("type") == "fixed"]
        assert len(toe_joints) == 1, lower_link
        txyz, trot = origin(toe_joints[0])
        chain["toe"] = txyz
        chain["toe_joint"] = toe_joints[0].get("name")
        legs.append(chain)
    return legs


if __name__ == "__main__":
    print(json.dumps(parse(sys.argv[1], sys.argv[2:14]), indent=1))

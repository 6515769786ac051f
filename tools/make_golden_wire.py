"""Records tests/golden/wire_msgs.npz: the ROS 1 wire bytes of planning_ros_msgs/Trajectory for a few plans, produced by a
route that shares nothing with the product's serialiser (mplb.cu k_serialize_traj) nor with the hand-written packer of
tests/test_gpu_wire.py:

  * the trajectory is planned by the REFERENCE'S OWN planner sources (oracle/_ref/libmplref.so) and its coefficient rows
    are read from the reference's own Primitive objects (ref_get_traj_coeffs = Primitive1D::coeff(), primitive.h:125);
  * the message is assembled as planning_ros_utils/primitive_ros_utils.h:11-55,78-113 (toPrimitiveROSMsg /
    toTrajectoryROSMsg) does: cx, cy, cz, cyaw = the six coefficients, cz = (0,0,0,0,0,z) for 2D, t = pr.t(), lambda empty;
  * the bytes come from a GENERIC ROS 1 serialiser driven by the text of the .msg files under
    /root/reference/planning_ros_msgs/msg (genpy's rules: fields in declaration order, little endian, uint32 length before
    variable arrays and strings, nested messages inline; std_msgs/Header = uint32 seq, time stamp (2 x uint32), string
    frame_id is a ROS built-in and is declared here).

Run in the build container: python tools/make_golden_wire.py"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from mpl_ros_b200 import maps  # noqa: E402
from helpers import load_config  # noqa: E402

MSG_DIR = "/root/reference/planning_ros_msgs/msg"
BUILTIN = {"bool": "<B", "int8": "<b", "uint8": "<B", "byte": "<b", "char": "<B", "int16": "<h", "uint16": "<H", "int32": "<i",
           "uint32": "<I", "int64": "<q", "uint64": "<Q", "float32": "<f", "float64": "<d"}
HEADER = [("uint32", "seq"), ("time", "stamp"), ("string", "frame_id")]  # std_msgs/Header (ROS built-in)


def parse_msg(name):
    if name in ("Header", "std_msgs/Header"):
        return HEADER
    fields = []
    for line in open(os.path.join(MSG_DIR, name.split("/")[-1] + ".msg")):
        line = line.split("#")[0].strip()
        if not line or "=" in line:
            continue
        t, n = line.split()[:2]
        fields.append((t, n))
    return fields


def ser(t, v):
    if t.endswith("]"):
        base, dimspec = t[:t.index("[")], t[t.index("[") + 1:-1]
        out = b"" if dimspec else struct.pack("<I", len(v))
        return out + b"".join(ser(base, x) for x in v)
    if t in BUILTIN:
        return struct.pack(BUILTIN[t], v)
    if t == "string":
        b = v.encode()
        return struct.pack("<I", len(b)) + b
    if t in ("time", "duration"):
        return struct.pack("<II", v[0], v[1])
    return b"".join(ser(ft, v[fn]) for ft, fn in parse_msg(t))


def trajectory_msg(coeffs, dim, dt, z, frame_id, seq, stamp):
    """toTrajectoryROSMsg (primitive_ros_utils.h:78-113) over coefficient rows [n_seg, 4, 6] of the reference's primitives."""
    prims = []
    for c in coeffs:
        cz = [0, 0, 0, 0, 0, z] if dim == 2 else list(c[2])
        prims.append({"cx": list(c[0]), "cy": list(c[1]), "cz": [float(x) for x in cz], "cyaw": list(c[3]), "t": float(dt)})
    return {"header": {"seq": seq, "stamp": stamp, "frame_id": frame_id}, "primitives": prims, "lambda": []}


CASES = {
    # name: (config, control, start/goal swap, U override, params override, z)
    "corridor_fwd": ("corridor", 3, False, None, {}, 0.25),
    "corridor_back": ("corridor", 3, True, None, {}, 0.25),
    "skir_acc": ("skir", 3, False, None, {}, 0.0),
    "skir_jrk_eps2": ("skir", 7, False, None, dict(v_max=2.0, a_max=2.0, dt=1.0, tol_pos=0.5, max_num=20000, epsilon=2.0), 0.0),
    "corridor_jrk_eps2": ("corridor", 7, False, None, dict(v_max=1.0, a_max=1.0, j_max=2.0, dt=1.0, tol_pos=0.5, max_num=30000, epsilon=2.0), 0.1),
    "simple_acc": ("simple", 3, False, None, {}, 0.0),
}


def main():
    assert ref.available()
    out = {}
    for name, (cfgname, control, swap, Uo, po, z) in CASES.items():
        m, dim, params, U, start, goal = load_config(cfgname)
        params = dict(params, **po)
        if swap:
            start, goal = goal, start
        rm = ref.RefMap(m.origin, m.dim, m.data, m.res)
        rm.free_unknown()
        rp = ref.RefPlanner(dim)
        rp.set_map(rm)
        for k, v in params.items():
            rp.set_param(k, v)
        rp.set_controls(U)
        s, g = oracle.make_waypoints(1), oracle.make_waypoints(1)
        s["pos"][0, :dim], g["pos"][0, :dim], s["control"], g["control"] = start, goal, control, control
        r = rp.plan(s, g)
        assert r["status"] == 0, name
        coeffs = rp.traj_coeffs(int(r["n_seg"]))
        msg = trajectory_msg(coeffs, dim, params["dt"], z, "map", 7, (12, 345))
        b = ser("Trajectory", msg)
        out[name + "/bytes"] = np.frombuffer(b, dtype=np.uint8)
        out[name + "/n_seg"] = np.int32(r["n_seg"])
        print(name, r["n_seg"], len(b))
    path = os.path.join(ROOT, "tests", "golden", "wire_msgs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()

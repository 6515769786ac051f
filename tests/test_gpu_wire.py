"""GPU test of SURVEY section 8(f).4, wire output: batched planning_ros_msgs/Trajectory serialisation
(planning_ros_utils/primitive_ros_utils.h:11-55,78-113) against an independent pure-Python ROS 1 serialiser fed with
the ORACLE's trajectories.  ROS itself is not in this image; the wire rules used here are the published ROS 1 ones
(little endian, uint32 length prefix for strings and variable arrays, fields in declaration order)."""
import math
import struct

import numpy as np
import pytest

import mpl_ros_b200 as mp
from mpl_ros_b200 import maps
from helpers import load_config
from helpers_gpu import make_pair, waypoint_pair

pytestmark = pytest.mark.gpu


def _ros_trajectory_bytes(dim, control, actions, seg_states, U, dt, z, frame_id, seq, stamp):
    """std_msgs/Header + Primitive[] + LambdaSeg[] (empty), built from (parent state, U[action], dt) like
    env_base::forward_action (env_base.h:228-231) and Primitive1D's constructors (primitive.h:35-52)."""
    order = {1: 1, 3: 2, 7: 3, 15: 4}[control & 15]
    b = struct.pack("<III", seq, stamp[0], stamp[1]) + struct.pack("<I", len(frame_id)) + frame_id.encode()
    b += struct.pack("<I", len(actions))
    for a, st in zip(actions, seg_states):
        rows = []
        for ax in range(3):
            c = [0.0] * 6
            if ax < dim:
                for d in range(order):
                    c[5 - d] = st[d * 3 + ax]
                c[5 - order] = U[a][ax]
            elif ax == 2:
                c[5] = z
            rows.append(c)
        cyaw = [0.0] * 6
        if control & 16:
            cyaw[4], cyaw[5] = U[a][dim], st[12]
        rows.append(cyaw)
        for c in rows:
            b += struct.pack("<I", 6) + struct.pack("<6d", *c)
        b += struct.pack("<d", dt)
    b += struct.pack("<I", 0)
    return b


def _check(pl, op, sg, gg, so, go, dim, control, U, dt, n, max_seg, z):
    rg, ag, segs = pl.plan_batch(sg, gg, max_seg=max_seg, want_states=True)
    msgs = pl.serialize_trajectories(rg, ag, segs, z=z, frame_id="map", seq=7, stamp=(12, 345))
    n_ok = 0
    for i in range(n):
        ro = op.plan(so[i:i + 1], go[i:i + 1])
        assert ro["status"] == rg[i]["status"]
        ns = int(ro["n_seg"]) if ro["status"] == 0 else 0
        if ns > max_seg:
            assert msgs[i] is None
            continue
        exp = _ros_trajectory_bytes(dim, control, op.actions(ns) if ns else [], op.seg_states(ns) if ns else [], U, dt, z, "map",
                                    7, (12, 345))
        assert msgs[i] == exp, i
        n_ok += ns > 0
    return n_ok


def test_wire_3d_batch():
    m = maps.load_fixture("levine")
    U = maps.make_U(1.0, 1, 3)
    pl, op = make_pair(m, 3, dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5), U)
    n = 32
    S, G = maps.sample_queries(m, n, seed=5)
    sg, so = waypoint_pair(S, mp.ACC)
    gg, go = waypoint_pair(G, mp.ACC)
    assert _check(pl, op, sg, gg, so, go, 3, mp.ACC, U, 1.0, n, 12, 0.0) >= 4  # max_seg 12 also exercises truncation


def test_wire_2d_with_z_and_yaw():
    m, dim, params, U2, start, goal = load_config("corridor")
    pl, op = make_pair(m, dim, params, U2)
    sg, so = waypoint_pair([start, goal], mp.ACC)
    gg, go = waypoint_pair([goal, start], mp.ACC)
    assert _check(pl, op, sg, gg, so, go, 2, mp.ACC, U2, params["dt"], 2, 64, 0.25) == 2
    U = np.array([[dx, dy, dyaw] for dx in (-0.5, 0, 0.5) for dy in (-0.5, 0, 0.5) for dyaw in (-0.5, 0, 0.5)])
    pl, op = make_pair(m, dim, dict(params, yaw_max=0.7), U)
    op.set_param("trig_mode", 1)
    sg, so = waypoint_pair([start], mp.ACCxYAW, yaw=math.pi / 2)
    gg, go = waypoint_pair([goal], mp.ACCxYAW)
    assert _check(pl, op, sg, gg, so, go, 2, mp.ACCxYAW, U, params["dt"], 1, 64, 0.1) == 1


def test_wire_jrk_3d():
    m, dim, params, U, start, goal = load_config("skir")
    Uj = maps.make_U(1.0, 1, 3)
    pl, op = make_pair(m, 3, dict(v_max=2.0, a_max=2.0, dt=1.0, tol_pos=0.5, max_num=4000), Uj)
    sg, so = waypoint_pair([start], mp.JRK)
    gg, go = waypoint_pair([goal], mp.JRK)
    _check(pl, op, sg, gg, so, go, 3, mp.JRK, Uj, 1.0, 1, 64, 0.0)


def test_wire_against_reference_generated_bytes():
    """Independent expectation (tests/golden/wire_msgs.npz, tools/make_golden_wire.py): trajectories planned by the
    reference's own sources, coefficient rows read from the reference's Primitive objects, message assembled as
    primitive_ros_utils.h does and serialised by a generic ROS 1 serialiser driven by the text of
    planning_ros_msgs/msg/{Trajectory,Primitive,LambdaSeg}.msg.  The GPU plans the same queries and writes the bytes itself."""
    import os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools")
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_msgs.npz"))
    src = open(os.path.join(sys_path, "make_golden_wire.py")).read()
    cases = eval(src[src.index("CASES = {") + len("CASES = "):src.index("}\n\n\ndef main") + 1])  # the generator's own case table
    for name, (cfgname, control, swap, _, po, z) in cases.items():
        m, dim, params, U, start, goal = load_config(cfgname)
        params = dict(params, **po)
        if swap:
            start, goal = goal, start
        pl, _ = make_pair(m, dim, params, U)
        sg, _ = waypoint_pair([start], control)
        gg, _ = waypoint_pair([goal], control)
        rg, ag, segs = pl.plan_batch(sg, gg, max_seg=64, want_states=True)
        assert rg[0]["status"] == 0 and rg[0]["n_seg"] == int(gold[name + "/n_seg"]), name
        msgs = pl.serialize_trajectories(rg, ag, segs, z=z, frame_id="map", seq=7, stamp=(12, 345))
        assert msgs[0] == gold[name + "/bytes"].tobytes(), name

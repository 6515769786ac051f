"""GPU parity tests of SURVEY section 8(f).2: yaw controls (Control::*xYAW; primitive.h:236-253,503-525, waypoint.h:114-117,
env_map.h:121-128) against the oracle.

Tolerance policy: the reference evaluates cos/sin with an unpinned libm; the product defines the branch with the
correctly rounded functions, and the oracle is run in the same definition (trig_mode 1), so every comparison here is
still bit-exact.  tests/test_oracle_yaw.py (CPU) bounds the distance between that definition and libm.
The reference publishes no numbers for its yaw tests (they draw pictures); the oracle's yaw branch (libm definition) is
pinned by the reference's own sources in tests/test_oracle_vs_reference.py.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle
import mpl_ros_b200 as mp
from mpl_ros_b200 import _lib, maps
from helpers import load_config
from helpers_gpu import assert_results_equal, make_pair, waypoint_pair

pytestmark = pytest.mark.gpu


def _U_yaw(dim, u, u_yaw, z=False):
    """test_planner_2d_with_yaw.cpp:49-57: {-u,0,u}^2 x {-u_yaw,0,u_yaw} (3D: z rate fixed at 0)."""
    rows = []
    for dx in (-u, 0.0, u):
        for dy in (-u, 0.0, u):
            for dyaw in (-u_yaw, 0.0, u_yaw):
                rows.append([dx, dy, dyaw] if dim == 2 else [dx, dy, 0.0, dyaw])
    return np.array(rows)


def _nodes_by_key(nodes):
    return {tuple(n["key"][:n["key"][15]]): n for n in nodes}


def _full_compare(pl, op, sg, gg, so, go, ctx, ns):
    ok = pl.plan(sg, gg)
    ro = op.plan(so, go)
    rg = pl.result()
    assert_results_equal(rg, ro, ctx)
    assert ok == (ro["status"] in (0, 5))
    gn = pl.getNodes()
    assert np.array_equal(gn["key"][pl.getPopLog()], op.pop_keys(ro["pops"]))
    on, gnk = _nodes_by_key(op.nodes(ro["n_nodes"])), _nodes_by_key(gn)
    assert set(on) == set(gnk)
    for k, a in gnk.items():
        b = on[k]
        assert np.array_equal(a["state"][:ns], b["state"][:ns]) and a["state"][12] == b["state"][12], k
        assert a["g"] == b["g"] and a["h"] == b["h"] and a["opened"] == b["opened"] and a["closed"] == b["closed"], k
    if ro["status"] == 0:
        assert np.array_equal(pl.getActions(), op.actions(ro["n_seg"]))
        sa, sb = pl.getSegStates(), op.seg_states(ro["n_seg"])
        assert np.array_equal(sa[:, :ns], sb[:, :ns]) and np.array_equal(sa[:, 12], sb[:, 12])
    return rg


def test_sincos_cr_device_matches_oracle():
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-math.pi, math.pi, 50000), rng.uniform(-50, 50, 5000),
                        [0.0, math.pi, -math.pi, math.pi / 2, 0.7, 1e-300, 1e-9, 3.0, 12345.678],
                        math.pi / 2 + 0.5 * np.arange(-8, 9), 0.05 * np.arange(-63, 64)])
    x = np.ascontiguousarray(x)
    s, c = np.zeros_like(x), np.zeros_like(x)
    _lib.check(_lib.lib().mplb_sincos_cr(_lib.ptr(x), x.size, _lib.ptr(s), _lib.ptr(c)))
    so, co = np.zeros_like(x), np.zeros_like(x)
    oracle.lib().orc_sincos_cr(x.ctypes.data_as(C.c_void_p), x.size, so.ctypes.data_as(C.c_void_p),
                               co.ctypes.data_as(C.c_void_p))
    assert np.array_equal(s, so) and np.array_equal(c, co)
    assert np.max(np.abs(s - np.sin(x))) < 2.3e-16 and np.max(np.abs(c - np.cos(x))) < 2.3e-16  # within an ulp of libm


@pytest.mark.parametrize("yaw_max,wyaw", [(0.7, 1.0), (-1.0, 1.0), (0.7, 0.0), (1.2, 2.5)])
def test_planner_2d_with_yaw(yaw_max, wyaw):
    """MPL/test/test_planner_2d_with_yaw.cpp:28-69 on corridor.yaml: ACCxYAW, start yaw pi/2, 27 controls."""
    m, dim, params, _, start, goal = load_config("corridor")
    U = _U_yaw(2, 0.5, 0.5)
    pl, op = make_pair(m, dim, dict(params, yaw_max=yaw_max, wyaw=wyaw), U)
    op.set_param("trig_mode", 1)
    sg, so = waypoint_pair(start, mp.ACCxYAW, yaw=math.pi / 2)
    gg, go = waypoint_pair(goal, mp.ACCxYAW)
    rg = _full_compare(pl, op, sg, gg, so, go, ("yaw", yaw_max, wyaw), 6)
    assert rg["status"] == 0
    if (yaw_max, wyaw) == (0.7, 1.0):  # regression pin of the oracle's own answer (same in libm and CR mode)
        assert rg["n_seg"] == 35 and rg["pops"] == 1342 and abs(rg["cost"] - 352.4275550988982) < 1e-9
    traj = pl.getTraj()
    ws = traj.getWaypoints()
    assert len(ws) == rg["n_seg"] + 1 and ws[0].yaw == math.pi / 2


def test_vel_yaw_2d_and_acc_yaw_3d():
    m, dim, params, _, start, goal = load_config("corridor")
    U = _U_yaw(2, 1.0, 0.4)
    pl, op = make_pair(m, dim, dict(dt=1.0, tol_pos=0.5, yaw_max=0.9, w=10.0, v_max=1.0), U)
    op.set_param("trig_mode", 1)
    sg, so = waypoint_pair(start, mp.VELxYAW, yaw=0.3)
    gg, go = waypoint_pair(goal, mp.VELxYAW)
    rg = _full_compare(pl, op, sg, gg, so, go, "VELxYAW", 3)
    assert rg["status"] == 0

    m, dim, params, _, start, goal = load_config("skir")
    U = _U_yaw(3, 1.0, 0.5)
    pl, op = make_pair(m, dim, dict(params, yaw_max=0.8), U)
    op.set_param("trig_mode", 1)
    sg, so = waypoint_pair(start, mp.ACCxYAW, yaw=-2.0)
    gg, go = waypoint_pair(goal, mp.ACCxYAW)
    _full_compare(pl, op, sg, gg, so, go, "ACCxYAW 3D (planar controls)", 6)


def test_distance_map_planner_2d_with_yaw():
    """MPL/test/test_distance_map_planner_2d_with_yaw.cpp: plain ACC plan, then tunnel + potential map + yaw."""
    m, dim, params, U2, start, goal = load_config("corridor")
    pl, op = make_pair(m, dim, params, U2)
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    assert pl.plan(sg, gg) and op.plan(so, go)["status"] == 0
    path = [np.array(w.pos) for w in pl.getTraj().getWaypoints()]
    opath = np.zeros((len(path), 3))
    opath[:, :2] = np.array(path)
    mu, om = pl._keep
    U = _U_yaw(2, 0.5, 0.5)
    pl2, op2 = make_pair(m, dim, dict(params, yaw_max=0.7), U)
    op2.set_param("trig_mode", 1)
    pl2.setMapUtil(mu)
    op2.set_map(om)
    pl2.setSearchRadius([0.5, 0.5])
    op2.set_vec("search_radius", [0.5, 0.5, 0.0])
    pl2.setSearchRegion(path)
    op2.set_search_region(opath, dense=False)
    pl2.setPotentialRadius([1.0, 1.0])
    op2.set_vec("potential_radius", [1.0, 1.0, 0.0])
    pl2.setPotentialWeight(0.5)
    op2.set_param("potential_weight", 0.5)
    pl2.setGradientWeight(0.1)
    op2.set_param("gradient_weight", 0.1)
    pl2.updatePotentialMap(start)
    op2.update_potential_map(np.array([start[0], start[1], 0.0]))
    sg, so = waypoint_pair(start, mp.ACCxYAW, yaw=math.pi / 2)
    gg, go = waypoint_pair(goal, mp.ACCxYAW)
    rg = _full_compare(pl2, op2, sg, gg, so, go, "distance map with yaw", 6)
    assert rg["status"] == 0


def test_yaw_batch_parity():
    m = maps.load_fixture("levine")
    U = _U_yaw(3, 1.0, 0.5)
    params = dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5, yaw_max=1.0)
    pl, op = make_pair(m, 3, params, U)
    op.set_param("trig_mode", 1)
    n = 48
    S, G = maps.sample_queries(m, n, seed=11)
    S[:, 2] = G[:, 2] = S[0, 2]  # planar controls: keep start and goal on one height
    yaws = np.linspace(-3.0, 3.0, n)
    sg, so = waypoint_pair(S, mp.ACCxYAW, yaw=yaws)
    gg, go = waypoint_pair(G, mp.ACCxYAW)
    rg, ag, _ = pl.plan_batch(sg, gg, max_seg=64, want_states=True)
    ro, ao = op.plan_batch(so, go, nthreads=8, max_seg=64)
    for i in range(n):
        assert_results_equal(rg[i], ro[i], ("yaw batch", i))
    assert np.array_equal(ag, ao)
    assert (ro["status"] == 0).sum() >= 4


def test_yaw_errors():
    m, dim, params, U2, start, goal = load_config("corridor")
    pl, _ = make_pair(m, dim, params, U2)  # rows without a yaw column
    sg, _ = waypoint_pair(start, mp.ACCxYAW, yaw=0.1)
    gg, _ = waypoint_pair(goal, mp.ACCxYAW)
    with pytest.raises(mp.MplbError):
        pl.plan(sg, gg)

"""GPU TrajSolver (mplb_traj_solve_batch, mpl_ros_b200/csrc/mplb_trajsolve.cu) against the oracle and the fixture recorded
from the reference's own sources.  The kernel performs the dense algorithm's operations in the dense algorithm's order on
the non-zero blocks only, so the comparison is EXACT (tolerance 0) against oracle/poly_oracle.cpp and against
tests/golden/trajsolver.npz.  Against a real Eigen build the expected difference is Eigen's blocked LU / GEMM for matrices
larger than 16 x 16 (rounding level); the north-star tolerance for trajectory coefficients is 1e-6 relative, and
test_conditioning_margin shows the margin: perturbing every input by 1 ulp moves the coefficients by < 1e-9 relative."""
import os

import numpy as np
import pytest

import mpl_ros_b200 as mp
from mpl_ros_b200 import traj_solver
import oracle
from trajsolver_cases import ACC, JRK, SNP, VEL, cases, random_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "trajsolver.npz")


def test_batch_equals_oracle_and_golden():
    gold = np.load(GOLD)
    groups = {}
    for c in cases():
        groups.setdefault((c[1], c[2], c[3]), []).append(c)
    n = 0
    for (dim, control, yaw_control), cs in groups.items():  # one launch per (dim, control, yaw order): mixed lengths inside
        got = traj_solver.solve_batch(dim, control, [c[4] for c in cs], [c[5] for c in cs], yaw_control)
        for c, g in zip(cs, got):
            want = oracle.traj_solve(dim, control, c[4], c[5], yaw_control)
            assert g.shape == want.shape, c[0]
            assert np.array_equal(g, want), (c[0], np.abs(g - want).max())
            assert np.array_equal(g, gold[c[0]]), c[0]
            n += 1
    assert n == len(gold.files)


def test_large_batch_random():
    rs = np.random.RandomState(21)
    for dim, control in ((3, JRK), (2, ACC), (3, VEL)):
        ws, ds = [], []
        for i in range(300):
            w, d = random_case(rs, dim, int(rs.randint(2, 40)), control, (VEL, ACC, JRK) if i % 5 == 0 else (VEL,), yaw=True)
            ws.append(w)
            ds.append(d)
        got = traj_solver.solve_batch(dim, control, ws, ds, ACC)
        for i in range(0, 300, 7):
            assert np.array_equal(got[i], oracle.traj_solve(dim, control, ws[i], ds[i], ACC)), (dim, control, i)


def test_empty_results_and_short_lists():
    rs = np.random.RandomState(3)
    w, d = random_case(rs, 3, 5, JRK)
    assert len(traj_solver.solve_batch(3, SNP, [w], [d])[0]) == 0            # no solver for SNP (traj_solver.h:28-30)
    assert len(traj_solver.solve_batch(3, JRK, [w], [d], yaw_control=SNP)[0]) == 0
    got = traj_solver.solve_batch(3, JRK, [w[:1], w, w[:0], w[:2]], [d[:0], d, d[:0], d[:1]])
    assert [len(g) for g in got] == [0, 4, 0, 1]
    assert np.array_equal(got[1], oracle.traj_solve(3, JRK, w, d))
    assert np.array_equal(got[3], oracle.traj_solve(3, JRK, w[:2], d[:1]))
    assert traj_solver.solve_batch(3, JRK, [], []) == []


def test_reference_shaped_flow():
    """MPL/test/test_traj_solver.cpp through the mirror class: setPath, setV(1), solve."""
    path = [(0, 0), (1, 0), (2, 1), (5, 1)]
    gold = np.load(GOLD)
    for cname, c in (("VEL", VEL), ("ACC", ACC), ("JRK", JRK)):
        ts = mp.TrajSolver2D(c)
        ts.setPath(path)
        ts.setV(1)
        traj = ts.solve()
        assert ts.getDts() == [1.0, 1.0, 3.0] and traj.getTotalTime() == 5.0
        got = np.array([np.vstack([pr.coeffs, pr.yaw_coeff]) for pr in traj.getPrimitives()])
        assert np.array_equal(got, gold["test_traj_solver_%s" % cname])
        ws = traj.getWaypoints()  # the spline passes through the key frames
        for w, p in zip(ws, path):
            assert np.allclose(w.pos, p, atol=1e-9)


def test_refines_a_planned_trajectory():
    """map_planner_node.cpp:216-227: plan, take the trajectory's waypoints (interior -> Control::VEL) and segment times,
    TrajSolver3D(Control::JRK); the refined trajectory equals the oracle's for the same inputs."""
    from mpl_ros_b200 import maps
    m = maps.load_fixture("skir")
    mu = mp.VoxelMapUtil()
    mu.setMap(m.origin, m.dim, m.data, m.res)
    mu.freeUnknown()
    pl = mp.VoxelMapPlanner(False)
    pl.setMapUtil(mu)
    pl.setVmax(2.0); pl.setAmax(1.0); pl.setDt(1.0); pl.setU(maps.make_U(1.0, 1, 3)); pl.setTol(0.5)
    s, g = mp.waypoints_array(1), mp.waypoints_array(1)
    s["pos"][0], g["pos"][0] = (5.5, 5.5, 0.5), (1.5, 1.5, 5.5)
    s["control"] = g["control"] = mp.ACC
    assert pl.plan(s, g)
    traj = pl.getTraj()
    ws = traj.getWaypoints()
    for w in ws[1:-1]:
        w.control = mp.VEL
    dts = [pr.t() for pr in traj.getPrimitives()]
    ts = mp.TrajSolver3D(mp.JRK)
    ts.setWaypoints(ws)
    ts.setDts(dts)
    refined = ts.solve()
    assert len(refined.getPrimitives()) == len(dts)
    rec = np.zeros(len(ws), dtype=oracle.WAYPOINT_DTYPE)
    for i, w in enumerate(ws):
        w.to_record(rec[i])
    want = oracle.traj_solve(3, mp.JRK, rec, dts)
    got = np.array([np.vstack([pr.coeffs, pr.yaw_coeff]) for pr in refined.getPrimitives()])
    assert np.array_equal(got, want)
    for w, r in zip(ws, refined.getWaypoints()):  # same key frames
        assert np.allclose(w.pos, r.pos, atol=1e-8)


def test_conditioning_margin():
    rs = np.random.RandomState(8)
    w, d = random_case(rs, 3, 30, JRK, (VEL,), yaw=True)
    base = traj_solver.solve_batch(3, JRK, [w], [d])[0]
    w2, d2 = w.copy(), np.nextafter(d, np.inf)
    for f in ("pos", "vel", "acc"):
        w2[f] = np.nextafter(w[f], np.inf)
    pert = traj_solver.solve_batch(3, JRK, [w2], [d2])[0]
    rel = np.abs(pert - base).max() / np.abs(base).max()
    assert rel < 1e-9, rel


def test_refine_batch_on_device():
    """plan_batch -> mplb_refine_trajectories (gather kernel + batched solve): every successful plan's refined trajectory equals
    the oracle's TrajSolver on the waypoints the reference's node would build (map_planner_node.cpp:216-227); the end waypoint is
    the last primitive evaluated at dt, taken here from the oracle's get_succ row of the last parent state."""
    from mpl_ros_b200 import maps
    m = maps.load_fixture("skir")
    U = maps.make_U(1.0, 1, 3)
    mu = mp.VoxelMapUtil()
    mu.setMap(m.origin, m.dim, m.data, m.res)
    mu.freeUnknown()
    pl = mp.VoxelMapPlanner(False)
    pl.setMapUtil(mu)
    pl.setVmax(2.0); pl.setAmax(1.0); pl.setDt(1.0); pl.setU(U); pl.setTol(0.5)
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    op = oracle.OraclePlanner(3)
    op.set_map(om)
    for k, v in dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5).items():
        op.set_param(k, v)
    op.set_controls(U)
    starts, goals = maps.sample_queries(m, 24, seed=5, min_dist=1.5)
    s, g = mp.waypoints_array(24), mp.waypoints_array(24)
    s["pos"], g["pos"] = starts, goals
    s["control"] = g["control"] = mp.ACC
    max_seg = 16
    res, acts, segs = pl.plan_batch(s, g, max_seg=max_seg, want_states=True)
    coefs, nseg = pl.refine_trajectories(res, acts, segs, mp.ACC, mp.JRK)
    n_ok = 0
    for i in range(24):
        good = res[i]["status"] == 0 and 1 <= res[i]["n_seg"] <= max_seg
        assert nseg[i] == (res[i]["n_seg"] if good else 0)
        if not good:
            assert not coefs[i].any()
            continue
        ns = int(res[i]["n_seg"])
        w = oracle.make_waypoints(ns + 1)
        for j in range(ns):
            st = segs[i, j]
            w["pos"][j], w["vel"][j], w["acc"][j], w["jrk"][j], w["yaw"][j] = st[0:3], st[3:6], st[6:9], st[9:12], st[12]
        last = oracle.make_waypoints(1)
        last["pos"][0], last["vel"][0], last["acc"][0] = segs[i, ns - 1, 0:3], segs[i, ns - 1, 3:6], segs[i, ns - 1, 6:9]
        last["control"] = mp.ACC
        end = op.succ_trace(last)[acts[i, ns - 1]]["succ"]
        w["pos"][ns], w["vel"][ns], w["acc"][ns], w["jrk"][ns], w["yaw"][ns] = end[0:3], end[3:6], end[6:9], end[9:12], end[12]
        w["control"] = mp.VEL
        w["control"][0] = w["control"][ns] = mp.ACC
        want = oracle.traj_solve(3, mp.JRK, w, np.ones(ns))
        assert np.array_equal(coefs[i, :ns], want), i
        assert not coefs[i, ns:].any()
        n_ok += 1
    assert n_ok >= 5

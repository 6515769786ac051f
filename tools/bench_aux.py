"""Secondary measurements of the two rows widened in round 2's last session (not the headline metric; bench.py is):
  * TrajSolver: refine a batch of planned-trajectory-shaped waypoint lists (36 waypoints, JRK) on the GPU through the public
    host-buffer call (H2D + kernel + D2H inside the timed region) vs the CPU checker on one core;
  * LPA*: the skir replanning flow (tests/lpa_flow.py) per call, GPU vs the CPU checker, and 64 replanners in one launch.
Prints one JSON object.  Run on a GPU box:  python tools/bench_aux.py > gpurun_out/aux.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpl_ros_b200 as mp  # noqa: E402
from mpl_ros_b200 import _lib, traj_solver  # noqa: E402
import oracle  # noqa: E402
import lpa_flow  # noqa: E402
from trajsolver_cases import JRK, VEL, random_case  # noqa: E402
from test_gpu_lpa import GpuMap, GpuPlanner  # noqa: E402


def trajsolve():
    rs = np.random.RandomState(0)
    n, W = 4096, 36
    ws, ds = [], []
    for _ in range(n):
        w, d = random_case(rs, 3, W, JRK, (VEL,), yaw=True)
        ws.append(w)
        ds.append(d)
    traj_solver.solve_batch(3, JRK, ws[:8], ds[:8])  # warm-up
    t = []
    for _ in range(5):
        t0 = time.perf_counter()
        got = traj_solver.solve_batch(3, JRK, ws, ds)
        t.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    m = 64
    for i in range(m):
        want = oracle.traj_solve(3, JRK, ws[i], ds[i])
    cpu = (time.perf_counter() - t0) / m
    assert np.array_equal(got[m - 1], want)
    return dict(batch=n, waypoints=W, control="JRK", gpu_e2e_ms_per_batch=min(t) * 1e3, gpu_traj_per_s=n / min(t),
                cpu_checker_ms_per_traj_1core=cpu * 1e3, cpu_checker_traj_per_s_1core=1 / cpu,
                note="GPU time includes packing in Python, H2D, the kernel and D2H")


def lpa():
    out = {}
    for cls_map, cls_pl, tag in ((GpuMap, GpuPlanner, "gpu"), (oracle.OracleMap, oracle.OraclePlanner, "cpu_checker")):
        m, mp_, pl, dim, start, goal = lpa_flow.build(cls_map, cls_pl, "skir")
        pl._lpa_control = mp.ACC
        s, g = oracle.make_waypoints(1), oracle.make_waypoints(1)
        lpa_flow.fill_waypoints(s, start, mp.ACC)
        lpa_flow.fill_waypoints(g, goal, mp.ACC)
        t0 = time.perf_counter(); r = pl.lpa_plan(s, g); t_first = time.perf_counter() - t0
        t0 = time.perf_counter(); pl.lpa_get_linked_nodes(); t_link = time.perf_counter() - t0
        path = pl.lpa_best_child_states()[:, :3]
        cells = lpa_flow.cells_on_path(m, dim, path[len(path) // 2:len(path) // 2 + 1], 2)
        mp_.set_cells(cells, 100)
        t0 = time.perf_counter(); pl.lpa_update_blocked_nodes(cells); t_upd = time.perf_counter() - t0
        t0 = time.perf_counter(); r2 = pl.lpa_plan(s, g); t_re = time.perf_counter() - t0
        t0 = time.perf_counter(); pl.lpa_get_sub_state_space(1); t_sub = time.perf_counter() - t0
        out[tag] = dict(first_plan_ms=t_first * 1e3, first_pops=int(r["pops"]), link_ms=t_link * 1e3, update_ms=t_upd * 1e3,
                        replan_ms=t_re * 1e3, replan_pops=int(r2["pops"]), subtree_ms=t_sub * 1e3)
    # 64 robots, one launch
    n = 64
    pls = [lpa_flow.build(GpuMap, GpuPlanner, "skir") for _ in range(n)]
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    rs = np.random.RandomState(1)
    for i in range(n):
        s["pos"][i] = pls[i][4]
        g["pos"][i] = (1.5 + rs.randint(0, 3), 1.5 + rs.randint(0, 3), 5.5)
    s["control"] = g["control"] = mp.ACC
    res = np.zeros(n, dtype=_lib.RESULT_DTYPE)
    handles = (C.c_void_p * n)(*[p[2].pl._h for p in pls])
    t0 = time.perf_counter()
    _lib.check(_lib.lib().mplb_lpa_plan_batch(handles, n, _lib.ptr(s), _lib.ptr(g), _lib.ptr(res)))
    out["gpu_batch64_first_plan_ms_incl_allocation"] = (time.perf_counter() - t0) * 1e3  # 64 x ~70 MB of cudaMalloc + memset inside
    out["gpu_batch64_pops_total"] = int(res["pops"].sum())
    # the same 64 robots replan towards new goals on their kept state spaces: no allocation, one launch
    for i in range(n):
        g["pos"][i] = (5.5 - rs.randint(0, 3), 1.5 + rs.randint(0, 3), 0.5 + rs.randint(0, 3))
    t0 = time.perf_counter()
    _lib.check(_lib.lib().mplb_lpa_plan_batch(handles, n, _lib.ptr(s), _lib.ptr(g), _lib.ptr(res)))
    out["gpu_batch64_second_plan_ms"] = (time.perf_counter() - t0) * 1e3
    out["gpu_batch64_second_pops_total"] = int(res["pops"].sum())
    out["gpu_batch64_second_ok"] = int((res["status"] == 0).sum())
    return out


if __name__ == "__main__":
    if "--lpa-only" in sys.argv:
        print(json.dumps(dict(lpastar=lpa())))
    else:
        print(json.dumps(dict(trajsolver=trajsolve(), lpastar=lpa())))

"""CPU checks of the oracle's cost-shaping branch: search region + potential map.

The reference publishes no numbers for this branch, so these are structural properties plus a regression pin of the
oracle's own answer for MPL/test/test_distance_map_planner_2d.cpp's flow (tests/test_oracle_vs_reference.py checks the
same flow against the reference's own sources where they can be compiled).
"""
import numpy as np

import oracle
from helpers import load_config


def _setup():
    m, dim, params, U, start, goal = load_config("corridor")
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()

    def planner():
        op = oracle.OraclePlanner(dim)
        op.set_map(om)
        for k, v in params.items():
            op.set_param(k, v)
        op.set_controls(U)
        return op
    s, g = oracle.make_waypoints(1), oracle.make_waypoints(1)
    s["pos"][0, :2], g["pos"][0, :2] = start, goal
    s["control"] = g["control"] = 3
    return m, om, planner, s, g, U, params


def test_distance_map_flow_oracle():
    m, om, planner, s, g, U, params = _setup()
    ncell = int(np.prod(m.dim))
    op = planner()
    r1 = op.plan(s, g)
    assert (r1["status"], r1["n_seg"], r1["n_closed"], r1["cost"]) == (0, 35, 615, 351.5)  # MPL/README.md:200-202
    acts, st = op.actions(35), op.seg_states(35)
    # waypoints of the trajectory (trajectory.h:277-289): segment starts plus the final end point
    path = np.zeros((36, 3))
    path[:35, :2] = st[:, :2]
    dt = params["dt"]
    last = st[-1]
    path[35, :2] = last[:2] + last[3:5] * dt + 0.5 * U[acts[-1]] * dt * dt
    before = om.get_data(ncell).copy()

    op2 = planner()
    op2.set_param("epsilon", 1.0)
    op2.set_vec("search_radius", [0.5, 0.5, 0.0])
    op2.set_search_region(path, dense=False)
    region = op2.get_search_region(ncell)
    assert region.size == ncell and 0 < region.sum() < ncell
    # every path point lies in the region
    for q in path:
        pn, idx = om.float_to_int(q)
        assert idx >= 0 and region[idx] == 1
    op2.set_vec("potential_radius", [1.0, 1.0, 0.0])
    op2.set_param("potential_weight", 0.5)
    op2.set_param("gradient_weight", 0.0)
    op2.update_potential_map(np.array([s["pos"][0, 0], s["pos"][0, 1], 0.0]))
    after = om.get_data(ncell)
    assert np.all(after[before > 0] == 100)          # every source cell is H_MAX
    assert np.all(after >= before)                   # stamping only raises values
    assert ((after > 0) & (after < 100)).sum() > 1000  # and there is a graded halo
    r2 = op2.plan(s, g)
    assert r2["status"] == 0 and r2["n_seg"] == 36 and r2["pops"] == 2732 and abs(r2["cost"] - 647.1) < 1e-9
    # shaping never makes a plan cheaper than the plain optimum
    assert r2["cost"] > r1["cost"]
    # clear_shaping restores the plain behaviour on the (rewritten) map: potential cells < 100 count as free
    op2.clear_shaping()
    r3 = op2.plan(s, g)
    assert r3["status"] == 0 and r3["cost"] == r1["cost"]

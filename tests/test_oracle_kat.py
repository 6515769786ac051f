"""Pin the oracle on the only numeric known-answer the reference publishes for this path.

MPL/README.md:195-202 (`test_planner_2d data/corridor.yaml`): expanded states 615, total time T = 35,
J(VEL) = 36.75, J(ACC) = 1.5  (=> cost g = 35*w + 1.5 = 351.5 with w = 10, env_base.h:370).
Set-up follows MPL/test/test_planner_2d.cpp:21-62.
"""
import numpy as np

import oracle
from mpl_ros_b200 import maps
from helpers import fill_waypoints, load_config, traj_J


def make_oracle(name):
    m, dim, params, U, start, goal = load_config(name)
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    p = oracle.OraclePlanner(dim)
    p.set_map(om)
    for k, v in params.items():
        p.set_param(k, v)
    p.set_controls(U)
    s = fill_waypoints(oracle.make_waypoints(1), start, maps.ACC)
    g = fill_waypoints(oracle.make_waypoints(1), goal, maps.ACC)
    return p, om, U, s, g, m


def test_corridor_kat():
    p, om, U, s, g, m = make_oracle("corridor")
    r = p.plan(s, g)
    assert r["status"] == 0
    assert r["n_closed"] == 615          # "MPL Planner expanded states: 615" (getCloseSet().size())
    assert r["pops"] == 615
    assert r["n_seg"] * 1.0 == 35.0      # "Total time T: 35.000000"
    assert r["cost"] == 351.5
    acts = p.actions(r["n_seg"])
    st = p.seg_states(r["n_seg"])
    jv, ja = traj_J(U, acts, st, 1.0, 2, 2)
    assert jv == 36.75 and ja == 1.5     # "J(VEL) = 36.750000, J(ACC) = 1.500000"
    # survey-time secondary values: hash-map nodes, verdict mix
    assert r["n_nodes"] == 1898 and r["n_valid"] == 2539 and r["n_prims"] == 615 * 9


def test_secondary_regressions():
    """Survey-time restatement values (not reference-certified), SURVEY.md §8(c).3."""
    p, om, U, s, g, m = make_oracle("simple")
    r = p.plan(s, g)
    assert (r["status"], r["n_closed"], r["n_seg"], r["cost"], r["n_prims"]) == (0, 143, 9, 96.0, 1287)
    p, om, U, s, g, m = make_oracle("skir")
    r = p.plan(s, g)
    assert (r["status"], r["n_closed"], r["n_seg"], r["cost"], r["n_prims"]) == (0, 348, 4, 47.0, 9396)


def test_trajectory_is_consistent():
    """Segments chain: evaluating segment i at dt lands on segment i+1's stored lattice node (same key)."""
    p, om, U, s, g, m = make_oracle("skir")
    r = p.plan(s, g)
    acts, st = p.actions(r["n_seg"]), p.seg_states(r["n_seg"])
    for i in range(r["n_seg"] - 1):
        u = U[acts[i]]
        pos = st[i][0:3] + st[i][3:6] * 1.0 + u / 2 * 1.0 * 1.0
        vel = st[i][3:6] + u * 1.0
        assert np.all(np.round(pos / 0.01) == np.round(st[i + 1][0:3] / 0.01))
        assert np.all(np.round(vel / 0.1) == np.round(st[i + 1][3:6] / 0.1))


def test_sample_time_accumulation():
    """SURVEY.md §7 hard part 2(a): `for (t = 0; t < T; t += T/n)` yields n or n+1 samples."""
    def count(n, T=1.0):
        t, c, d = 0.0, 0, T / n
        while t < T:
            c += 1
            t += d
        return c
    assert count(10) == 11 and count(5) == 5 and count(20) == 20

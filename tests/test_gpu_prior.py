"""Prior-trajectory heuristic on the CUDA path (env_base.h:46-53,249-256, env_map.h:187-225; flow of
MPL/test/test_planner_2d_with_prior_traj.cpp:29-105): a VEL-control plan becomes the prior of a second planner; the
heuristic then measures the distance to the prior's waypoint at the state's own time plus the prior's remaining cost,
and the prior's end point replaces the requested goal.  Compared exactly with the oracle, which
tests/test_oracle_vs_reference.py::test_prior_trajectory_heuristic holds to the reference's own sources on this flow."""
import numpy as np
import pytest

import oracle  # noqa: F401
import mpl_ros_b200 as mp
from mpl_ros_b200 import maps
from helpers import load_config
from helpers_gpu import assert_results_equal, make_pair, waypoint_pair

pytestmark = pytest.mark.gpu


def _full_compare(pl, op, ro, ncol):
    """ncol = 3 * control order: the stored coord keeps the derivatives below the control order"""
    assert_results_equal(pl.result(), ro)
    assert np.array_equal(pl.getActions(), op.actions(ro["n_seg"]))
    assert np.array_equal(pl.getSegStates()[:, :ncol], op.seg_states(ro["n_seg"])[:, :ncol])
    nodes = pl.getNodes()
    onodes = op.nodes(ro["n_nodes"])
    kg = {tuple(k[:k[15]]): (g, h) for k, g, h in zip(nodes["key"], nodes["g"], nodes["h"])}
    ko = {tuple(k[:k[15]]): (g, h) for k, g, h in zip(onodes["key"], onodes["g"], onodes["h"])}
    assert kg == ko  # every node with the same g and the same (prior-shaped) h
    keys = np.array([nodes["key"][i] for i in pl.getPopLog()])
    assert np.array_equal(keys, op.pop_keys(ro["pops"]))


def test_prior_trajectory_flow_2d():
    m, dim, params, _, start, goal = load_config("corridor")
    U1 = maps.make_U(1.0, 1, 2)
    pl1, op1 = make_pair(m, dim, dict(v_max=1.0, a_max=1.0, dt=1.0), U1)
    sg, so = waypoint_pair(start, mp.VEL)
    gg, go = waypoint_pair(goal, mp.VEL)
    assert pl1.plan(sg, gg)
    ro1 = op1.plan(so, go)
    _full_compare(pl1, op1, ro1, 3)
    prior = pl1.getTraj()
    U2 = maps.make_U(1.0, 1, 2) * 0.5
    for ctl, prm in ((mp.JRK, dict(epsilon=1.0, v_max=1.0, a_max=1.0, dt=1.0, w=10.0, tol_pos=0.5, max_num=20000)),
                     (mp.ACC, dict(epsilon=1.0, v_max=1.0, a_max=1.0, dt=1.0, w=10.0, tol_pos=0.5))):
        pl2, op2 = make_pair(m, dim, prm, U2)
        pl2.setPriorTrajectory(prior)
        op2.set_prior_trajectory(op1)
        sg, so = waypoint_pair(start, ctl)
        gg, go = waypoint_pair(goal, ctl)
        pl2.plan(sg, gg)
        ro2 = op2.plan(so, go)
        _full_compare(pl2, op2, ro2, 9 if ctl == mp.JRK else 6)
        # the requested goal is ignored while a prior is installed (env_base.h:295-298): a different goal, same answer
        other = np.array(goal) + np.array([-3.0, 0.5])
        g2, _ = waypoint_pair(other, ctl)
        pl2.plan(sg, g2)
        assert_results_equal(pl2.result(), ro2)
        # clearing the prior gives the plain plan again
        pl2.setPriorTrajectory(None)
        op3 = make_pair(m, dim, prm, U2)[1]
        pl2.plan(sg, gg)
        ro3 = op3.plan(so, go)
        assert_results_equal(pl2.result(), ro3)
        assert ro3["pop_hash"] != ro2["pop_hash"]


def test_prior_trajectory_batch_3d():
    """A batch sharing one prior (3D, ACC prior for a JRK search), starts spread around the prior's start."""
    m, dim, params, U, start, goal = load_config("skir")
    pl1, op1 = make_pair(m, dim, params, U)
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    assert pl1.plan(sg, gg)
    op1.plan(so, go)
    prm = dict(v_max=2.0, a_max=1.0, j_max=2.0, dt=1.0, tol_pos=0.5, max_num=1500)
    pl2, op2 = make_pair(m, dim, prm, U)
    pl2.setPriorTrajectory(pl1.getTraj())
    op2.set_prior_trajectory(op1)
    S = np.array(start) + np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.2, 0], [0.1, 0.1, 0.1], [-0.1, 0, 0.1]])
    G = np.tile(goal, (len(S), 1))
    sg, so = waypoint_pair(S, mp.JRK)
    gg, go = waypoint_pair(G, mp.JRK)
    res, acts, _ = pl2.plan_batch(sg, gg, max_seg=64)
    ro, ao = op2.plan_batch(so, go, nthreads=4, max_seg=64)
    for f in ("status", "n_seg", "pops", "n_nodes", "n_open", "n_closed", "pop_hash", "closed_hash", "n_valid", "n_samples"):
        assert np.array_equal(res[f], ro[f]), f
    assert np.array_equal(res["cost"][np.isfinite(ro["cost"])], ro["cost"][np.isfinite(ro["cost"])])
    assert np.array_equal(acts, ao)

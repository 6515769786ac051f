"""GPU parity tests of SURVEY section 8(f).1: search region + potential map + iterativePlan (env_map.h:104-128,
map_planner.cpp:20-114,286-434) against the oracle, bit-exact.

The reference publishes no numbers for this branch (test_distance_map_planner_2d.cpp only draws a picture), so the
oracle side is pinned by the reference's own sources run on the CPU (tests/test_oracle_vs_reference.py); the plain-map
first plan of the same flow is the README known answer.
"""
import numpy as np
import pytest

import mpl_ros_b200 as mp
from mpl_ros_b200 import maps
from helpers import load_config
from helpers_gpu import assert_results_equal, make_pair, waypoint_pair

pytestmark = pytest.mark.gpu


def _ncell(m):
    return int(np.prod(np.asarray(m.dim, dtype=np.int64)))


def _path_of(pl):
    return [np.array(w.pos, dtype=np.float64) for w in pl.getTraj().getWaypoints()]


def _oracle_path(path):
    out = np.zeros((len(path), 3))
    for i, q in enumerate(path):
        out[i, :len(q)] = q
    return out


def _compare_plan(pl, op, sg, gg, so, go, ctx):
    ok = pl.plan(sg, gg)
    ro = op.plan(so, go)
    rg = pl.result()
    assert_results_equal(rg, ro, ctx)
    assert ok == (ro["status"] in (0, 5))
    if ro["status"] == 0:
        assert np.array_equal(pl.getActions(), op.actions(ro["n_seg"]))
        assert np.array_equal(pl.getSegStates()[:, :6], op.seg_states(ro["n_seg"])[:, :6])
        gn = pl.getNodes()
        assert np.array_equal(gn["key"][pl.getPopLog()], op.pop_keys(ro["pops"]))
    return rg, ro


@pytest.mark.parametrize("grad_w", [0.0, 0.3])
def test_distance_map_planner_2d_flow(grad_w):
    """MPL/test/test_distance_map_planner_2d.cpp:46-93, step for step, on corridor.yaml."""
    m, dim, params, U, start, goal = load_config("corridor")
    pl, op = make_pair(m, dim, params, U)
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    rg, _ = _compare_plan(pl, op, sg, gg, so, go, "plain")
    assert rg["n_closed"] == 615 and rg["n_seg"] == 35  # MPL/README.md:200-202
    path = _path_of(pl)
    mu, om = pl._keep

    # "planner.reset(new OccMapPlanner)" on the same map_util
    pl2, op2 = make_pair(m, dim, dict(params, epsilon=1.0), U)
    pl2.setMapUtil(mu)
    op2.set_map(om)
    pl2.setSearchRadius([0.5, 0.5])
    op2.set_vec("search_radius", [0.5, 0.5, 0.0])
    pl2.setSearchRegion(path)
    op2.set_search_region(_oracle_path(path), dense=False)
    region = pl2.getSearchRegionMask()
    assert np.array_equal(region, op2.get_search_region(_ncell(m)))
    assert 0 < region.sum() < region.size
    assert len(pl2.getSearchRegion()) == int(region.sum())

    pl2.setPotentialRadius([1.0, 1.0])
    op2.set_vec("potential_radius", [1.0, 1.0, 0.0])
    pl2.setPotentialWeight(0.5)
    op2.set_param("potential_weight", 0.5)
    pl2.setGradientWeight(grad_w)
    op2.set_param("gradient_weight", grad_w)
    pl2.updatePotentialMap(start)
    op2.update_potential_map(np.array([start[0], start[1], 0.0]))
    dmap = mu.getMap()
    assert np.array_equal(dmap, om.get_data(_ncell(m)))  # the shared map itself was rewritten (map_planner.cpp:387)
    assert ((dmap > 0) & (dmap < 100)).sum() > 1000

    rg2, ro2 = _compare_plan(pl2, op2, sg, gg, so, go, "shaped")
    assert rg2["status"] == 0 and rg2["cost"] > rg["cost"]
    if grad_w == 0.0:  # regression pin of the oracle's own answer for this flow
        assert rg2["n_seg"] == 36 and rg2["pops"] == 2732

    # the first planner has no potential map of its own but now sees the rewritten map (0 < v < 100 stays free)
    _compare_plan(pl, op, sg, gg, so, go, "plain planner on the rewritten map")


def test_search_region_only_3d():
    """em:104-106 without a potential map: the occupancy test stays on the bit-bricks."""
    m, dim, params, U, start, goal = load_config("skir")
    pl, op = make_pair(m, dim, params, U)
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    rg, _ = _compare_plan(pl, op, sg, gg, so, go, "plain")
    assert rg["status"] == 0
    path = _path_of(pl)
    for radius, dense in (([0.5, 0.5, 0.5], False), ([0.2, 0.2, 0.1], False), ([1.0, 1.0, 0.3], True)):
        pl.setSearchRadius(radius)
        op.set_vec("search_radius", radius)
        pl.setSearchRegion(path, dense)
        op.set_search_region(_oracle_path(path), dense=dense)
        assert np.array_equal(pl.getSearchRegionMask(), op.get_search_region(_ncell(m)))
        _compare_plan(pl, op, sg, gg, so, go, ("region", radius, dense))
    # clearing the region restores the plain answer
    pl.setSearchRegionMask(None)
    op.clear_shaping()
    rg3, _ = _compare_plan(pl, op, sg, gg, so, go, "cleared")
    assert rg3["cost"] == rg["cost"] and rg3["pops"] == rg["pops"]


def test_potential_map_3d_local_range():
    """3D createMask (radius + half-height) stamped only inside pos +- range (map_planner.cpp:330-347)."""
    m, dim, params, U, start, goal = load_config("skir")
    pl, op = make_pair(m, dim, params, U)
    mu, om = pl._keep
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    pl.setPotentialRadius([0.4, 0.4, 0.2])
    op.set_vec("potential_radius", [0.4, 0.4, 0.2])
    pl.setPotentialMapRange([3.0, 2.5, 1.0])
    op.set_vec("potential_map_range", [3.0, 2.5, 1.0])
    pl.setPotentialWeight(0.2)
    op.set_param("potential_weight", 0.2)
    pl.setGradientWeight(0.1)
    op.set_param("gradient_weight", 0.1)
    pl.updatePotentialMap(start)
    op.update_potential_map(np.asarray(start, dtype=np.float64))
    dmap = mu.getMap()
    assert np.array_equal(dmap, om.get_data(_ncell(m)))
    assert ((dmap > 0) & (dmap < 100)).sum() > 100
    _compare_plan(pl, op, sg, gg, so, go, "potential 3d")
    # a second stamping pass works on the already rewritten map (cells > 0 are sources, as in the reference)
    pl.updatePotentialMap(goal)
    op.update_potential_map(np.asarray(goal, dtype=np.float64))
    assert np.array_equal(mu.getMap(), om.get_data(_ncell(m)))
    _compare_plan(pl, op, sg, gg, so, go, "potential 3d, second pass")


def test_iterative_plan():
    """MapPlanner::iterativePlan (map_planner.cpp:394-434): tunnel around the previous trajectory until the cost repeats."""
    m, dim, params, U, start, goal = load_config("corridor")
    pl, op = make_pair(m, dim, params, U)
    sg, so = waypoint_pair(start, mp.ACC)
    gg, go = waypoint_pair(goal, mp.ACC)
    _compare_plan(pl, op, sg, gg, so, go, "raw")
    raw = pl.getTraj()
    pl.setSearchRadius([0.3, 0.3])
    op.set_vec("search_radius", [0.3, 0.3, 0.0])
    pl.setPotentialRadius([0.6, 0.6])
    op.set_vec("potential_radius", [0.6, 0.6, 0.0])
    pl.setPotentialWeight(0.3)
    op.set_param("potential_weight", 0.3)
    pl.updatePotentialMap(start)
    op.update_potential_map(np.array([start[0], start[1], 0.0]))
    assert pl.iterativePlan(sg, gg, raw, 4)
    # the same loop on the oracle, fed with the oracle's own trajectories
    from mpl_ros_b200.planner import Primitive, Trajectory
    traj, prev, costs = raw, 0.0, []
    for _ in range(4):
        op.set_search_region(_oracle_path([w.pos for w in traj.getWaypoints()]), dense=False)
        ro = op.plan(so, go)
        assert ro["status"] == 0
        acts, st = op.actions(ro["n_seg"]), op.seg_states(ro["n_seg"])
        traj = Trajectory([Primitive(dim, mp.ACC, st[i], U[acts[i]], params["dt"]) for i in range(len(acts))])
        costs.append(float(ro["cost"]))
        if prev == ro["cost"]:
            break
        prev = ro["cost"]
    assert pl.getTrajCost() == costs[-1]
    assert_results_equal(pl.result(), ro, "iterative")
    assert np.array_equal(pl.getActions(), acts)


def test_shaped_batch_parity():
    """plan_batch with a search region and a potential map installed: every plan equals the oracle's."""
    m = maps.load_fixture("levine")
    U = maps.make_U(1.0, 1, 3)
    params = dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5)
    pl, op = make_pair(m, 3, params, U)
    mu, om = pl._keep
    n = 64
    S, G = maps.sample_queries(m, n, seed=3)
    sg, so = waypoint_pair(S, mp.ACC)
    gg, go = waypoint_pair(G, mp.ACC)
    pl.setPotentialRadius([0.5, 0.5, 0.3])
    op.set_vec("potential_radius", [0.5, 0.5, 0.3])
    pl.setPotentialWeight(0.1)
    op.set_param("potential_weight", 0.1)
    pl.updatePotentialMap(S[0])
    op.update_potential_map(np.asarray(S[0], dtype=np.float64))
    assert np.array_equal(mu.getMap(), om.get_data(_ncell(m)))
    # region: everything except a slab, given as an explicit mask (env_base::set_search_region)
    nd = np.asarray(m.dim)
    mask = np.ones(tuple(nd[::-1]), dtype=np.uint8)
    mask[:, :, nd[0] // 2] = 0
    mask[:, : nd[1] // 3, nd[0] // 2] = 1
    pl.setSearchRegionMask(mask)
    # the oracle takes regions as paths only; a dense "path" over the in-region cell centres with radius 0 is the same mask
    idx = np.flatnonzero(mask.ravel())
    cells = np.stack(np.unravel_index(idx, tuple(nd[::-1])), axis=1)[:, ::-1]
    centres = (cells + 0.5) * float(np.float32(m.res)) + np.asarray(m.origin)
    op.set_vec("search_radius", [0.0, 0.0, 0.0])
    op.set_search_region(np.ascontiguousarray(centres), dense=True)
    assert np.array_equal(pl.getSearchRegionMask(), op.get_search_region(_ncell(m)))
    rg, ag, _ = pl.plan_batch(sg, gg, max_seg=64, want_states=True)
    ro, ao = op.plan_batch(so, go, nthreads=8, max_seg=64)
    for i in range(n):
        assert_results_equal(rg[i], ro[i], ("shaped batch", i))
    assert np.array_equal(ag, ao)
    assert (ro["status"] == 0).sum() >= 8 and (ro["status"] == 3).sum() >= 8  # both outcomes are exercised


def test_shaping_errors():
    m, dim, params, U, start, goal = load_config("corridor")
    pl, _ = make_pair(m, dim, params, U)
    with pytest.raises(mp.MplbError):
        pl.setSearchRegionMask(np.ones(7, dtype=np.uint8))
    with pytest.raises(mp.MplbError):
        pl.setPotentialMap(np.zeros(5, dtype=np.int8))
    pl.setSearchRegionMask(np.ones(_ncell(m), dtype=np.uint8))
    sg, _ = waypoint_pair(start, mp.ACC)
    with pytest.raises(mp.MplbError):  # the get_succ trace models the plain map only
        pl.expand(sg)

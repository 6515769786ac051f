"""BASELINE configs[4] at FULL size: synthetic 1024^3 box map (1 GiB int8, 128 MiB of bit-bricks — the one
configuration whose voxel reads leave L2), |U| = 125 jerk controls, max_num = 50 000.

* test_c5_slice_vs_golden: the first 512 queries of the 65 536-query list through the C ABI against the committed fixture
  tests/golden/c5_results.npz (tools/make_golden_c5.py: oracle port on 512 queries, the reference's own sources on a
  prefix, asserted equal to each other when the fixture was recorded).  Exact comparison of every counter and hash; the
  statuses 7 (KEY_RANGE) and 8 (NOMEM), which do not exist in the reference, must not occur.
* test_c5_live_oracle_small_budget: a few of the same queries with max_num = 1500, planned live by the oracle on the box.
"""
import os

import numpy as np
import pytest

import mpl_ros_b200 as mp
from mpl_ros_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_results.npz")
FIELDS = ("status", "n_seg", "cost", "pops", "n_nodes", "n_open", "n_closed", "n_prims", "n_samples", "n_valid", "pop_hash",
          "closed_hash")


@pytest.fixture(scope="module")
def c5():
    m = W.c5_map()
    mu = mp.VoxelMapUtil()
    mu.setMap(m.origin, m.dim, m.data, m.res)
    pl = mp.VoxelMapPlanner(False)
    pl.setMapUtil(mu)
    p = W.C5["params"]
    pl.setVmax(p["v_max"]); pl.setAmax(p["a_max"]); pl.setDt(p["dt"]); pl.setU(W.controls(W.C5)); pl.setTol(p["tol_pos"])
    pl.setMemFraction(0.85)
    pl._keep = mu
    return m, pl


@pytest.mark.gpu
def test_c5_slice_vs_golden(c5):
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/c5_results.npz not recorded")
    m, pl = c5
    z = np.load(GOLD)
    n = len(z["oracle/status"])
    S, G = W.c5_queries(m, n)
    assert np.array_equal(S, z["starts"]) and np.array_equal(G, z["goals"]), "query list differs from the recorded one"
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    W.fill(s, g, S, G, W.C5["control"])
    pl.setMaxNum(W.C5["params"]["max_num"])
    res, acts, _ = pl.plan_batch(s, g, max_seg=64)
    assert not np.isin(res["status"], (7, 8)).any(), np.unique(res["status"], return_counts=True)
    for f in FIELDS:
        a, b = res[f], z["oracle/" + f]
        assert np.array_equal(a, b) or (f == "cost" and np.array_equal(np.isinf(a), np.isinf(b))
                                        and np.array_equal(a[np.isfinite(a)], b[np.isfinite(b)])), ("oracle", f)
    if "reference/pops" in z.files:  # the reference's own sources (status -1: its bool does not say why a plan failed)
        k = len(z["reference/pops"])
        for f in ("pops", "n_nodes", "n_open", "n_closed", "n_prims", "n_valid", "pop_hash", "closed_hash"):
            assert np.array_equal(res[f][:k], z["reference/" + f]), ("reference", f)
    st = pl.last_batch_stats()
    print("C5 slice: %d plans, kernel %.1f ms, tiers %d, %.3g prim/s" % (n, st["kernel_ms"], st["tiers"],
                                                                        res["n_prims"].sum() / (st["kernel_ms"] * 1e-3)))


@pytest.mark.gpu
def test_c5_live_oracle_small_budget(c5):
    import oracle
    m, pl = c5
    n = 6
    S, G = W.c5_queries(m, n)
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    W.fill(s, g, S, G, W.C5["control"])
    pl.setMaxNum(1500)
    res, acts, _ = pl.plan_batch(s, g, max_seg=64)
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    op = oracle.OraclePlanner(3)
    op.set_map(om)
    for k, v in dict(W.C5["params"], max_num=1500).items():
        op.set_param(k, v)
    op.set_controls(W.controls(W.C5))
    so, go = oracle.make_waypoints(n), oracle.make_waypoints(n)
    W.fill(so, go, S, G, W.C5["control"])
    ro, ao = op.plan_batch(so, go, nthreads=n, max_seg=64)
    for f in FIELDS:
        a, b = res[f], ro[f]
        assert np.array_equal(a, b) or (f == "cost" and np.array_equal(a[np.isfinite(a)], b[np.isfinite(b)])), f
    assert np.array_equal(acts, ao)

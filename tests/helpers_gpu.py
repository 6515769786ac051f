"""Build the same problem on both sides: the product planner (GPU, through the C ABI) and the oracle."""
import numpy as np

import oracle
import mpl_ros_b200 as mp
from mpl_ros_b200 import maps

RESULT_FIELDS = ("status", "n_seg", "cost", "pops", "n_nodes", "n_open", "n_closed", "n_prims", "n_samples", "n_valid",
                 "pop_hash", "closed_hash")
PARAM_SETTERS = dict(v_max="setVmax", a_max="setAmax", j_max="setJmax", dt="setDt", w="setW", epsilon="setEpsilon",
                     max_num="setMaxNum", yaw_max="setYawmax", wyaw="setWyaw")


def make_pair(m, dim, params, U):
    """returns (gpu planner, oracle planner) configured identically (setter order as map_planner_node.cpp:173-182)."""
    mu = mp.MapUtil(dim)
    mu.setMap(m.origin, m.dim, m.data, m.res)
    mu.freeUnknown()
    pl = mp.MapPlanner(dim, False)
    pl.setMapUtil(mu)
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    op = oracle.OraclePlanner(dim)
    op.set_map(om)
    tol = {}
    for k, v in params.items():
        if k.startswith("tol_"):
            tol[k] = v
            continue
        getattr(pl, PARAM_SETTERS[k])(v)
        op.set_param(k, v)
    if tol:
        pl.setTol(tol.get("tol_pos", 0.5), tol.get("tol_vel", -1), tol.get("tol_acc", -1))
        for k in ("tol_pos", "tol_vel", "tol_acc"):
            op.set_param(k, tol.get(k, 0.5 if k == "tol_pos" else -1))
    pl.setU(U)
    op.set_controls(U)
    pl._keep = (mu, om)
    return pl, op


def waypoint_pair(pos, control, vel=None, acc=None, yaw=None):
    a, b = mp.waypoints_array(len(np.atleast_2d(pos))), oracle.make_waypoints(len(np.atleast_2d(pos)))
    for w in (a, b):
        p = np.atleast_2d(np.asarray(pos, dtype=np.float64))
        w["pos"][:, :p.shape[1]] = p
        if vel is not None:
            v = np.atleast_2d(np.asarray(vel, dtype=np.float64))
            w["vel"][:, :v.shape[1]] = v
        if acc is not None:
            v = np.atleast_2d(np.asarray(acc, dtype=np.float64))
            w["acc"][:, :v.shape[1]] = v
        if yaw is not None:
            w["yaw"] = yaw
        w["control"] = control
    return a, b


def assert_results_equal(rg, ro, ctx=""):
    for f in RESULT_FIELDS:
        a, b = rg[f], ro[f]
        if f == "cost":
            assert (a == b) or (np.isinf(a) and np.isinf(b)), (ctx, f, a, b)
        else:
            assert a == b, (ctx, f, a, b)

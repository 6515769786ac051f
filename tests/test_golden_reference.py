"""tests/golden/reference_results.npz (outputs of the reference's own planner sources, recorded by
tools/make_golden_reference.py) against the oracle (CPU) and against the CUDA path (GPU).  Unlike
tests/test_oracle_vs_reference.py and tests/test_gpu_vs_reference.py these need neither /root/reference nor the harness
binary: the fixture is committed."""
import os

import numpy as np
import pytest

import oracle
import golden_cases as gc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_results.npz")


def _gold(name):
    z = np.load(GOLD)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def _make_oracle(case):
    m = case["map"]
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    op = oracle.OraclePlanner(case["dim"])
    op.set_map(om)
    for k, v in case["params"].items():
        op.set_param(k, v)
    op.set_controls(case["U"])
    op._keep = om
    return op


def _plan_oracle(op, s, g, control):
    ws, wg = oracle.make_waypoints(1), oracle.make_waypoints(1)
    ws["pos"][0, :len(s)], wg["pos"][0, :len(g)] = s, g
    ws["control"] = wg["control"] = control
    return op.plan(ws, wg)


CASES = gc.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_recorded_reference(name):
    gc.compare(gc.pack(gc.run_case(CASES[name], _make_oracle, _plan_oracle)), _gold(name), name)


def _make_gpu(case):
    import mpl_ros_b200 as mp
    m, dim, prm = case["map"], case["dim"], case["params"]
    mu = mp.MapUtil(dim)
    mu.setMap(m.origin, m.dim, m.data, m.res)
    mu.freeUnknown()
    pl = mp.MapPlanner(dim, False)
    pl.setMapUtil(mu)
    setters = dict(v_max="setVmax", a_max="setAmax", j_max="setJmax", dt="setDt", w="setW", epsilon="setEpsilon",
                   max_num="setMaxNum")
    for k, v in prm.items():
        if k in setters:
            getattr(pl, setters[k])(v)
    pl.setTol(prm.get("tol_pos", 0.5), prm.get("tol_vel", -1), prm.get("tol_acc", -1))
    pl.setU(case["U"])
    pl._keep = mu
    return pl


def _plan_gpu(pl, s, g, control):
    import mpl_ros_b200 as mp
    ws, wg = mp.waypoints_array(1), mp.waypoints_array(1)
    ws["pos"][0, :len(s)], wg["pos"][0, :len(g)] = s, g
    ws["control"] = wg["control"] = control
    pl.plan(ws, wg)
    return pl.result().copy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_recorded_reference(name):
    gc.compare(gc.pack(gc.run_case(CASES[name], _make_gpu, _plan_gpu)), _gold(name), name)

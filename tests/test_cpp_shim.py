"""The C++ host layer (include/mpl_b200/map_planner.hpp): compiles on the CPU box; on the GPU box the reference-shaped
C++ test (tests/cpp/test_planner_2d.cpp, mirroring MPL/test/test_planner_2d.cpp) must print the README's known answers."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="test_planner_2d"):
    from mpl_ros_b200.build import build_lib
    so = build_lib()
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, so,
                           "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def _write_corridor(tmp_path):
    from mpl_ros_b200 import maps
    m = maps.load_fixture("corridor")
    p = str(tmp_path / "corridor.bin")
    with open(p, "wb") as f:
        f.write(struct.pack("<2i", *m.dim.tolist()))
        f.write(struct.pack("<2d", *m.origin.tolist()))
        f.write(struct.pack("<d", m.res))
        f.write(struct.pack("<2d", *m.extra["start"].tolist()))
        f.write(struct.pack("<2d", *m.extra["goal"].tolist()))
        f.write(m.data.tobytes())
    return p


@pytest.mark.parametrize("name", ["test_planner_2d", "test_distance_map_planner_2d", "test_planner_2d_with_yaw", "test_replanner"])
def test_cpp_shim_compiles_and_links(tmp_path, name):
    assert os.path.exists(_build(tmp_path, name))


@pytest.mark.gpu
def test_cpp_distance_map_planner_2d(tmp_path):
    """tests/cpp/test_distance_map_planner_2d.cpp (the reference's test_distance_map_planner_2d.cpp flow) against the
    oracle's answer for the same flow (tests/test_oracle_shaping.py pins it on the CPU)."""
    exe = _build(tmp_path, "test_distance_map_planner_2d")
    out = subprocess.check_output([exe, _write_corridor(tmp_path)]).decode()
    assert "MPL Planner expanded states: 615" in out, out
    assert "distance: cost 647.1000000000 pops 2732 segs 36" in out, out
    assert "iterative: ok 1" in out, out


@pytest.mark.gpu
def test_cpp_planner_2d_known_answer(tmp_path):
    from mpl_ros_b200 import maps
    exe = _build(tmp_path)
    m = maps.load_fixture("corridor")
    p = str(tmp_path / "corridor.bin")
    with open(p, "wb") as f:
        f.write(struct.pack("<2i", *m.dim.tolist()))
        f.write(struct.pack("<2d", *m.origin.tolist()))
        f.write(struct.pack("<d", m.res))
        f.write(struct.pack("<2d", *m.extra["start"].tolist()))
        f.write(struct.pack("<2d", *m.extra["goal"].tolist()))
        f.write(m.data.tobytes())
    out = subprocess.check_output([exe, p]).decode()
    assert "MPL Planner expanded states: 615" in out, out        # MPL/README.md:200
    assert "Total time T: 35.000000" in out, out                 # MPL/README.md:201
    assert "J(VEL) = 36.750000, J(ACC) = 1.500000" in out, out   # MPL/README.md:202
    assert "cost: 351.500000" in out and "expanded: 615" in out and "waypoints: 36" in out, out
    # getExpandedEdges / getValidPrimitives = the 2539 finite-cost primitives the survey measured; getCloud = occupied cells
    n_occ = int((m.data == 100).sum())
    assert "edges: 2539 valid: 2539 cloud: %d ray:" % n_occ in out, out
    assert "initialized after reset: 0" in out, out


@pytest.mark.gpu
def test_cpp_planner_2d_with_yaw(tmp_path):
    """tests/cpp/test_planner_2d_with_yaw.cpp (the reference's test_planner_2d_with_yaw.cpp flow) against the oracle's
    answer for the same flow (tests/test_oracle_yaw.py pins it on the CPU in both trig definitions)."""
    exe = _build(tmp_path, "test_planner_2d_with_yaw")
    out = subprocess.check_output([exe, _write_corridor(tmp_path)]).decode()
    assert "MPL Planner expanded states: 1342" in out, out
    assert "yaw: cost 352.4275550989 pops 1342 segs 35 first yaw 1.570796" in out, out


@pytest.mark.gpu
def test_cpp_replanner_flow(tmp_path):
    """tests/cpp/test_replanner.cpp (the callbacks of mpl_test_node/src/map_replanner_node.cpp:107-241 without ROS) through
    the C++ header: A* and LPA* planners on one shared MapUtil edited with getMap / setMap.  Compared with the oracle driven
    through the same sequence."""
    import oracle
    from helpers import fill_waypoints, load_config
    exe = _build(tmp_path, "test_replanner")
    out = subprocess.check_output([exe, _write_corridor(tmp_path)]).decode()
    m, dim, params, U, start, goal = load_config("corridor")
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    pa, pl = oracle.OraclePlanner(2), oracle.OraclePlanner(2)
    for p in (pa, pl):
        p.set_map(om)
        for k, v in params.items():
            p.set_param(k, v)
        p.set_controls(U)
    s, g = oracle.make_waypoints(1), oracle.make_waypoints(1)
    fill_waypoints(s, start, 3)
    fill_waypoints(g, goal, 3)
    grid = m.data.reshape(-1).copy()

    def expect(tag):
        ra, rl = pa.plan(s, g), pl.lpa_plan(s, g)
        line = "%s: astar ok %d cost %.10f closed %d | lpastar ok %d cost %.10f expand %d open %d closed %d segs %d" % (
            tag, ra["status"] == 0, ra["cost"] if ra["status"] == 0 else -1.0, ra["n_closed"], rl["status"] == 0,
            rl["cost"] if rl["status"] == 0 else -1.0, rl["pops"], rl["n_open"], rl["n_closed"], rl["n_seg"])
        assert line in out, (line, out)
        assert "%s: linked %d" % (tag, len(pl.lpa_get_linked_nodes())) in out, (tag, out)

    expect("first")
    ws = pl.lpa_best_child_states()
    c = np.round((ws[int(len(ws) * 0.45)][:2] - m.origin) / m.res - 0.5).astype(int)
    new_obs = []
    for nx in range(-2, 3):
        for ny in range(-2, 3):
            pn = (c[0] + nx, c[1] + ny)
            if 0 <= pn[0] < m.dim[0] and 0 <= pn[1] < m.dim[1] and 0 <= grid[pn[0] + m.dim[0] * pn[1]] < 100:
                grid[pn[0] + m.dim[0] * pn[1]] = 100
                new_obs.append(pn)
    om.set_cells(new_obs, 100)
    pl.lpa_update_blocked_nodes(new_obs)
    assert "blocked %d cells" % len(new_obs) in out
    expect("blocked")
    clear = new_obs[:len(new_obs) // 2]
    om.set_cells(clear, 0)
    pl.lpa_update_cleared_nodes(clear)
    expect("cleared")
    pl._lpa_control = 3
    nxt = pl.lpa_waypoint(1)
    pl.lpa_get_sub_state_space(1)
    s = nxt
    expect("subtree")

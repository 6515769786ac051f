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


@pytest.mark.parametrize("name", ["test_planner_2d", "test_distance_map_planner_2d", "test_planner_2d_with_yaw"])
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

"""The C++ host layer (include/mpl_b200/map_planner.hpp): compiles on the CPU box; on the GPU box the reference-shaped
C++ test (tests/cpp/test_planner_2d.cpp, mirroring MPL/test/test_planner_2d.cpp) must print the README's known answers."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from mpl_ros_b200.build import build_lib
    so = build_lib()
    exe = str(tmp_path / "test_planner_2d")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_planner_2d.cpp"), "-o", exe, so,
                           "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def test_cpp_shim_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


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

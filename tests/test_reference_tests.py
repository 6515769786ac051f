"""The reference's own seven CTests (MPL/test/test_planner_2d.cpp, its five planner siblings and test_traj_solver.cpp) compiled UNMODIFIED against the drop-in headers (tests/cpp/build_reference_tests.py)
and run through the C ABI on the GPU.  The sources print timings and closed-set sizes and assert nothing;
what they print is compared with MPL/README.md:199-202 (615 expanded states) and with the oracle's answers for the same
flows (pinned against the reference's sources by tests/test_oracle_vs_reference.py)."""
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpp"))
import build_reference_tests as brt  # noqa: E402


def _corridor_bin(tmp_path):
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


def test_reference_tests_compile_unmodified():
    """Eigen-mode compile of the header against the reference's own callers (here, where /root/reference exists)."""
    if not brt.available():
        pytest.skip("/root/reference is not present on this box")
    exes = brt.build()
    assert all(os.path.exists(p) for p in exes.values())


def _run(name, tmp_path):
    exe = os.path.join(brt.OUT, name)
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_refbin/%s was not built (needs /root/reference at build time)" % name)
    out = subprocess.check_output([exe, _corridor_bin(tmp_path)], cwd=str(tmp_path), stderr=subprocess.STDOUT, timeout=300).decode()
    return out, [int(x) for x in re.findall(r"expanded states: (\d+)", out)], [int(x) for x in re.findall(r"Expand \[(\d+)\] nodes", out)]


# what the runs must print: closed-set sizes ("expanded states", the reference's own printf) and pops per plan ("Expand [n]
# nodes!", the planner's verbose line, graph_search.h:168) — MPL/README.md:200 for test_planner_2d, the oracle's answers for the
# distance-map flow (tests/test_oracle_shaping.py: 2732 expansions), both pinned against the reference's sources
EXPECTED_STATES = {"test_planner_2d": [615], "test_distance_map_planner_2d": [615, 2732]}
EXPECTED_FIRST_POPS = {"test_planner_2d": 615, "test_distance_map_planner_2d": 615, "test_distance_map_planner_2d_iterative": 2732}  # the iterative test keeps its plain planner quiet: the first verbose plan is the shaped one


@pytest.mark.gpu
@pytest.mark.parametrize("name", brt.TESTS)
def test_reference_test_runs_on_gpu(name, tmp_path):
    out, counts, pops = _run(name, tmp_path)
    if name == "test_traj_solver":  # no planner: three splines (min vel / acc / jrk) through four key frames, 3 segments each
        assert "4 points, 0 circles, 9 trajectory segments" in out, out
        return
    assert pops, out  # every test plans at least once with a verbose planner
    assert "[stand-in drawing]" in out  # the run reached its plotting section, i.e. every planner call returned
    if name in EXPECTED_STATES:
        assert counts == EXPECTED_STATES[name], out
    if name in EXPECTED_FIRST_POPS:
        assert pops[0] == EXPECTED_FIRST_POPS[name], out
    print(name, counts, pops)

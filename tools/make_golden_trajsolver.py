"""Records tests/golden/trajsolver.npz from the REFERENCE'S OWN TrajSolver sources (oracle/_ref, built from
/root/reference against oracle/shim): the MPL/test/test_traj_solver.cpp path with all three solvers, and seeded random
waypoint lists (2D / 3D, mixed control flags as map_planner_node.cpp:216-227 produces them, yaw key frames, the three yaw
orders).  Run in the build container:  python tools/make_golden_trajsolver.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from trajsolver_cases import cases  # noqa: E402


def main():
    out = {}
    for name, dim, control, yaw_control, wps, dts in cases():
        co = ref.traj_solve(dim, control, wps, dts, yaw_control)
        oo = oracle.traj_solve(dim, control, wps, dts, yaw_control)
        assert np.array_equal(co, oo), name  # the oracle restatement equals the reference's sources bit for bit
        out[name] = co
    # the setPath / setV flow of MPL/test/test_traj_solver.cpp through the reference's own TrajSolver::setPath
    path = [(0, 0), (1, 0), (2, 1), (5, 1)]
    for cname, c in (("VEL", 1), ("ACC", 3), ("JRK", 7)):
        co, dts = ref.traj_solve_path(2, c, path, 1.0)
        assert np.array_equal(co, out["test_traj_solver_%s" % cname]) and np.array_equal(dts, [1, 1, 3])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "trajsolver.npz"), **out)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()

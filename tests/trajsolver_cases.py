"""Seeded inputs of the TrajSolver parity tests (shared by tools/make_golden_trajsolver.py and the tests)."""
import numpy as np

WAYPOINT_DTYPE = np.dtype([("pos", "f8", 3), ("vel", "f8", 3), ("acc", "f8", 3), ("jrk", "f8", 3),
                           ("yaw", "f8"), ("t", "f8"), ("control", "i4"), ("enable_t", "i4")], align=True)
VEL, ACC, JRK, SNP = 1, 3, 7, 15


def path_case(dim, control, path):
    """TrajSolver::setPath (traj_solver.h:54-70) + allocate_time with v = 1."""
    w = np.zeros(len(path), dtype=WAYPOINT_DTYPE)
    for i, p in enumerate(path):
        w["pos"][i, :dim] = p
        w["control"][i] = VEL
    w["control"][0] = w["control"][-1] = control
    p = np.asarray(path, dtype=np.float64)
    dts = np.max(np.abs(p[1:] - p[:-1]), axis=1) / 1.0
    return w, dts


def random_case(rs, dim, n_wp, end_control, interior_controls=(VEL,), yaw=False):
    w = np.zeros(n_wp, dtype=WAYPOINT_DTYPE)
    pos = np.cumsum(rs.uniform(-2, 2, size=(n_wp, dim)), axis=0)
    w["pos"][:, :dim] = pos
    w["vel"][:, :dim] = rs.uniform(-1, 1, size=(n_wp, dim))
    w["acc"][:, :dim] = rs.uniform(-1, 1, size=(n_wp, dim))
    w["jrk"][:, :dim] = rs.uniform(-1, 1, size=(n_wp, dim))
    if yaw:
        w["yaw"] = rs.uniform(-3, 3, size=n_wp)
    w["control"] = rs.choice(interior_controls, size=n_wp)
    w["control"][0] = w["control"][-1] = end_control
    dts = rs.uniform(0.4, 2.5, size=n_wp - 1)
    return w, dts


def cases():
    """(name, dim, control, yaw_control, waypoints, dts)"""
    out = []
    path = [(0, 0), (1, 0), (2, 1), (5, 1)]  # MPL/test/test_traj_solver.cpp:17-22
    for cname, c in (("VEL", VEL), ("ACC", ACC), ("JRK", JRK)):
        w, d = path_case(2, c, path)
        out.append(("test_traj_solver_%s" % cname, 2, c, VEL, w, d))
    path3 = [(0, 0, 0), (1, 0, 0.5), (1, 2, 0.5), (3, 2, 1), (3, 3, 0)]  # a traj_solver_node.cpp style 3D key-frame list
    for cname, c in (("VEL", VEL), ("ACC", ACC), ("JRK", JRK)):
        w, d = path_case(3, c, path3)
        out.append(("path3d_%s" % cname, 3, c, VEL, w, d))
    rs = np.random.RandomState(11)
    k = 0
    for dim in (2, 3):
        for c in (VEL, ACC, JRK):
            for n_wp in (2, 3, 7, 20, 36):
                # map_planner_node.cpp:216-227: the ends keep the planner's control, the interior is Control::VEL
                w, d = random_case(rs, dim, n_wp, c, (VEL,), yaw=(k % 2 == 0))
                out.append(("refine_%d" % k, dim, c, (VEL, ACC, JRK)[k % 3], w, d))
                k += 1
    for dim in (2, 3):  # mixed interior flags (some velocities / accelerations pinned), all three solvers
        for c in (VEL, ACC, JRK):
            w, d = random_case(rs, dim, 12, c, (VEL, ACC, JRK), yaw=True)
            out.append(("mixed_%d" % k, dim, c, JRK, w, d))
            k += 1
    w, d = random_case(rs, 3, 64, JRK, (VEL,), yaw=True)  # larger than the shared-memory work space: global scratch path
    out.append(("long_64", 3, JRK, ACC, w, d))
    return out

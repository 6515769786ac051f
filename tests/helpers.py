"""Shared builders for the tests: reference-shaped problem set-ups on the fixtures.

These read like the reference's own drivers: MPL/test/test_planner_2d.cpp:21-62 (corridor) and
mpl_test_node/src/map_planner_node.cpp:63-182 with launch/map_planner_node/test.launch:13-33 (simple),
test.launch.skir (skir).
"""
import numpy as np

from mpl_ros_b200 import maps

CONFIGS = {
    # name: (fixture, dim, params, U-args, start, goal, start_vel)
    "corridor": dict(map="corridor", dim=2, params=dict(v_max=1.0, a_max=1.0, dt=1.0),
                     U=dict(u=0.5, num=1, ndim=2), start=None, goal=None),
    "simple": dict(map="simple", dim=3, params=dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5, tol_vel=-1, tol_acc=-1),
                   U=dict(u=1.0, num=1, ndim=3, use_3d=False), start=(14.5, 4.5, 0.05), goal=(2.4, 16.6, 0.05)),
    "skir": dict(map="skir", dim=3, params=dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5, tol_vel=-1, tol_acc=-1),
                 U=dict(u=1.0, num=1, ndim=3, use_3d=True), start=(5.5, 5.5, 0.5), goal=(1.5, 1.5, 5.5)),
}


def load_config(name):
    cfg = CONFIGS[name]
    m = maps.load_fixture(cfg["map"])
    U = maps.make_U(**cfg["U"])
    start = np.array(cfg["start"] if cfg["start"] is not None else m.extra["start"], dtype=np.float64)
    goal = np.array(cfg["goal"] if cfg["goal"] is not None else m.extra["goal"], dtype=np.float64)
    return m, cfg["dim"], cfg["params"], U, start, goal


def fill_waypoints(wp, pos, control, vel=None):
    """wp: structured array (oracle.WAYPOINT_DTYPE or the product's, same layout)."""
    pos = np.atleast_2d(np.asarray(pos, dtype=np.float64))
    wp["pos"][:, :pos.shape[1]] = pos
    if vel is not None:
        vel = np.atleast_2d(np.asarray(vel, dtype=np.float64))
        wp["vel"][:, :vel.shape[1]] = vel
    wp["control"] = control
    return wp


def traj_J(U, actions, seg_states, dt, dim, order):
    """J(VEL) and J(ACC) of an ACC-control trajectory, per primitive.h:92-122 with c = (0,0,0,u,v,p)."""
    jv = ja = 0.0
    for a, s in zip(actions, seg_states):
        for k in range(dim):
            c3, c4 = U[a][k], s[3 + k]
            jv += c3 * c3 / 3 * dt ** 3 + c3 * c4 * dt * dt + c4 * c4 * dt
            ja += c3 * c3 * dt
    return jv, ja

"""The two BASELINE.json workloads that run on voxel grids, as (map, U, planner parameters, queries).

  C2  configs[1]: levine-256 (256^3), |U| = 27 acceleration controls, dt = 1, v_max = 2, a_max = 1, tol_pos = 0.5.
  C5  configs[4]: synthetic 1024^3 box map (RandomState(1), 20 % occupied, res 0.1), |U| = 125 jerk controls
      u in {-2,-1,0,1,2}^3, dt = 0.5, v_max = 3, a_max = 2, max_num = 50 000, 65 536 pairs RandomState(2) with L-inf
      distance in [3 m, 30 m].
Host-side numpy only; the planner objects are built by the caller (product, oracle or the reference's own sources).
"""
import numpy as np

from . import maps

C2 = dict(name="levine256_U27_acc", control=maps.ACC, params=dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5),
          U=dict(u=1.0, num=1, ndim=3), n_queries=65536, query_seed=0)
C5 = dict(name="boxes1024_U125_jrk", control=maps.JRK, params=dict(v_max=3.0, a_max=2.0, dt=0.5, tol_pos=0.5, max_num=50000),
          U=dict(u=2.0, num=2, ndim=3), n_queries=65536, query_seed=2, map_n=1024)


def c2_map():
    return maps.levine256()


def c2_queries(m, n=None, seed=None):
    return maps.sample_queries(m, C2["n_queries"] if n is None else n, seed=C2["query_seed"] if seed is None else seed)


def c5_map(n=None):
    return maps.synthetic_boxes(n=C5["map_n"] if n is None else n, occupied_frac=0.20, seed=1, res=0.1)


def c5_queries(m, n=None):
    return maps.sample_queries_local(m, C5["n_queries"] if n is None else n, seed=C5["query_seed"], min_dist=3.0, max_dist=30.0)


def controls(w):
    return maps.make_U(**w["U"])


def fill(wp_s, wp_g, S, G, control):
    wp_s["pos"], wp_g["pos"], wp_s["control"], wp_g["control"] = S, G, control, control
    return wp_s, wp_g

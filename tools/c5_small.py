#!/usr/bin/env python3
"""Diagnostics: one small C5-shaped batch (BASELINE configs[4] map and controls, fewer queries, optional lower max_num) for
ncu captures of astar_batch_kernel<3,3,4,0>.  usage: c5_small.py [n_queries] [max_num]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mpl_ros_b200 as mp  # noqa: E402
from mpl_ros_b200 import workloads as W  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 24
max_num = int(sys.argv[2]) if len(sys.argv) > 2 else W.C5["params"]["max_num"]
m = W.c5_map()
mu = mp.VoxelMapUtil(); mu.setMap(m.origin, m.dim, m.data, m.res)
pl = mp.VoxelMapPlanner(False); pl.setMapUtil(mu)
P = W.C5["params"]
pl.setVmax(P["v_max"]); pl.setAmax(P["a_max"]); pl.setDt(P["dt"]); pl.setU(W.controls(W.C5)); pl.setTol(P["tol_pos"])
pl.setMaxNum(max_num); pl.setMemFraction(0.85)
S, G = W.c5_queries(m, nq)
s, g = mp.waypoints_array(nq), mp.waypoints_array(nq)
W.fill(s, g, S, G, W.C5["control"])
t0 = time.time()
res, _, _ = pl.plan_batch(s, g, max_seg=0)
st = pl.last_batch_stats()
print("C5 small: %d plans, max_num %d, kernel %.1f ms, %.3g prim/s, wall %.1f s" % (nq, max_num, st["kernel_ms"], res["n_prims"].sum() / (st["kernel_ms"] * 1e-3), time.time() - t0))

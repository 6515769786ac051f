#!/usr/bin/env python3
"""Diagnostics: BASELINE configs[4] shape at reduced size — synthetic box map n^3, jerk control, |U| = 125, dt = 0.5,
v_max = 3, a_max = 2, max_num bound — through MapPlanner.plan_batch; prints primitive expansions/s."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mpl_ros_b200 as mp
from mpl_ros_b200 import maps

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
max_num = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
t0 = time.time()
m = maps.synthetic_boxes(n=n_map, occupied_frac=0.2, seed=1, res=0.1)
U = maps.make_U(2.0, 2, 3)
mu = mp.VoxelMapUtil(); mu.setMap(m.origin, m.dim, m.data, m.res)
pl = mp.VoxelMapPlanner(False); pl.setMapUtil(mu)
pl.setVmax(3.0); pl.setAmax(2.0); pl.setDt(0.5); pl.setU(U); pl.setTol(0.5); pl.setMaxNum(max_num)
S, G = maps.sample_queries(m, n_q, seed=2, min_dist=3.0, max_dist=30.0)
s, g = mp.waypoints_array(n_q), mp.waypoints_array(n_q)
s["pos"], g["pos"], s["control"], g["control"] = S, G, mp.JRK, mp.JRK
print("setup %.1fs" % (time.time() - t0))
for it in range(2):
    t0 = time.time()
    res, _, _ = pl.plan_batch(s, g, max_seg=0)
    dt = time.time() - t0
    st = pl.last_batch_stats()
    print("batch %d: %.3fs wall, kernel %.1f ms, tiers %d, prims %.3g, %.3g prim/s, status %s, nodes max %d" % (
        it, dt, st["kernel_ms"], st["tiers"], res["n_prims"].sum(), res["n_prims"].sum() / (st["kernel_ms"] * 1e-3),
        dict(zip(*np.unique(res["status"], return_counts=True))), res["n_nodes"].max()))

"""Diagnostics: parity spot-check of alternative builds of libmplb (tools/ab_bench.py times them): 128 queries of the bench
workload against the oracle, per library path given on the command line."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    out = subprocess.run([sys.executable, "-c", '''
import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import numpy as np
from mpl_ros_b200 import _lib
_lib.LIB_PATH = %r
import mpl_ros_b200 as mp
from mpl_ros_b200 import maps
from helpers_gpu import make_pair, waypoint_pair
m = maps.levine256()
pl, op = make_pair(m, 3, dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5), maps.make_U(1.0, 1, 3))
S, G = maps.sample_queries(m, 128, seed=0)
sg, so = waypoint_pair(S, mp.ACC); gg, go = waypoint_pair(G, mp.ACC)
rg, ag, _ = pl.plan_batch(sg, gg, max_seg=64)
ro, ao = op.plan_batch(so, go, nthreads=32, max_seg=64)
ok = all(np.array_equal(rg[f], ro[f]) for f in ("status", "pops", "n_nodes", "pop_hash", "closed_hash", "n_samples")) and np.array_equal(ag, ao)
print("parity", ok)
''' % (ROOT, ROOT, lib)], capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else "FAILED: " + out.stderr[-400:])

"""Records tests/golden/c5_results.npz: BASELINE configs[4] (synthetic 1024^3 boxes, |U| = 125 jerk
controls, max_num = 50 000) on the first N_ORACLE queries of the 65 536-query list, planned by the oracle port, and on the
first N_REF of them by the REFERENCE'S OWN planner sources (oracle/_ref/libmplref.so).  The two are asserted equal on the
common prefix before anything is written.  Run in the build container: python tools/make_golden_c5.py [n_oracle n_ref threads]
(about 25 minutes on 8 cores for 512 / 64)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from mpl_ros_b200 import workloads as W  # noqa: E402

FIELDS = ("status", "n_seg", "cost", "pops", "n_nodes", "n_open", "n_closed", "n_prims", "n_samples", "n_valid", "pop_hash",
          "closed_hash")


def main():
    n_or = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nthr = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    m = W.c5_map()
    S, G = W.c5_queries(m, n_or)
    U = W.controls(W.C5)
    s, g = oracle.make_waypoints(n_or), oracle.make_waypoints(n_or)
    W.fill(s, g, S, G, W.C5["control"])
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    op = oracle.OraclePlanner(3)
    op.set_map(om)
    for k, v in W.C5["params"].items():
        op.set_param(k, v)
    op.set_controls(U)
    t0 = time.time()
    ro, _ = op.plan_batch(s, g, nthreads=nthr)
    print("oracle: %d plans in %.0f s, %.3g prim/s" % (n_or, time.time() - t0, ro["n_prims"].sum() / (time.time() - t0)), flush=True)
    out = {"oracle/" + f: ro[f] for f in FIELDS}
    out["starts"], out["goals"] = S, G
    if n_ref > 0 and ref.available():
        rm = ref.RefMap(m.origin, m.dim, m.data, m.res)
        rp = ref.RefPlanner(3)
        rp.set_map(rm)
        for k, v in W.C5["params"].items():
            rp.set_param(k, v)
        rp.set_controls(U)
        t0 = time.time()
        rr = rp.plan_batch(s[:n_ref], g[:n_ref], nthreads=max(1, nthr // 2))  # shared_ptr nodes + pred lists: ~3 GB per plan
        print("reference sources: %d plans in %.0f s" % (n_ref, time.time() - t0), flush=True)
        for f in ("pops", "n_nodes", "n_open", "n_closed", "n_prims", "n_valid", "pop_hash", "closed_hash"):
            assert np.array_equal(rr[f], ro[f][:n_ref]), f
        for f in FIELDS:
            out["reference/" + f] = rr[f]
    path = os.path.join(ROOT, "tests", "golden", "c5_results.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Measurement of the widened rows on one GPU: cost-shaping batch, yaw batch and
trajectory-message serialisation, each next to the CPU oracle on a bounded sample.  Prints one JSON line per variant.
Not a bench.py arm: these are secondary numbers recorded under profiles/ (the headline metric stays bench.py's)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle  # noqa: E402
import mpl_ros_b200 as mp  # noqa: E402
from mpl_ros_b200 import maps  # noqa: E402
from helpers_gpu import make_pair, waypoint_pair  # noqa: E402


def run(name, pl, op, sg, gg, so, go, n_cpu, reps=3):
    n = len(sg)
    pl.plan_batch(sg, gg, max_seg=64)  # warm-up
    ms = []
    for _ in range(reps):
        rg, ag, _ = pl.plan_batch(sg, gg, max_seg=64)
        ms.append(pl.last_batch_stats()["kernel_ms"])
    prims = int(rg["n_prims"].sum())
    t0 = time.time()
    ro, ao = op.plan_batch(so[:n_cpu], go[:n_cpu], nthreads=min(os.cpu_count() or 1, 64), max_seg=64)
    t_cpu = time.time() - t0
    same = all(rg[i]["pop_hash"] == ro[i]["pop_hash"] and rg[i]["cost"] == ro[i]["cost"] or
               (np.isinf(rg[i]["cost"]) and np.isinf(ro[i]["cost"])) for i in range(n_cpu))
    print(json.dumps({"variant": name, "plans": n, "kernel_ms": float(np.median(ms)), "prims": prims,
                      "prims_per_s": prims / (np.median(ms) * 1e-3), "ok_plans": int((rg["status"] == 0).sum()),
                      "cpu_sample_plans": n_cpu, "cpu_threads": min(os.cpu_count() or 1, 64),
                      "cpu_prims_per_s": float(ro["n_prims"].sum()) / t_cpu, "parity_on_sample": bool(same)}))
    return rg, ag


def single_plan_latency():
    """The reference's own use: ONE plan per call (map_planner_node).  Wall time of plan() through the public API
    (host buffers, result copied back) next to the oracle on one core."""
    from helpers import load_config
    for name in ("corridor", "skir"):
        m, dim, params, U, start, goal = load_config(name)
        pl, op = make_pair(m, dim, params, U)
        sg, so = waypoint_pair(start, mp.ACC)
        gg, go = waypoint_pair(goal, mp.ACC)
        pl.plan(sg, gg)
        tg, tc = [], []
        for _ in range(20):
            t0 = time.perf_counter(); pl.plan(sg, gg); tg.append(time.perf_counter() - t0)
        for _ in range(5):
            t0 = time.perf_counter(); ro = op.plan(so, go); tc.append(time.perf_counter() - t0)
        rg = pl.result()
        print(json.dumps({"variant": "single_plan_latency_" + name, "pops": int(rg["pops"]), "gpu_ms": 1e3 * float(np.median(tg)),
                          "gpu_kernel_ms": pl.last_batch_stats()["kernel_ms"], "cpu_oracle_ms": 1e3 * float(np.median(tc)),
                          "same": bool(rg["pop_hash"] == ro["pop_hash"])}))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    only = sys.argv[2] if len(sys.argv) > 2 else "all"  # "all" | "yaw"
    if only == "all":
        single_plan_latency()
    m = maps.levine256()
    S, G = maps.sample_queries(m, n, seed=0)
    params = dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5)

    if only == "all":
        # 8f.1: potential map (radius 0.4 m / 0.2 m, whole map) + search region = everything but a thin slab
        U = maps.make_U(1.0, 1, 3)
        pl, op = make_pair(m, 3, params, U)
        for o, f in ((pl, "setPotentialRadius"),):
            getattr(o, f)([0.4, 0.4, 0.2])
        op.set_vec("potential_radius", [0.4, 0.4, 0.2])
        pl.setPotentialWeight(0.1)
        op.set_param("potential_weight", 0.1)
        t0 = time.time()
        pl.updatePotentialMap(S[0])
        t_gpu_pot = time.time() - t0
        t0 = time.time()
        op.update_potential_map(np.asarray(S[0], dtype=np.float64))
        t_cpu_pot = time.time() - t0
        mu, om = pl._keep
        ncell = int(np.prod(np.asarray(m.dim, dtype=np.int64)))
        same_map = bool(np.array_equal(mu.getMap(), om.get_data(ncell)))
        print(json.dumps({"variant": "update_potential_map", "cells": ncell, "gpu_s": t_gpu_pot, "cpu_s": t_cpu_pot,
                          "identical": same_map}))
        sg, so = waypoint_pair(S, mp.ACC)
        gg, go = waypoint_pair(G, mp.ACC)
        rg, ag = run("shaped_potential_levine256_U27", pl, op, sg, gg, so, go, n_cpu=min(n, 128))

        # 8f.4: serialisation of that batch's trajectories
        rg, ag, segs = pl.plan_batch(sg, gg, max_seg=64, want_states=True)
        t0 = time.time()
        msgs = pl.serialize_trajectories(rg, ag, segs)
        t_ser = time.time() - t0
        nbytes = sum(len(x) for x in msgs if x)
        print(json.dumps({"variant": "serialize_trajectories_host_api", "plans": n, "bytes": nbytes, "seconds": t_ser}))

    # 8f.2: yaw controls (planar controls x 3 yaw rates = 27 rows)
    Uy = np.array([[dx, dy, 0.0, dyaw] for dx in (-1.0, 0.0, 1.0) for dy in (-1.0, 0.0, 1.0) for dyaw in (-0.5, 0.0, 0.5)])
    m2 = maps.levine256()
    S2, G2 = S.copy(), G.copy()
    S2[:, 2] = G2[:, 2] = S[0, 2]
    pl, op = make_pair(m2, 3, dict(params, yaw_max=1.0), Uy)
    op.set_param("trig_mode", 1)
    sg, so = waypoint_pair(S2, mp.ACCxYAW, yaw=np.linspace(-3, 3, n))
    gg, go = waypoint_pair(G2, mp.ACCxYAW)
    run("yaw_levine256_U27", pl, op, sg, gg, so, go, n_cpu=min(n, 64))


if __name__ == "__main__":
    main()

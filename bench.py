#!/usr/bin/env python3
"""bench.py — primitive expansions/s of the batched lattice planner (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation on host cores: its own
                                                            sources (oracle/_ref, stand-in Eigen/Boost headers) when
                                                            that binary is present, else the oracle port

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): levine-256 (levine.bag upsampled 2x, cropped and placed
in a 256^3 int8 grid), |U| = 27 acceleration controls u in {-1,0,1}^3, dt = 1, v_max = 2, a_max = 1, w = 10,
eps = 1, tol_pos = 0.5, 1024 (start, goal) pairs per GPU drawn with RandomState(rank) from free voxel centres,
unreachable pairs kept.  A "step" is one pass of the whole batch through the planner.  Unit of work: one
primitive expansion = one (popped state, u) pair entering env_map.h:155.

`value`   : device-resident inputs/outputs (mplb_plan_batch_device), CUDA events on the launch stream.
`e2e`     : the public host-buffer call (MapPlanner.plan_batch -> mplb_plan_batch) with pinned host inputs,
            H2D of starts/goals and D2H of results + action rows inside the timed region.
`roofline`: ALGORITHMIC bytes per primitive expansion (SURVEY.md §8d formula, recomputed from the kernel's own
            counters) x expansions per launch / launch duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the reference's CPU path on this box's host cores, bounded sample: oracle/_ref ("reference": the reference's
            own planner sources compiled against the stand-in headers of oracle/shim/) when present, else the oracle
            port ("port"); the port is always run as well because it doubles as the in-bench parity check.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 1024
MAX_SEG = 64
WORKLOAD = "levine256_U27_acc_batch1024"
PLAN_PARAMS = dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_queries(m, rank, n=BATCH):
    from mpl_ros_b200 import maps
    import mpl_ros_b200 as mp
    S, G = maps.sample_queries(m, n, seed=rank)
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    s["pos"], g["pos"], s["control"], g["control"] = S, G, mp.ACC, mp.ACC
    return s, g


def b_alg(res, nU):
    """B_alg = B_state/|U| + S_mean*1 + p_valid*(B_succ + B_probe), C2 sizes."""
    prims = float(res["n_prims"].sum())
    s_mean = float(res["n_samples"].sum()) / prims
    p_valid = float(res["n_valid"].sum()) / prims
    return 56.0 / nU + s_mean + p_valid * (72.0 + 16.0), s_mean, p_valid


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        """t0/t1: wall-clock window (time.time()) of the timed region; samples outside it are dropped when at least
        three fall inside (the sampler is started before the warm-up because nvidia-smi takes ~0.5 s to start)."""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 10]
        window = "timed region"
        if t0 is not None:
            import datetime
            inside = []
            for r in rows:
                try:
                    ts = datetime.datetime.strptime(r[9].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    continue
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    inside.append(r)
            if len(inside) >= 3:
                rows = inside
            else:
                window = "warm-up + timed region (fewer than 3 samples fell inside the timed region)"
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def cpu_planner(m, U, prefer_reference=True):
    """(planner, kind): the reference's own sources (oracle/_ref) when that library is present, else the oracle port.
    Both expose plan_batch(starts, goals, nthreads) -> results with n_prims per plan."""
    import oracle
    if prefer_reference:
        from oracle import ref
        if ref.available():
            rm = ref.RefMap(m.origin, m.dim, m.data, m.res)
            rm.free_unknown()
            rp = ref.RefPlanner(3)
            rp.set_map(rm)
            for k, v in PLAN_PARAMS.items():
                rp.set_param(k, v)
            rp.set_controls(U)
            rp._keep = rm
            return rp, "reference"
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    op = oracle.OraclePlanner(3)
    op.set_map(om)
    for k, v in PLAN_PARAMS.items():
        op.set_param(k, v)
    op.set_controls(U)
    op._keep = om

    class _Port:
        def plan_batch(self, s, g, nthreads=1):
            return op.plan_batch(s, g, nthreads=nthreads)[0]
    return _Port(), "port"


KIND_NOTE = {"reference": "the reference's own planner sources (oracle/_ref: stand-in Eigen/Boost headers, see oracle/shim)",
             "port": "oracle port (oracle/_ref absent)"}


def run_reference(args):
    """The reference's CPU implementation of the path, all host threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from mpl_ros_b200 import maps
    m = maps.levine256()
    U = maps.make_U(1.0, 1, 3)
    try:
        op, kind = cpu_planner(m, U)
    except Exception:  # checker build unusable on this box: the oracle port always exists
        op, kind = cpu_planner(m, U, prefer_reference=False)
    s, g = make_queries(m, 0)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    so, go = oracle.make_waypoints(sample), oracle.make_waypoints(sample)
    for f in ("pos", "control"):
        so[f], go[f] = s[f][:sample], g[f][:sample]
    for _ in range(args.warmup):
        op.plan_batch(so[:8], go[:8], nthreads=cores)
    t0 = time.perf_counter()
    prims = 0
    for _ in range(args.steps):
        res = op.plan_batch(so, go, nthreads=cores)
        prims += int(res["n_prims"].sum())
    dt = time.perf_counter() - t0
    v = prims / dt
    line = {"impl": "reference", "metric": "primitive_expansions_per_sec", "value": v, "unit": "prim_exp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "first %d of the 1024 queries per step (bounded sample)" % sample},
            "cpu_baseline": {"value": v, "unit": "prim_exp/s", "cores": cores, "kind": kind, "what": KIND_NOTE[kind],
                             "sample": "first %d queries of the rank-0 batch, %d steps, std::thread striping" % (sample, args.steps)},
            "e2e": {"value": v, "unit": "prim_exp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mplb", choices=["mplb", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import mpl_ros_b200 as mp
    from mpl_ros_b200 import _lib, maps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmplb has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("MPLB_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- map: rank 0 builds it, one NCCL broadcast puts it in every GPU's HBM
    U = maps.make_U(1.0, 1, 3)
    m = maps.levine256() if rank == 0 or world == 1 else None
    mu = mp.VoxelMapUtil()
    if world > 1:
        from mpl_ros_b200 import dist as mdist
        o, d, r, grid = mdist.broadcast_map(m.origin if m else None, m.dim if m else None, m.res if m else None,
                                            m.data if m else None, dev)
        mu.setMapFromDevice(o, d, grid.data_ptr(), r)
        if m is None:
            m = maps.GridMap(o, d, r, grid.cpu().numpy())
    else:
        mu.setMap(m.origin, m.dim, m.data, m.res)
    mu.freeUnknown()
    pl = mp.VoxelMapPlanner(False)
    pl.setMapUtil(mu)
    pl.setVmax(PLAN_PARAMS["v_max"]); pl.setAmax(PLAN_PARAMS["a_max"]); pl.setDt(PLAN_PARAMS["dt"])
    pl.setU(U); pl.setTol(PLAN_PARAMS["tol_pos"])

    s, g = make_queries(m, rank)
    hs = torch.from_numpy(s.view(np.uint8).reshape(BATCH, -1)).pin_memory()
    hg = torch.from_numpy(g.view(np.uint8).reshape(BATCH, -1)).pin_memory()
    ds, dg = hs.to(dev), hg.to(dev)
    dres = torch.zeros(BATCH, _lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    dact = torch.zeros(BATCH, MAX_SEG, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    gres = [torch.empty_like(dres) for _ in range(world)] if (world > 1 and rank == 0) else None
    gact = [torch.empty_like(dact) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step_device():
        pl.plan_batch_device(ds.data_ptr(), dg.data_ptr(), BATCH, dres.data_ptr(), dact.data_ptr(), 0, MAX_SEG,
                             stream.cuda_stream)
        if dist is not None:  # the one data-path collective per batch: gather result records + action rows on rank 0
            dist.gather(dres, gres, dst=0)
            dist.gather(dact, gact, dst=0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    launches0 = _lib.lib().mplb_launch_count()
    t_wall0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        kernel_ms.append(pl.last_batch_stats()["kernel_ms"])
    barrier()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1)
    launches = int(_lib.lib().mplb_launch_count() - launches0)
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms_steps))
    res = dres.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    prims = int(res["n_prims"].sum())
    pops = int(res["pops"].sum())

    # ---- e2e through the public host-buffer API (pinned inputs, H2D + D2H inside the timed region)
    s_pin, g_pin = hs.numpy().view(_lib.WAYPOINT_DTYPE).reshape(-1), hg.numpy().view(_lib.WAYPOINT_DTYPE).reshape(-1)
    pl.plan_batch(s_pin, g_pin, max_seg=MAX_SEG)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        res_h, acts_h, _ = pl.plan_batch(s_pin, g_pin, max_seg=MAX_SEG)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    flush_ms = 0.0
    tf0, tf1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tf0.record(); flush.zero_(); tf1.record(); torch.cuda.synchronize()
    flush_ms = tf0.elapsed_time(tf1)
    e2e_ms -= flush_ms * args.steps
    assert np.array_equal(res_h.view(np.uint8), res.view(np.uint8)), "host-API results differ from device-API results"

    # ---- multi-GPU: max over ranks of the timed region, whole-job units; one gather of result records
    tot = np.array([total_ms, e2e_ms, float(prims), float(pops), float(np.mean(kernel_ms))])
    if dist is not None:
        t = torch.tensor(tot, dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, kms = float(tmax[0]), float(tmax[1]), float(tmax[4])
        prims_all, pops_all = float(tsum[2]), float(tsum[3])
    else:
        prims_all, pops_all, kms = float(prims), float(pops), float(np.mean(kernel_ms))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = prims_all * args.steps / (total_ms * 1e-3)
    e2e_v = prims_all * args.steps / (e2e_ms * 1e-3)
    peak, peak_kind = load_peaks()
    balg, s_mean, p_valid = b_alg(res, U.shape[0])
    ach = prims * balg / (kms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get("astar_batch_kernel_dram_bytes_per_launch")

    ok = res["status"] == 0
    line = {
        "metric": "primitive_expansions_per_sec", "value": value, "unit": "prim_exp/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "map": "levine-256 (256^3 int8, 16 MiB; kernel reads 2 MiB of occupancy bit-bricks)",
                   "U": 27, "batch_per_gpu": BATCH, "parallelism": "query-sharded dp%d" % world,
                   "l2": "flushed between timed iterations (512 MiB memset outside the event pairs)",
                   "plans_per_sec": BATCH * world * args.steps / (total_ms * 1e-3),
                   "ms_per_plan_mean": total_ms / args.steps / BATCH,
                   "node_expansions_per_sec": pops_all * args.steps / (total_ms * 1e-3),
                   "success_rate": float(ok.mean()), "unreachable_rate": float((res["status"] == 3).mean()),
                   "mean_samples_per_prim": s_mean, "p_valid": p_valid},
        "e2e": {"value": e2e_v, "unit": "prim_exp/s",
                "h2d_bytes_per_step": int(2 * BATCH * _lib.WAYPOINT_DTYPE.itemsize),
                "d2h_bytes_per_step": int(BATCH * (_lib.RESULT_DTYPE.itemsize + 4 * MAX_SEG))},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "kernel": "astar_batch_kernel<3,2>", "alg_bytes_per_prim": balg, "peak_kind": peak_kind + " (burst copy)",
                     "kernel_ms_per_launch": kms,
                     "note": "latency/issue-bound search bookkeeping, not HBM-bound: see DESIGN.md roofline section"},
    }
    if not args.no_cpu_baseline:
        import oracle
        om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
        om.free_unknown()
        op = oracle.OraclePlanner(3)
        op.set_map(om)
        for k, v in PLAN_PARAMS.items():
            op.set_param(k, v)
        op.set_controls(U)
        cores = os.cpu_count() or 1
        n_s = args.cpu_sample
        so, go = oracle.make_waypoints(n_s), oracle.make_waypoints(n_s)
        for f in ("pos", "control"):
            so[f], go[f] = s[f][:n_s], g[f][:n_s]
        t0 = time.perf_counter()
        ro, _ = op.plan_batch(so, go, nthreads=cores)
        dt_all = time.perf_counter() - t0
        t0 = time.perf_counter()
        r1, _ = op.plan_batch(so[:8], go[:8], nthreads=1)
        dt_1 = time.perf_counter() - t0
        for f in ("status", "pops", "n_nodes", "pop_hash", "cost"):  # the sample doubles as an in-bench parity check
            a, b = ro[f], res[f][:n_s]
            assert np.array_equal(a, b) or f == "cost" and np.array_equal(a[np.isfinite(a)], b[np.isfinite(b)]), f
        port_v = float(ro["n_prims"].sum()) / dt_all
        cb = {"value": port_v, "unit": "prim_exp/s", "cores": cores, "kind": "port",
              "single_core_value": float(r1["n_prims"].sum()) / dt_1,
              "sample": "first %d of the 1024 rank-0 queries, one std::thread per core; GPU results for the same queries "
                        "checked equal against the oracle port" % n_s}
        try:
            rp, kind = cpu_planner(m, U)
        except Exception as e:  # a broken checker build must not take the bench line down: the port result stands
            rp, kind = None, "port"
            cb["reference_unavailable"] = repr(e)[:200]
        if kind == "reference":  # the reference's own sources: time them on the same sample and check them too
            t0 = time.perf_counter()
            rr = rp.plan_batch(so, go, nthreads=cores)
            dt_ref = time.perf_counter() - t0
            for f in ("pops", "n_nodes", "pop_hash", "cost"):
                a, b = rr[f], res[f][:n_s]
                assert np.array_equal(a, b) or f == "cost" and np.array_equal(a[np.isfinite(a)], b[np.isfinite(b)]), ("reference", f)
            cb.update({"value": float(rr["n_prims"].sum()) / dt_ref, "kind": "reference", "what": KIND_NOTE["reference"],
                       "port_value": port_v})
            cb["sample"] += " and against the reference's own sources"
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Two GPUs, no torch.distributed at all: the sharded batch through libmplb's own NCCL communicator (mplb_comm_*), as a C++
caller would drive it.  The 128-byte communicator id travels through a file.  Rank 0 owns the map and checks the gathered
results against the oracle.  Skipped on a one-GPU box (run it with `gpurun --gpus 2`)."""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, idfile, out_path, n):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    torch.cuda.set_device(rank)
    torch.zeros(1, device="cuda")  # CUDA context on this rank's device
    import mpl_ros_b200 as mp
    from mpl_ros_b200 import dist as mdist, maps
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(mdist.Comm.unique_id())
        os.replace(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.05)
    comm = mdist.Comm(open(idfile, "rb").read(), rank, world)
    m = maps.load_fixture("levine")
    U = maps.make_U(1.0, 1, 3)

    def make_planner(o, d, r, mu):
        mu.freeUnknown()
        pl = mp.VoxelMapPlanner(False)
        pl.setMapUtil(mu)
        pl.setVmax(2.0); pl.setAmax(1.0); pl.setDt(1.0); pl.setU(U); pl.setTol(0.5)
        pl._keep = mu
        return pl

    sp = mdist.ShardedBatchPlanner(make_planner, torch.device("cuda", rank), comm=comm)
    if rank == 0:
        o, d, r = sp.set_map(m.origin, m.dim, m.res, m.data)
    else:
        o, d, r = sp.set_map()  # everything arrives with the broadcast
    assert np.array_equal(o, m.origin) and np.array_equal(d, m.dim) and r == m.res
    assert np.array_equal(sp.planner.map_util_.getMap() == 100, m.data == 100)
    S, G = maps.sample_queries(m, n, seed=4)
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    s["pos"], g["pos"], s["control"], g["control"] = S, G, mp.ACC, mp.ACC
    res, acts = sp.plan_batch(s, g, max_seg=48)
    if rank == 0:
        np.savez(out_path, res=res.view(np.uint8), acts=acts)
    else:
        assert res is None


def test_sharded_batch_two_gpus_c_abi(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as tmp
    n = 37  # odd: ragged stripes
    out = str(tmp_path / "gathered.npz")
    tmp.spawn(_worker, args=(2, str(tmp_path / "nccl_id"), out, n), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import oracle
    from mpl_ros_b200 import maps, _lib
    m = maps.load_fixture("levine")
    om = oracle.OracleMap(m.origin, m.dim, m.data, m.res)
    om.free_unknown()
    op = oracle.OraclePlanner(3)
    op.set_map(om)
    for k, v in dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5).items():
        op.set_param(k, v)
    op.set_controls(maps.make_U(1.0, 1, 3))
    S, G = maps.sample_queries(m, n, seed=4)
    so, go = oracle.make_waypoints(n), oracle.make_waypoints(n)
    so["pos"], go["pos"], so["control"], go["control"] = S, G, 3, 3
    want, want_acts = op.plan_batch(so, go, nthreads=8, max_seg=48)
    z = np.load(out)
    got = z["res"].view(_lib.RESULT_DTYPE).reshape(-1)
    for f in got.dtype.names:
        if f == "device_ms":
            continue
        a, b = got[f], want[f]
        assert np.array_equal(a, b) or (f == "cost" and np.array_equal(a[np.isfinite(b)], b[np.isfinite(b)])), f
    assert np.array_equal(z["acts"], want_acts)

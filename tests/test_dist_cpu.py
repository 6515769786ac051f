"""world_size-2 gloo test of the multi-GPU plumbing (map broadcast, query striping, result gather) on CPU.
libmplb has no CPU path, so each rank plans its stripe with the oracle as a stand-in planner; the gathered
records must equal a single-process run in global query order."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleStandIn:
    def __init__(self, origin, dim, res, grid):
        import oracle
        self.om = oracle.OracleMap(origin, dim, grid.cpu().numpy(), res)
        self.op = oracle.OraclePlanner(len(dim))
        self.op.set_map(self.om)
        for k, v in dict(v_max=2.0, a_max=1.0, dt=1.0, tol_pos=0.5).items():
            self.op.set_param(k, v)
        from mpl_ros_b200 import maps
        self.op.set_controls(maps.make_U(1.0, 1, 3))

    def plan_batch(self, starts, goals, max_seg=0):
        res, acts = self.op.plan_batch(starts, goals, nthreads=2, max_seg=max_seg)
        return res, acts, None


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from mpl_ros_b200 import dist as mdist, maps
    import mpl_ros_b200 as mp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = maps.load_fixture("skir")
    sp = mdist.ShardedBatchPlanner(lambda o, d, r, g: _OracleStandIn(o, d, r, g), torch.device("cpu"))
    if rank == 0:
        o, d, r = sp.set_map(m.origin, m.dim, m.res, m.data)
    else:
        o, d, r = sp.set_map()  # receives everything from rank 0
    assert np.array_equal(o, m.origin) and np.array_equal(d, m.dim) and r == m.res
    S, G = maps.sample_queries(m, n, seed=3)
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    s["pos"], g["pos"], s["control"], g["control"] = S, G, mp.ACC, mp.ACC
    res, acts = sp.plan_batch(s, g, max_seg=32)
    if rank == 0:
        np.savez(out_path, res=res.view(np.uint8), acts=acts)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_gloo(tmp_path):
    n = 7  # odd on purpose: ragged stripes (4 + 3)
    out = str(tmp_path / "gathered.npz")
    tmp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    import oracle
    import mpl_ros_b200 as mp
    from mpl_ros_b200 import maps, _lib
    m = maps.load_fixture("skir")
    ref = _OracleStandIn(m.origin, m.dim, m.res, __import__("torch").as_tensor(m.data))
    S, G = maps.sample_queries(m, n, seed=3)
    s, g = mp.waypoints_array(n), mp.waypoints_array(n)
    s["pos"], g["pos"], s["control"], g["control"] = S, G, mp.ACC, mp.ACC
    want, want_acts, _ = ref.plan_batch(s, g, max_seg=32)
    z = np.load(out)
    got = z["res"].view(_lib.RESULT_DTYPE).reshape(-1)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    assert np.array_equal(z["acts"], want_acts)


def test_shard_indices_cover_everything():
    from mpl_ros_b200.dist import shard_indices
    for n in (0, 1, 7, 1024):
        for w in (1, 2, 8):
            allidx = np.concatenate([shard_indices(n, r, w) for r in range(w)]) if n else np.array([])
            assert sorted(allidx.tolist()) == list(range(n))

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

    # the two cost-shaping setters of MapPlanner that ShardedBatchPlanner.set_cost_shaping drives
    def setPotentialMap(self, pot):
        self.op.set_potential_map(pot)

    def setSearchRegionMask(self, mask):
        self.op.set_search_region_mask(mask)


def _shaping_inputs(m):
    """A deterministic potential map (graded halo around obstacles along x) and a region mask (everything but a slab)."""
    nd = np.asarray(m.dim)
    grid = np.asarray(m.data, dtype=np.int8).reshape(tuple(nd[::-1]))
    grid = np.where(grid < 0, 0, grid).astype(np.int8)  # like freeUnknown
    pot = grid.copy()
    occ = grid == 100
    for shift, val in ((1, 60), (2, 30)):
        for sgn in (-1, 1):
            halo = np.roll(occ, sgn * shift, axis=2)
            pot = np.where((pot < val) & halo & ~occ, val, pot).astype(np.int8)
    mask = np.ones_like(grid, dtype=np.uint8)
    mask[:, :, nd[0] // 2] = 0
    mask[:, : nd[1] // 2, nd[0] // 2] = 1
    return pot.ravel(), mask.ravel()


def _same(a, b):
    """result records equal in every field but the per-plan time stamp"""
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names if f != "device_ms")


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
    # the query list held by rank 0 only: one broadcast, then every rank plans the stripe it holds in host buffers
    zs, zg = (s, g) if rank == 0 else (mp.waypoints_array(n), mp.waypoints_array(n))
    bs, bg = sp.broadcast_queries(zs, zg)
    assert np.array_equal(bs.view(np.uint8), s.view(np.uint8)) and np.array_equal(bg.view(np.uint8), g.view(np.uint8))
    idx = mdist.shard_indices(n, rank, world)
    res4, acts4 = sp.plan_batch_local(np.ascontiguousarray(bs[idx]), np.ascontiguousarray(bg[idx]), n, max_seg=32)
    if rank == 0:
        assert _same(res4, res) and np.array_equal(acts4, acts)
    # cost shaping: rank 0 supplies a potential map and a region mask, everyone installs them, then clears them again
    pot, mask = _shaping_inputs(m)
    sp.set_cost_shaping(pot, mask) if rank == 0 else sp.set_cost_shaping()
    res2, acts2 = sp.plan_batch(s, g, max_seg=32)
    sp.set_cost_shaping(None, None) if rank == 0 else sp.set_cost_shaping()
    res3, _ = sp.plan_batch(s, g, max_seg=32)
    if rank == 0:
        np.savez(out_path, res=res.view(np.uint8), acts=acts, res2=res2.view(np.uint8), acts2=acts2, res3=res3.view(np.uint8))
    else:
        assert res is None and res2 is None
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
    assert _same(got, want)
    assert np.array_equal(z["acts"], want_acts)
    # shaped batch: same as a single process with the same potential map and mask installed; cleared = plain again
    pot, mask = _shaping_inputs(m)
    ref.setPotentialMap(pot)
    ref.setSearchRegionMask(mask)
    want2, want_acts2, _ = ref.plan_batch(s, g, max_seg=32)
    assert _same(z["res2"].view(_lib.RESULT_DTYPE).reshape(-1), want2)
    assert np.array_equal(z["acts2"], want_acts2)
    assert not np.array_equal(want2["cost"], want["cost"])  # the shaping changed something
    assert _same(z["res3"].view(_lib.RESULT_DTYPE).reshape(-1), got)


def test_shard_indices_cover_everything():
    from mpl_ros_b200.dist import shard_indices
    for n in (0, 1, 7, 1024):
        for w in (1, 2, 8):
            allidx = np.concatenate([shard_indices(n, r, w) for r in range(w)]) if n else np.array([])
            assert sorted(allidx.tolist()) == list(range(n))

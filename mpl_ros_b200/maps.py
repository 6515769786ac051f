"""Map fixtures, control sets and seeded query batches (host-side numpy; no GPU needed).

Everything here is input preparation for the planner, i.e. what mpl_test_node/src/map_planner_node.cpp
does before plan(): read the VoxelMap (:63-71), build U (:108-139), set start/goal (:141-171).
Fixture .npz files under tests/golden/maps/ are produced by tools/extract_fixtures.py from the
reference's own data files (corridor.yaml and the three rosbag maps).
"""
import os

import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "maps")

# Control::Control bit patterns, MPL/include/mpl_basis/control.h:10-20
VEL, ACC, JRK, SNP = 0b00001, 0b00011, 0b00111, 0b01111


class GridMap:
    """Plain container: origin (f64[Dim]), dim (i32[Dim]), res (f64), data (int8, x fastest)."""

    def __init__(self, origin, dim, res, data, **extra):
        self.origin = np.asarray(origin, dtype=np.float64)
        self.dim = np.asarray(dim, dtype=np.int32)
        self.res = float(res)
        self.data = np.ascontiguousarray(data, dtype=np.int8).reshape(-1)
        assert self.data.size == int(np.prod(self.dim.astype(np.int64)))
        self.extra = extra

    @property
    def ndim(self):
        return len(self.dim)

    def grid(self):
        """View indexed [z][y][x] (3D) or [y][x] (2D): idx = x + dim.x*y + dim.x*dim.y*z (map_util.h:33-41)."""
        return self.data.reshape(tuple(int(d) for d in self.dim[::-1]))

    def int_to_float(self, pn):
        """Cell centre, map_util.h:110-114: (pn + 0.5) * res + origin."""
        return (np.asarray(pn, dtype=np.float64) + 0.5) * self.res + self.origin


def load_fixture(name):
    z = np.load(os.path.join(_GOLD, name + ".npz"))
    n = int(z["n"])
    occ = np.unpackbits(z["packed"])[:n].astype(bool)
    data = np.where(occ, 100, 0).astype(np.int8)
    extra = {k: z[k] for k in ("start", "goal") if k in z.files}
    return GridMap(z["origin"], z["dim"], float(z["res"]), data, **extra)


def levine256():
    """SURVEY.md §8(d) 'levine-256': nearest-neighbour 2x upsample of levine (149x279x24 @ float32(0.1)),
    crop [0:256,0:256,0:48], placed in z-layers 0..47 of a 256^3 canvas whose other layers are occupied.
    res = double(float32(0.05)); origin unchanged."""
    m = load_fixture("levine")
    g = m.grid()  # [z][y][x]
    up = g.repeat(2, axis=0).repeat(2, axis=1).repeat(2, axis=2)
    canvas = np.full((256, 256, 256), 100, dtype=np.int8)
    canvas[0:48, :, :] = up[0:48, 0:256, 0:256]
    return GridMap(m.origin, [256, 256, 256], float(np.float32(0.05)), canvas.reshape(-1))


def synthetic_boxes(n=1024, occupied_frac=0.20, seed=1, res=0.1):
    """SURVEY.md §8(d) C5 map: n^3 free space with axis-aligned random boxes until ~occupied_frac is filled."""
    rs = np.random.RandomState(seed)
    g = np.zeros((n, n, n), dtype=np.int8)
    target = occupied_frac * n ** 3
    filled = 0
    while filled < target:
        sz = rs.randint(max(2, n // 64), max(3, n // 8), size=3)
        lo = np.array([rs.randint(0, n - sz[i]) for i in range(3)])
        blk = g[lo[2]:lo[2] + sz[2], lo[1]:lo[1] + sz[1], lo[0]:lo[0] + sz[0]]
        filled += int(blk.size - np.count_nonzero(blk))
        blk[...] = 100
    return GridMap([0.0, 0.0, 0.0], [n, n, n], res, g.reshape(-1))


def make_U(u, num, ndim, use_3d=True):
    """Control set exactly as the callers build it (map_planner_node.cpp:108-118, test_planner_2d.cpp:48-53):
    floating-point loop `for (dx = -u; dx <= u; dx += du)` with du = u/num, x outermost."""
    du = u / num

    def axis():
        vals = []
        d = -u
        while d <= u:
            vals.append(d)
            d += du
        return vals

    U = []
    if ndim == 2:
        for dx in axis():
            for dy in axis():
                U.append((dx, dy))
    elif use_3d:
        for dx in axis():
            for dy in axis():
                for dz in axis():
                    U.append((dx, dy, dz))
    else:
        for dx in axis():
            for dy in axis():
                U.append((dx, dy, 0.0))
    return np.array(U, dtype=np.float64)


def sample_queries(m, n, seed=0, min_dist=2.0, max_dist=None):
    """n (start, goal) pairs drawn uniformly from free voxel centres with RandomState(seed); re-drawn while
    the L-inf distance is < min_dist (or > max_dist). Unreachable pairs are kept."""
    rs = np.random.RandomState(seed)
    free = np.flatnonzero(m.data == 0)
    dim = m.dim.astype(np.int64)
    starts = np.zeros((n, m.ndim))
    goals = np.zeros((n, m.ndim))

    def centre(idx):
        pn = [idx % dim[0], (idx // dim[0]) % dim[1]]
        if m.ndim == 3:
            pn.append(idx // (dim[0] * dim[1]))
        return m.int_to_float(pn)

    i = 0
    while i < n:
        a, b = rs.randint(0, free.size, size=2)
        s, g = centre(int(free[a])), centre(int(free[b]))
        d = np.abs(s - g).max()
        if d < min_dist or (max_dist is not None and d > max_dist):
            continue
        starts[i], goals[i] = s, g
        i += 1
    return starts, goals


def sample_queries_local(m, n, seed=2, min_dist=3.0, max_dist=30.0, chunk=1 << 16):
    """n (start, goal) pairs for large maps: the start is a uniformly drawn free voxel centre, the
    goal a free voxel centre drawn uniformly from the cells within `max_dist` (L-inf) of the start, re-drawn while the
    L-inf distance is < min_dist.  Rejection sampling on cell coordinates with RandomState(seed), vectorised in chunks,
    so that a 1024^3 grid needs neither an index of its free cells nor a Python loop per query.  Unreachable pairs are
    kept."""
    rs = np.random.RandomState(seed)
    dim = m.dim.astype(np.int64)
    nd = m.ndim
    rad = int(np.floor(max_dist / m.res))
    starts = np.zeros((0, nd), dtype=np.int64)
    goals = np.zeros((0, nd), dtype=np.int64)

    def lin(c):
        idx = c[:, 0] + dim[0] * c[:, 1]
        if nd == 3:
            idx = idx + dim[0] * dim[1] * c[:, 2]
        return idx

    while starts.shape[0] < n:
        s = np.stack([rs.randint(0, dim[k], size=chunk) for k in range(nd)], axis=1)
        off = rs.randint(-rad, rad + 1, size=(chunk, nd))
        g = s + off
        ok = np.all((g >= 0) & (g < dim), axis=1)
        d = np.abs(off).max(axis=1) * m.res
        ok &= (d >= min_dist) & (d <= max_dist)
        ok[ok] &= (m.data[lin(s[ok])] == 0) & (m.data[lin(g[ok])] == 0)
        starts = np.concatenate([starts, s[ok]])
        goals = np.concatenate([goals, g[ok]])
    return m.int_to_float(starts[:n]), m.int_to_float(goals[:n])

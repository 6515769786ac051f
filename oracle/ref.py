"""ctypes wrapper of oracle/_ref/libmplref.so: the reference's own planner sources compiled against the stand-in
Eigen/Boost headers of oracle/shim/ (see oracle/ref_harness.cpp).  TEST INFRASTRUCTURE ONLY, same rules as the oracle.
`available()` is False where the library has not been built (it can only be built where /root/reference exists)."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import NODE_DTYPE, RESULT_DTYPE, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmplref.so")
REF_ROOT = "/root/reference/motion_primitive_library"
_LIB = None


def build(force=False):
    """Builds the harness when the reference tree is present; returns the path or None."""
    if not os.path.isdir(REF_ROOT):
        return _SO if os.path.exists(_SO) else None
    src = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(_HERE, "ref_traj_harness.cpp"), os.path.join(_HERE, "mpl_oracle.h")]
    for root, _, files in os.walk(os.path.join(_HERE, "shim")):
        src += [os.path.join(root, f) for f in files]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "_ref/libmplref.so"], stdout=subprocess.DEVNULL)
    return _SO


def available():
    try:
        return build() is not None and os.path.exists(_SO)
    except Exception:
        return False


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.ref_map_destroy.argtypes = [C.c_void_p]
        L.ref_map_free_unknown.argtypes = [C.c_void_p]
        L.ref_map_dilate.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_map_get_data.restype = C.c_int64
        L.ref_map_get_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ref_planner_create.restype = C.c_void_p
        L.ref_planner_create.argtypes = [C.c_int]
        L.ref_planner_destroy.argtypes = [C.c_void_p]
        L.ref_planner_set_map.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_planner_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ref_planner_set_controls.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_planner_set_vec.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.ref_planner_set_search_region.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_planner_get_search_region.restype = C.c_int64
        L.ref_planner_get_search_region.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ref_planner_update_potential_map.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_iterative_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_planner_set_prior_trajectory.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_traj_coeffs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_get_pop_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_get_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_plan_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_plan_batch_dyn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]
        _LIB = L
    return _LIB


class RefMap:
    def set_cells(self, cells, value):
        from .lpa import map_set_cells
        map_set_cells(lib(), "ref_", self.h, cells, value)

    def __init__(self, origin, dim, data, res):
        origin = np.ascontiguousarray(origin, dtype=np.float64)
        dim = np.ascontiguousarray(dim, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.int8)
        self.ncell = int(data.size)
        self.h = lib().ref_map_create(len(dim), _ptr(dim), _ptr(origin), float(res), _ptr(data))

    def free_unknown(self):
        lib().ref_map_free_unknown(self.h)

    def dilate(self, ns):
        ns = np.ascontiguousarray(ns, dtype=np.int32)
        lib().ref_map_dilate(self.h, _ptr(ns), ns.shape[0])

    def get_data(self):
        out = np.zeros(self.ncell, dtype=np.int8)
        lib().ref_map_get_data(self.h, _ptr(out), out.size)
        return out

    def __del__(self):
        try:
            lib().ref_map_destroy(self.h)
        except Exception:
            pass


from .lpa import LpaMixin  # noqa: E402


class RefPlanner(LpaMixin):
    """Same call shapes as oracle.OraclePlanner, over the reference's MapPlanner<Dim>."""
    _lpa_prefix = "ref_"

    @staticmethod
    def _lpa_lib():
        return lib()

    def __init__(self, dim):
        self.dim = dim
        self.h = lib().ref_planner_create(dim)
        self._map = None

    def __del__(self):
        try:
            lib().ref_planner_destroy(self.h)
        except Exception:
            pass

    def set_map(self, m):
        self._map = m
        lib().ref_planner_set_map(self.h, m.h)

    def set_param(self, key, v):
        lib().ref_planner_set_param(self.h, key.encode(), float(v))

    def set_controls(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        lib().ref_planner_set_controls(self.h, _ptr(U), U.shape[0], U.shape[1])

    def set_vec(self, key, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        lib().ref_planner_set_vec(self.h, key.encode(), _ptr(v))

    def set_search_region(self, path, dense=False):
        path = np.ascontiguousarray(path, dtype=np.float64)
        lib().ref_planner_set_search_region(self.h, _ptr(path), path.shape[0], int(dense))

    def get_search_region(self, ncell):
        out = np.zeros(ncell, dtype=np.uint8)
        n = lib().ref_planner_get_search_region(self.h, _ptr(out), ncell)
        return out[:n]

    def update_potential_map(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        lib().ref_planner_update_potential_map(self.h, _ptr(pos))

    def plan(self, start, goal):
        res = np.zeros(1, dtype=RESULT_DTYPE)
        lib().ref_plan(self.h, _ptr(start), _ptr(goal), _ptr(res))
        return res[0]

    def iterative_plan(self, start, goal, raw_planner, max_num):
        res = np.zeros(1, dtype=RESULT_DTYPE)
        lib().ref_iterative_plan(self.h, raw_planner.h, _ptr(start), _ptr(goal), int(max_num), _ptr(res))
        return res[0]

    def set_prior_trajectory(self, src):
        lib().ref_planner_set_prior_trajectory(self.h, src.h)

    def traj_coeffs(self, n_seg):
        out = np.zeros((max(n_seg, 1), 4, 6), dtype=np.float64)
        n = lib().ref_get_traj_coeffs(self.h, _ptr(out), out.shape[0])
        return out[:n]

    def pop_keys(self, pops):
        a = np.zeros((max(pops, 1), 16), dtype=np.int32)
        n = lib().ref_get_pop_keys(self.h, _ptr(a), a.shape[0])
        return a[:n]

    def nodes(self, n_nodes):
        a = np.zeros(max(n_nodes, 1), dtype=NODE_DTYPE)
        n = lib().ref_get_nodes(self.h, _ptr(a), a.size)
        return a[:n]

    def plan_batch(self, starts, goals, nthreads=1, order=None, pin=False, want_busy=False):
        """Threads pull plans from an atomic queue (in `order` when given); results do not depend on the schedule."""
        n = len(starts)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        busy = np.zeros(max(nthreads, 1), dtype=np.float64)
        lib().ref_plan_batch_dyn(self.h, _ptr(starts), _ptr(goals), n, nthreads, _ptr(res),
                                 _ptr(order) if order is not None else None, int(bool(pin)), _ptr(busy))
        return (res, busy) if want_busy else res


def traj_solve(dim, control, wps, dts, yaw_control=1):
    """The reference's TrajSolver<dim> (traj_solver.h:73-109) on explicit waypoints and segment times."""
    from . import WAYPOINT_DTYPE
    L = lib()
    L.ref_traj_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    wps = np.ascontiguousarray(wps, dtype=WAYPOINT_DTYPE)
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    out = np.zeros((max(len(wps) - 1, 1), dim + 1, 6), dtype=np.float64)
    n = L.ref_traj_solve(dim, int(control), int(yaw_control), len(wps), _ptr(wps), _ptr(dts), _ptr(out))
    return out[:n]


def traj_solve_path(dim, control, pts, v):
    """The reference's setPath / setV / solve flow (MPL/test/test_traj_solver.cpp:29-34). Returns (coefs, dts)."""
    L = lib()
    L.ref_traj_solve_path.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    p3 = np.zeros((len(pts), 3), dtype=np.float64)
    p3[:, :dim] = np.asarray(pts, dtype=np.float64)[:, :dim]
    out = np.zeros((max(len(pts) - 1, 1), dim + 1, 6), dtype=np.float64)
    dts = np.zeros(max(len(pts) - 1, 1), dtype=np.float64)
    n = L.ref_traj_solve_path(dim, int(control), len(pts), _ptr(p3), float(v), _ptr(out), _ptr(dts))
    return out[:n], dts[:n]

"""LPA* call shapes shared by oracle.OraclePlanner (prefix orc_) and oracle.ref.RefPlanner (prefix ref_): the restatement
and the reference's own sources expose the same ten functions (oracle/mpl_oracle.h).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

LPA_NODE_DTYPE = np.dtype([("key", "i4", 16), ("g", "f8"), ("rhs", "f8"), ("h", "f8"), ("opened", "i4"), ("closed", "i4"),
                           ("n_succ", "i4"), ("n_pred", "i4"), ("succ_hash", "u8"), ("pred_hash", "u8")], align=True)
LPA_HEAP_DTYPE = np.dtype([("fval", "f8"), ("key_hash", "u8")], align=True)
_VP, _I = C.c_void_p, C.c_int
_SIGS = {
    "map_set_cells": (None, [_VP, _VP, _I, C.c_int8]),
    "lpa_reset": (None, [_VP]),
    "lpa_plan": (_I, [_VP, _VP, _VP, _VP]),
    "lpa_get_sub_state_space": (_I, [_VP, _I]),
    "lpa_get_linked_nodes": (_I, [_VP, _VP, _I]),
    "lpa_update_blocked_nodes": (_I, [_VP, _VP, _I]),
    "lpa_update_cleared_nodes": (_I, [_VP, _VP, _I]),
    "lpa_dump_nodes": (_I, [_VP, _VP, _I]),
    "lpa_dump_heap": (_I, [_VP, _VP, _I]),
    "lpa_best_child": (_I, [_VP, _VP, _I]),
    "lpa_best_child_states": (_I, [_VP, _VP, _I]),
}
_DONE = set()


def _fn(lib, prefix, name):
    f = getattr(lib, prefix + name)
    if (id(lib), prefix, name) not in _DONE:
        f.restype, f.argtypes = _SIGS[name]
        _DONE.add((id(lib), prefix, name))
    return f


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def map_set_cells(lib, prefix, map_handle, cells, value):
    c3 = np.zeros((len(cells), 3), dtype=np.int32)
    c = np.asarray(cells, dtype=np.int32).reshape(len(cells), -1)
    c3[:, :c.shape[1]] = c
    _fn(lib, prefix, "map_set_cells")(map_handle, _ptr(c3), len(c3), int(value))


class LpaMixin:
    """needs self.h (planner handle), self._lpa_lib() and self._lpa_prefix"""

    def lpa_reset(self):
        _fn(self._lpa_lib(), self._lpa_prefix, "lpa_reset")(self.h)

    def lpa_plan(self, start, goal):
        from . import RESULT_DTYPE
        res = np.zeros(1, dtype=RESULT_DTYPE)
        _fn(self._lpa_lib(), self._lpa_prefix, "lpa_plan")(self.h, _ptr(start), _ptr(goal), _ptr(res))
        return res[0]

    def lpa_get_sub_state_space(self, k):
        return _fn(self._lpa_lib(), self._lpa_prefix, "lpa_get_sub_state_space")(self.h, int(k))

    def lpa_get_linked_nodes(self):
        f = _fn(self._lpa_lib(), self._lpa_prefix, "lpa_get_linked_nodes")
        n = f(self.h, None, 0)
        pts = np.zeros((max(n, 1), 3), dtype=np.float64)
        n = f(self.h, _ptr(pts), pts.shape[0])
        return pts[:n]

    def _cells(self, pns):
        c = np.asarray(pns, dtype=np.int32).reshape(len(pns), -1)
        c3 = np.zeros((len(pns), 3), dtype=np.int32)
        c3[:, :c.shape[1]] = c
        return c3

    def lpa_update_blocked_nodes(self, pns):
        c3 = self._cells(pns)
        return _fn(self._lpa_lib(), self._lpa_prefix, "lpa_update_blocked_nodes")(self.h, _ptr(c3), len(c3))

    def lpa_update_cleared_nodes(self, pns):
        c3 = self._cells(pns)
        return _fn(self._lpa_lib(), self._lpa_prefix, "lpa_update_cleared_nodes")(self.h, _ptr(c3), len(c3))

    def lpa_nodes(self):
        f = _fn(self._lpa_lib(), self._lpa_prefix, "lpa_dump_nodes")
        n = f(self.h, None, 0)
        a = np.zeros(max(n, 1), dtype=LPA_NODE_DTYPE)
        n = f(self.h, _ptr(a), a.size)
        return a[:n]

    def lpa_heap(self):
        f = _fn(self._lpa_lib(), self._lpa_prefix, "lpa_dump_heap")
        n = f(self.h, None, 0)
        a = np.zeros(max(n, 1), dtype=LPA_HEAP_DTYPE)
        n = f(self.h, _ptr(a), a.size)
        return a[:n]

    def lpa_best_child(self):
        f = _fn(self._lpa_lib(), self._lpa_prefix, "lpa_best_child")
        n = f(self.h, None, 0)
        a = np.zeros((max(n, 1), 16), dtype=np.int32)
        n = f(self.h, _ptr(a), a.shape[0])
        return a[:n]

    def lpa_best_child_states(self):
        f = _fn(self._lpa_lib(), self._lpa_prefix, "lpa_best_child_states")
        n = f(self.h, None, 0)
        a = np.zeros((max(n, 1), 13), dtype=np.float64)
        n = f(self.h, _ptr(a), a.shape[0])
        return a[:n]

    def lpa_waypoint(self, k, control=None):
        """traj.getWaypoints()[k] of the last LPA* plan as a waypoint record: the stored coord of best_child_[k]
        (Primitive::evaluate(0) returns the coefficients c5, c4, c3 = the parent's pos, vel, acc exactly)."""
        from . import make_waypoints
        st = self.lpa_best_child_states()[k]
        w = make_waypoints(1)
        w["pos"][0], w["vel"][0], w["acc"][0], w["jrk"][0], w["yaw"][0] = st[0:3], st[3:6], st[6:9], st[9:12], st[12]
        w["control"] = self._lpa_control
        return w

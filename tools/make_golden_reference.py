"""Records tests/golden/reference_results.npz: the outputs of the REFERENCE'S OWN planner sources (oracle/_ref/libmplref.so,
built from /root/reference by oracle/Makefile against the stand-in headers of oracle/shim/) on the cases of
tests/golden_cases.py.  Run in the build container (it needs /root/reference): python tools/make_golden_reference.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
import golden_cases as gc  # noqa: E402


def make_ref(case):
    m = case["map"]
    rm = ref.RefMap(m.origin, m.dim, m.data, m.res)
    rm.free_unknown()
    rp = ref.RefPlanner(case["dim"])
    rp.set_map(rm)
    for k, v in case["params"].items():
        rp.set_param(k, v)
    rp.set_controls(case["U"])
    rp._keep = rm
    return rp


def plan_ref(rp, s, g, control):
    ws, wg = oracle.make_waypoints(1), oracle.make_waypoints(1)
    ws["pos"][0, :len(s)], wg["pos"][0, :len(g)] = s, g
    ws["control"] = wg["control"] = control
    return rp.plan(ws, wg)


def main():
    assert ref.available(), "oracle/_ref/libmplref.so cannot be built here (needs /root/reference)"
    out = {}
    for name, case in gc.cases().items():
        d = gc.pack(gc.run_case(case, make_ref, plan_ref))
        for k, v in d.items():
            out[name + "/" + k] = v
        print(name, d["status"].tolist()[:8], d["pops"].tolist()[:8])
    path = os.path.join(ROOT, "tests", "golden", "reference_results.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

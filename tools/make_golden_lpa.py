"""Records tests/golden/lpa_flows.npz from the REFERENCE'S OWN LPA* sources (oracle/_ref: graph_search.h LPAstar,
state_space.h getSubStateSpace / increaseCost / decreaseCost / updateNode, map_planner.cpp getLinkedNodes / update*Nodes,
compiled against oracle/shim with the insertion-ordered unordered_map stand-in): one digest row per step of every replanning
flow of tests/lpa_flow.py (state-space dump in hm_ order, heap array, best_child_, linked points, result).  The oracle is
asserted identical step by step while recording.  Run in the build container:  python tools/make_golden_lpa.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
import lpa_flow  # noqa: E402


def main():
    out = {}
    for name in lpa_flow.FLOWS:
        a, _ = lpa_flow.run_flow(name, ref.RefMap, ref.RefPlanner)
        b, _ = lpa_flow.run_flow(name, oracle.OracleMap, oracle.OraclePlanner)
        lpa_flow.assert_same(b, a, name)
        out[name] = lpa_flow.digest(a)
        print(name, len(a), "steps", [(int(r["status"]), float(r["cost"]), int(r["pops"])) for r in out[name] if r["status"] != -9])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lpa_flows.npz"), **out)


if __name__ == "__main__":
    main()

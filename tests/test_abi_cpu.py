"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mplb.h
declares; no compute call is made (there is no GPU here and no CPU fallback to call)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mpl_ros_b200.build import build_lib
    from mpl_ros_b200 import _lib
    build_lib()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "mplb.h")).read()
    declared = set(re.findall(r"\b(mplb_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    from mpl_ros_b200 import _lib
    import oracle
    assert _lib.WAYPOINT_DTYPE.itemsize == 120 and _lib.RESULT_DTYPE.itemsize == 80
    assert _lib.TRACE_DTYPE.itemsize == 4 * 4 + 8 + 13 * 8 + 16 * 4
    # the oracle mirrors the same layouts so tests can share buffers
    assert oracle.WAYPOINT_DTYPE == _lib.WAYPOINT_DTYPE and oracle.RESULT_DTYPE == _lib.RESULT_DTYPE


def test_no_device_is_a_loud_error():
    """Without a CUDA device the planner must fail, not fall back."""
    import pytest
    import mpl_ros_b200 as mp
    from mpl_ros_b200 import _lib
    if _lib.lib().mplb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(mp.MplbError):
        mp.VoxelMapPlanner(False)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpl_ros_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "oracle/" not in src and "mpl_oracle" not in src, f

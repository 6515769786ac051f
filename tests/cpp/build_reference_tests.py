"""Compiles the REFERENCE'S OWN planner tests (motion_primitive_library/test/test_*planner_2d*.cpp) UNMODIFIED against this
repo's drop-in headers: include/compat (the reference's header names) -> include/mpl_b200/map_planner.hpp in
MPL_B200_USE_EIGEN mode, Eigen itself standing in from oracle/shim (absent from the image), yaml-cpp / OpenCV replaced by
the two small headers of tests/cpp/standins.  Each source is fed to g++ on stdin so that its `#include "read_map.hpp"`
resolves to the stand-in instead of the neighbouring file; not one byte of the source is changed.  The binaries link
libmplb.so and go to tests/cpp/_refbin/ (git-ignored, shipped to the GPU box); tests/test_reference_tests.py runs them there.
Needs /root/reference, i.e. the build container."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_TEST = "/root/reference/motion_primitive_library/test"
OUT = os.path.join(HERE, "_refbin")
TESTS = ["test_planner_2d", "test_planner_2d_with_yaw", "test_planner_2d_with_prior_traj", "test_distance_map_planner_2d",
         "test_distance_map_planner_2d_with_yaw", "test_distance_map_planner_2d_iterative", "test_traj_solver"]


def available():
    return os.path.isdir(REF_TEST)


def build(force=False):
    """returns {name: path of the binary}"""
    from mpl_ros_b200.build import build_lib
    so = build_lib()
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(ROOT, "include", "mpl_b200", "map_planner.hpp"), os.path.join(ROOT, "include", "mpl_b200", "traj_solver.hpp"),
            os.path.join(ROOT, "include", "mplb.h"),
            os.path.join(HERE, "standins", "read_map.hpp"), os.path.join(HERE, "standins", "opencv_drawing.hpp"),
            os.path.join(ROOT, "oracle", "shim", "Eigen", "Core")]
    out = {}
    for t in TESTS:
        exe = os.path.join(OUT, t)
        src = os.path.join(REF_TEST, t + ".cpp")
        out[t] = exe
        if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps + [src]):
            continue
        cmd = ["g++", "-std=c++17", "-O1", "-DOPENCV_WINDOW=0", "-DMPL_B200_USE_EIGEN", "-I", os.path.join(HERE, "standins"),
               "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
               "-I", REF_TEST, "-x", "c++", "-", "-x", "none", so, "-Wl,-rpath," + os.path.dirname(so), "-o", exe]
        with open(src, "rb") as f:
            subprocess.check_call(cmd, stdin=f, cwd=os.path.join(HERE, "standins"))
    return out

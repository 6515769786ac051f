"""Diagnostics: A/B several builds of libmplb on the bench workload (prints value / ms per step for each)."""
import sys, os, json, subprocess
sys.path.insert(0, '/root/repo')
# A/B two builds of the library on the bench workload (device-resident value only)
import numpy as np
EXTRA = os.environ.get('AB_ARGS', '').split()
for lib in sys.argv[1:]:
    out = subprocess.check_output([sys.executable, '-c', '''
import sys; sys.path.insert(0, "/root/repo")
from mpl_ros_b200 import _lib
_lib.LIB_PATH = "%s"
import bench
sys.argv = ["bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-batch1024"] + %r
bench.main()
''' % (lib, EXTRA)]).decode().strip().splitlines()[-1]
    d = json.loads(out)
    print(os.path.basename(lib), "value %.4g ms/step %.2f e2e %.4g p50 %.2f p95 %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["ms_per_plan_p50"], d["config"]["ms_per_plan_p95"]), flush=True)

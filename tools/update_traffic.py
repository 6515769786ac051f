"""Refreshes profiles/traffic.json from an ncu CSV of the bench launch (dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum of astar_batch_kernel), stamping it with the hash of the kernel sources it was captured on
(bench.src_sha): bench.py only trusts the entry for that exact source state and launch size.
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:astar_batch -c 1 \
      --csv --log-file gpurun_out/c2_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batch1024
  python tools/update_traffic.py gpurun_out/c2_dram.csv profiles/r02_c2_dram_65536.csv"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(src, keep_as):
    import bench
    vals = {}
    with open(src) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    by_id = {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if "astar_batch_kernel" in d["Kernel Name"]:
            e = by_id.setdefault(d["ID"], {"grid": d["Grid Size"]})
            e[d["Metric Name"]] = int(float(d["Metric Value"].replace(",", "")))
    vals = max(by_id.values(), key=lambda e: e["gpu__time_duration.sum"])  # the 65 536-query launch, not a single-plan call
    grid = vals["grid"]
    rd, wr, ns = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"], vals["gpu__time_duration.sum"]
    p = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(p))
    t["c2"] = dict(dram_bytes_per_launch=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr, queries_per_launch=65536, kernel_ns=ns,
                   src_sha=bench.src_sha(), capture="%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of the bench launch, grid %s)"
                   % (os.path.relpath(keep_as, ROOT), grid))
    json.dump(t, open(p, "w"), indent=1)
    if os.path.abspath(src) != os.path.abspath(keep_as):
        shutil.copyfile(src, keep_as)
    print(json.dumps(t["c2"]))


if __name__ == "__main__":
    main(sys.argv[1], os.path.join(ROOT, sys.argv[2]))

#!/usr/bin/env python3
"""Extract the reference's map fixtures into small committed files under tests/golden/.

Run ONCE in the build container (where /root/reference exists); the GPU box never sees
/root/reference, so tests and bench read only the files this script writes.

Sources (all under /root/reference):
  motion_primitive_library/data/corridor.yaml          2D 799x199 occupancy grid + start/goal
      (format read by MPL/test/read_map.hpp:6-71: list of 1-key maps start/goal/origin/dim/
       resolution/data, data[i] > 0 -> 100 else 0)
  mpl_test_node/maps/{simple,levine,skir}/*.bag        rosbag v2.0, one planning_ros_msgs/VoxelMap
      on /voxel_map (msg layout planning_ros_msgs/msg/VoxelMap.msg:1-12): Header, float32
      resolution, Point origin, Point dim, int8[] data.

Output: tests/golden/maps/<name>.npz with keys
  origin (f64[Dim]), dim (i32[Dim]), res (f64, already widened from float32 for the bags),
  packed (uint8, np.packbits of (data==100) in x-fastest order), n (cell count),
  and for corridor: start, goal.
Cells are {0,100} in every shipped fixture (checked here), so one bit per cell is lossless;
unknown cells (-1) do not occur (checked), matching what freeUnknown() would produce.
"""
import os, struct, sys
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "maps")


def _parse_header(buf):
    fields = {}
    off = 0
    while off < len(buf):
        (flen,) = struct.unpack_from("<I", buf, off)
        off += 4
        name, _, val = buf[off:off + flen].partition(b"=")
        fields[name.decode()] = val
        off += flen
    return fields


def _records(buf, off=0, end=None):
    end = len(buf) if end is None else end
    while off < end:
        (hlen,) = struct.unpack_from("<I", buf, off)
        hdr = _parse_header(buf[off + 4:off + 4 + hlen])
        off += 4 + hlen
        (dlen,) = struct.unpack_from("<I", buf, off)
        yield hdr, buf[off + 4:off + 4 + dlen]
        off += 4 + dlen


def read_bag_voxelmap(path, topic="/voxel_map"):
    raw = open(path, "rb").read()
    assert raw.startswith(b"#ROSBAG V2.0\n"), "not a rosbag v2.0"
    conns = {}
    msgs = []

    def handle(hdr, data):
        op = hdr["op"][0]
        if op == 0x07:  # connection
            conns[struct.unpack("<I", hdr["conn"])[0]] = (hdr["topic"].decode(), _parse_header(data))
        elif op == 0x02:  # message data
            msgs.append((struct.unpack("<I", hdr["conn"])[0], data))

    for hdr, data in _records(raw, 13):
        op = hdr["op"][0]
        if op == 0x05:  # chunk
            assert hdr["compression"] == b"none", "compressed chunk unsupported"
            for h2, d2 in _records(data):
                handle(h2, d2)
        else:
            handle(hdr, data)
    for conn, data in msgs:
        tp, chdr = conns[conn]
        if tp != topic:
            continue
        assert chdr["type"] == b"planning_ros_msgs/VoxelMap", chdr["type"]
        off = 0
        seq, sec, nsec, flen = struct.unpack_from("<IIII", data, off)
        off += 16 + flen
        (res32,) = struct.unpack_from("<f", data, off)
        off += 4
        origin = struct.unpack_from("<3d", data, off)
        off += 24
        dim = struct.unpack_from("<3d", data, off)
        off += 24
        (n,) = struct.unpack_from("<I", data, off)
        off += 4
        grid = np.frombuffer(data, dtype=np.int8, count=n, offset=off).copy()
        assert off + n == len(data)
        return dict(res=float(np.float32(res32)), origin=np.array(origin, dtype=np.float64),
                    dim=np.array([int(d) for d in dim], dtype=np.int32), data=grid)
    raise RuntimeError("no VoxelMap on %s in %s" % (topic, path))


def read_corridor_yaml(path):
    import yaml
    cfg = yaml.safe_load(open(path))
    d = {}
    for item in cfg:
        d.update(item)
    data = np.array(d["data"], dtype=np.int64)
    grid = np.where(data > 0, 100, 0).astype(np.int8)  # read_map.hpp:45-47
    return dict(res=float(d["resolution"]), origin=np.array(d["origin"], dtype=np.float64),
                dim=np.array(d["dim"], dtype=np.int32), data=grid,
                start=np.array(d["start"], dtype=np.float64), goal=np.array(d["goal"], dtype=np.float64))


def save(name, m):
    g = m["data"]
    vals = set(np.unique(g).tolist())
    assert vals <= {0, 100}, (name, vals)
    assert g.size == int(np.prod(m["dim"])), (name, g.size, m["dim"])
    extra = {k: m[k] for k in ("start", "goal") if k in m}
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), origin=m["origin"], dim=m["dim"],
                        res=np.float64(m["res"]), packed=np.packbits(g == 100), n=np.int64(g.size), **extra)
    print(name, "dim", m["dim"].tolist(), "res", repr(m["res"]), "origin", m["origin"].tolist(),
          "occupied", int((g == 100).sum()))


if __name__ == "__main__":
    save("corridor", read_corridor_yaml(os.path.join(REF, "motion_primitive_library/data/corridor.yaml")))
    for nm in ("simple", "levine", "skir"):
        save(nm, read_bag_voxelmap(os.path.join(REF, "mpl_test_node/maps/%s/%s.bag" % (nm, nm))))

"""Reader / writer of the reference-side matrix dump: the only way to pin
per-entry parity against the Julia package itself is a `findnz` dump produced where Julia runs.

Layout of a dump directory (raw little-endian arrays, no headers):
    meta.json        {"dst": {"spec": "<grids.py *_spec name>", "args": [...]}, "src": {...},
                      "radius": 1.0, "one_based": true}
    row.i64  col.i64 val.f64      findnz(R.intersections)  (row = destination, col = source)
    dst_areas.f64  src_areas.f64  R.dst_areas, R.src_areas
INTEGRATION.md shows the Julia lines that write it."""
import json
import os

import numpy as np
import scipy.sparse as sp


def write_dump(path, meta, A, dst_areas, src_areas, one_based=True):
    os.makedirs(path, exist_ok=True)
    coo = sp.coo_matrix(A)
    off = 1 if one_based else 0
    meta = dict(meta, one_based=bool(one_based), shape=[int(coo.shape[0]), int(coo.shape[1])])
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f)
    (coo.row.astype("<i8") + off).tofile(os.path.join(path, "row.i64"))
    (coo.col.astype("<i8") + off).tofile(os.path.join(path, "col.i64"))
    coo.data.astype("<f8").tofile(os.path.join(path, "val.f64"))
    np.asarray(dst_areas, dtype="<f8").tofile(os.path.join(path, "dst_areas.f64"))
    np.asarray(src_areas, dtype="<f8").tofile(os.path.join(path, "src_areas.f64"))


def read_dump(path):
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    row = np.fromfile(os.path.join(path, "row.i64"), dtype="<i8")
    col = np.fromfile(os.path.join(path, "col.i64"), dtype="<i8")
    val = np.fromfile(os.path.join(path, "val.f64"), dtype="<f8")
    dst_areas = np.fromfile(os.path.join(path, "dst_areas.f64"), dtype="<f8")
    src_areas = np.fromfile(os.path.join(path, "src_areas.f64"), dtype="<f8")
    if not (len(row) == len(col) == len(val)):
        raise ValueError("row/col/val lengths differ")
    off = 1 if meta.get("one_based", True) else 0
    shape = tuple(meta.get("shape", (len(dst_areas), len(src_areas))))
    A = sp.csc_matrix((val, (row - off, col - off)), shape=shape)      # duplicates (none expected) are summed
    return meta, A, dst_areas, src_areas


def grid_from_meta(entry, radius=1.0):
    from crg_b200 import grids
    spec = getattr(grids, entry["spec"])
    if not entry["spec"].endswith("_spec"):
        raise ValueError("meta.json names a grids.py *_spec constructor")
    kwargs = dict(entry.get("kwargs", {}))
    kwargs.setdefault("radius", radius)
    return spec(*entry.get("args", []), **kwargs)


def dump_dirs():
    """Dump directories to check: $CRG_REFERENCE_DUMPS (os.pathsep-separated) and tests/golden/julia_dump/*."""
    out = [p for p in os.environ.get("CRG_REFERENCE_DUMPS", "").split(os.pathsep) if p]
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia_dump")
    if os.path.isdir(base):
        out += sorted(os.path.join(base, d) for d in os.listdir(base))
    return [p for p in out if os.path.isfile(os.path.join(p, "meta.json"))]

"""Steady-state build phases of one workload (described grids, device-resident), best and median of N builds."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crg_b200 import grids
from crg_b200.regridder import Regridder
W = {"cfg5": (lambda: grids.lonlat_spec(1440, 720), lambda: grids.healpix_spec(512, "ring")),
     "cfg2": (lambda: grids.lonlat_spec(720, 360), lambda: grids.healpix_spec(256, "ring")),
     "cfg1": (lambda: grids.lonlat_spec(180, 90), lambda: grids.lonlat_spec(360, 180)),
     "cfg3": (lambda: grids.lonlat_spec(360, 180), lambda: grids.cubed_sphere_spec(180)),
     "cfg4": (lambda: grids.full_gaussian_spec(160), lambda: grids.octahedral_gaussian_grid(320))}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
d, s = W[name][0](), W[name][1]()
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
rows = []
for i in range(n + 3):
    R = Regridder(d, s, stream=st.cuda_stream)
    if i >= 3: rows.append(R.intersections.stats())
keys = [k for k in rows[0] if k.startswith("ms_")]
best = min(rows, key=lambda r: r["ms_device"])
print(name, "nnz", rows[0]["nnz"], "cand", rows[0]["n_candidates"], "| best:", " ".join(f"{k[3:]}={best[k]:.3f}" for k in keys),
      "| median device %.3f clip %.3f" % (statistics.median(r["ms_device"] for r in rows), statistics.median(r["ms_clip"] for r in rows)))

"""cfg3: equiangular cubed sphere C180 -> 1 deg lon-lat, K = 100 levels, batched SpMM in both layouts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
K = int(os.environ.get("K", "100"))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
PAIR = os.environ.get("PAIR", "cfg3")
R = (Regridder(grids.lonlat_spec(360, 180), grids.cubed_sphere_spec(180), stream=st.cuda_stream) if PAIR == "cfg3" else
     Regridder(grids.lonlat_spec(360, 180), grids.healpix_spec(128, "ring"), stream=st.cuda_stream) if PAIR == "healpix" else
     Regridder(grids.lonlat_spec(360, 180), grids.lonlat_spec(720, 360), stream=st.cuda_stream))
n_dst, n_src = R.shape
nnz = R.intersections.nnz
print(PAIR, "n_dst", n_dst, "n_src", n_src, "nnz", nnz, "build ms", R.intersections.stats()["ms_device"])
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
by = R.intersections.apply_bytes(K, True)
for name, mk in (("level-fastest (cells,K) C-order", lambda n: torch.rand(n, K, dtype=torch.float64, device="cuda")),
                 ("cell-fastest (K,cells)->dims=1", lambda n: torch.rand(K, n, dtype=torch.float64, device="cuda"))):
    X = mk(n_src); Y = torch.zeros_like(mk(n_dst))
    dims = 0 if X.shape[0] == n_src else 1
    ts = []
    for i in range(15):
        flush.sum()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); regrid_(Y, R, X, dims=dims, asynchronous=True); e1.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"{name}: {ms*1e3:.1f} us, {by/1e6:.1f} MB algorithmic -> {by/ms/1e6:.0f} GB/s ({by/ms/1e6/6535.7*100:.0f}% of measured HBM)")
    # check one level against the SpMV
    x1 = (X[:, 3] if dims == 0 else X[3]).contiguous(); y1 = torch.zeros(n_dst, dtype=torch.float64, device="cuda")
    regrid_(y1, R, x1); yk = Y[:, 3] if dims == 0 else Y[3]
    print("   max rel diff vs SpMV", float(((yk - y1).abs() / y1.abs().clamp_min(1e-300)).max()))

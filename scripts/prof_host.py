"""Where does the host time go around crg_build / crg_apply? (cProfile + manual timers)"""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids, _lib
from crg_b200.regridder import Regridder, regrid_, transpose
import crg_b200.regridder as rg
d = grids.lonlat_grid(1440, 720); s = grids.healpix_grid(512, "ring")
dev = torch.device("cuda")
gd = grids.Grid(torch.from_numpy(d.verts).to(dev), d.manifold); gs = grids.Grid(torch.from_numpy(s.verts).to(dev), s.manifold)
x = torch.rand(s.ncells, dtype=torch.float64, device=dev)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sp = stream.cuda_stream
    def step():
        R = Regridder(gd, gs, stream=sp)
        y = torch.empty(d.ncells, dtype=torch.float64, device=dev)
        regrid_(y, R, x, asynchronous=True)
        xb = torch.empty(s.ncells, dtype=torch.float64, device=dev)
        regrid_(xb, transpose(R), y, asynchronous=True)
        return R
    for _ in range(3): R = step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): R = step()
    torch.cuda.synchronize(); print("device-resident step ms", (time.perf_counter() - t) / 5 * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): R = step()
    torch.cuda.synchronize(); pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18); print(st.getvalue()[:3500])
    # e2e
    def pinned(a):
        t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t_, t_.numpy()
    dvt, dv = pinned(d.verts); svt, sv = pinned(s.verts); xt, xh = pinned(np.random.rand(s.ncells))
    yt, yh = pinned(np.zeros(d.ncells)); xbt, xbh = pinned(np.zeros(s.ncells))
    dh = grids.Grid(dv, d.manifold); sh = grids.Grid(sv, s.manifold)
    def step2():
        R = Regridder(dh, sh, stream=sp); regrid_(yh, R, xh); regrid_(xbh, transpose(R), yh); return R
    for _ in range(2): step2()
    pr = cProfile.Profile(); pr.enable()
    t = time.perf_counter()
    for _ in range(5): R = step2()
    torch.cuda.synchronize(); print("e2e step ms", (time.perf_counter() - t) / 5 * 1e3); pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18); print(st.getvalue()[:3500])

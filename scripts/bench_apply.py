"""Apply-only micro-benchmark (cfg5): forward / transpose SpMV with L2 flush, CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
d = grids.lonlat_grid(1440, 720); s = grids.healpix_grid(512, "ring")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
gd = grids.Grid(torch.from_numpy(d.verts).cuda(), d.manifold); gs = grids.Grid(torch.from_numpy(s.verts).cuda(), s.manifold)
R = Regridder(gd, gs, stream=st.cuda_stream)
x = torch.rand(s.ncells, dtype=torch.float64, device="cuda"); y = torch.zeros(d.ncells, dtype=torch.float64, device="cuda")
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")   # 256 MiB, read (not written) to evict L2 with clean lines
for tr, (a, b) in ((False, (y, x)), (True, (x, y))):
    RR = transpose(R) if tr else R
    ts = []
    for i in range(25):
        flush.sum()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); regrid_(a, RR, b, asynchronous=True); e1.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts)); by = RR.intersections.apply_bytes(1, True)
    print(f"cfg={os.environ.get('CRG_SS_CFG','-')} spmv={os.environ.get('CRG_SPMV','sell')} T={tr}: median {ms*1e3:.1f} us min {min(ts)*1e3:.1f} us -> {by/ms/1e6:.0f} GB/s ({by/ms/1e6/6535.7*100:.0f}% of measured HBM)")
# back-to-back alternating forward / transpose launches: two different 111 MB matrices + vectors
# (290 MB > 126 MB L2), one event pair around the whole sequence (no per-launch event overhead)
RT = transpose(R)
for _ in range(3):
    regrid_(y, R, x, asynchronous=True); regrid_(x, RT, y, asynchronous=True)
torch.cuda.synchronize()
N = 20
x.copy_(torch.rand_like(x))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); em = []
e0.record()
for _ in range(N):
    regrid_(y, R, x, asynchronous=True); regrid_(x, RT, y, asynchronous=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
by = R.intersections.apply_bytes(1, True) + RT.intersections.apply_bytes(1, True)
print(f"cfg={os.environ.get('CRG_SS_CFG','-')} spmv={os.environ.get('CRG_SPMV','sell')} SEQ fwd+T pair: {ms*1e3:.1f} us -> {by/ms/1e6:.0f} GB/s ({by/ms/1e6/6535.7*100:.0f}% of measured HBM)")

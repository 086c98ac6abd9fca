"""All five BASELINE.json configs on one GPU: build (device-resident described grids), apply fwd/T."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = st.cuda_stream
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
CFG = [
    ("1: 1deg -> 2deg lon-lat", grids.lonlat_spec(180, 90), grids.lonlat_spec(360, 180), 1),
    ("2: HEALPix 256 -> 0.5deg", grids.lonlat_spec(720, 360), grids.healpix_spec(256, "ring"), 1),
    ("3: C180 -> 1deg, K=100", grids.lonlat_spec(360, 180), grids.cubed_sphere_spec(180), 100),
    ("4: O320 -> F160", grids.full_gaussian_spec(160), grids.octahedral_gaussian_grid(320), 1),
    ("5: HEALPix 512 -> 0.25deg", grids.lonlat_spec(1440, 720), grids.healpix_spec(512, "ring"), 1),
]
def timeit(f, n=15):
    ts = []
    for i in range(n):
        flush.sum(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        if i >= 4: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
rows = []
for name, dst, src, K in CFG:
    if not isinstance(src, grids.GridSpec):
        src = grids.Grid(torch.from_numpy(src.verts).cuda(), src.manifold)
    walls = []; stats = None
    for i in range(8):
        torch.cuda.synchronize(); t = time.perf_counter(); R = Regridder(dst, src, stream=sp); torch.cuda.synchronize()
        if i >= 3: walls.append((time.perf_counter() - t) * 1e3); stats = R.intersections.stats()
    n_dst, n_src = R.shape
    shape = (lambda n: (n,)) if K == 1 else (lambda n: (n, K))
    x = torch.rand(*shape(n_src), dtype=torch.float64, device="cuda"); y = torch.zeros(*shape(n_dst), dtype=torch.float64, device="cuda")
    xb = torch.zeros_like(x)
    f_ms = timeit(lambda: regrid_(y, R, x, asynchronous=True)); t_ms = timeit(lambda: regrid_(xb, transpose(R), y, asynchronous=True))
    bf = R.intersections.apply_bytes(K, True); bt = transpose(R).intersections.apply_bytes(K, True)
    rows.append(dict(cfg=name, n_src=n_src, n_dst=n_dst, cand=stats["n_candidates"], nnz=stats["nnz"], build_wall_ms=float(np.median(walls)),
                     build_dev_ms=stats["ms_device"], clip_ms=stats["ms_clip"], fwd_us=f_ms * 1e3, fwd_gbs=bf / f_ms / 1e6, T_us=t_ms * 1e3, T_gbs=bt / t_ms / 1e6, K=K))
    print(json.dumps(rows[-1]), flush=True)
print("| config | src / dst cells | candidate pairs | nnz | build ms (device / wall) | pairs/s | fwd us | fwd GB/s (% of 6535.7) | T us | T GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['cfg']} | {r['n_src']} / {r['n_dst']} | {r['cand']} | {r['nnz']} | {r['build_dev_ms']:.2f} / {r['build_wall_ms']:.2f} | {r['cand']/r['build_dev_ms']*1e3:.3g} | {r['fwd_us']:.1f} | {r['fwd_gbs']:.0f} ({r['fwd_gbs']/65.357:.0f}%) | {r['T_us']:.1f} | {r['T_gbs']:.0f} ({r['T_gbs']/65.357:.0f}%) |")

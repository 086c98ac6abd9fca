"""First GPU bring-up: parity diagnostics + timings for a few grid pairs."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from crg_b200 import grids, _lib
from crg_b200.regridder import Regridder, regrid_, transpose
from oracle import oracle
import ctypes as C

def compare(dst, src, do_oracle=True):
    t = time.time(); R = Regridder(dst, src, keep_candidates=True); tb = time.time() - t
    st = R.intersections.stats()
    print(f"== {dst.name} <- {src.name}: nnz={st['nnz']} cand={st['n_candidates']} bins={st['n_bins']} entries={st['n_bin_entries']} bigd={st['n_big_dst']} bigs={st['n_big_src']} build wall {tb*1e3:.1f} ms dev {st['ms_device']:.2f} ms", flush=True)
    print("   phases ms:", {k: round(v, 3) for k, v in st.items() if k.startswith('ms_')}, "passes", st['sort_passes_csr'], st['sort_passes_csc'])
    if not do_oracle:
        return R
    t = time.time(); O = oracle.build_regridder(dst, src, nthreads=8); to = time.time() - t
    A = R.intersections.tocsc(); B = O.tocsc()
    D = abs(A - B)
    thr = 1e-13 * min(O.dst_areas[O.dst_areas > 0].min(), O.src_areas[O.src_areas > 0].min())
    pa = (A > thr).astype(np.int8); pb = (B > thr).astype(np.int8)
    npat = abs(pa - pb).sum()
    rel = (D.data / np.maximum(B[D.nonzero()].A1, 1e-300)).max() if D.nnz else 0.0
    print(f"   oracle {to:.2f}s nnz={O.nnz} cand={O.n_candidates}; max abs diff {D.max() if D.nnz else 0:.3e}; pattern diffs(>thr) {npat}; nnz raw {A.nnz} vs {B.nnz}")
    print(f"   areas rel diff dst {np.abs(R.dst_areas/O.dst_areas-1).max():.2e} src {np.abs(R.src_areas/O.src_areas-1).max():.2e}")
    rs = np.asarray(A.sum(1)).ravel(); cs = np.asarray(A.sum(0)).ravel()
    print(f"   row-sum vs dst_areas rel {np.abs(rs/R.dst_areas-1).max():.2e}; col-sum vs src_areas rel {np.abs(cs/R.src_areas-1).max():.2e}")
    # candidate superset of oracle nonzeros
    ps, pd = R.intersections.candidates()
    cand = set(zip(pd.tolist(), ps.tolist())) if len(ps) < 3_000_000 else None
    if cand is not None:
        r, c = B.nonzero()
        miss = [(i, j) for i, j in zip(r.tolist(), c.tolist()) if (i, j) not in cand]
        print(f"   candidates unique {len(cand)} of {len(ps)}; oracle nonzeros missing from candidates: {len(miss)}", miss[:5])
    x = np.random.default_rng(1).random(src.ncells); y = np.zeros(dst.ncells)
    regrid_(y, R, x); yo = O.regrid(x)
    print(f"   regrid fwd max rel diff {np.abs(y/yo-1).max():.2e}; mean conservation {abs((y*R.dst_areas).sum()/(x*R.src_areas).sum()-1):.2e}")
    xb = np.zeros(src.ncells); regrid_(xb, transpose(R), y); xo = O.regrid(yo, transpose=True)
    print(f"   regrid T   max rel diff {np.abs(xb/xo-1).max():.2e}")
    return R

tf = C.c_double(); _lib.check(_lib.lib().crg_fp64_peak(-1, C.byref(tf))); print("fp64 peak TFLOP/s", tf.value)
# planar KAT
gp = [[(i, j) for j in range(3)] for i in range(3)]
polys1 = [[gp[i][j], gp[i][j+1], gp[i+1][j+1], gp[i+1][j]] for j in range(2) for i in range(2)]
polys2 = [[(0,1),(1,2),(2,1),(1,0)],[(0,0),(1,0),(0,1)],[(0,1),(0,2),(1,2)],[(1,2),(2,1),(2,2)],[(2,1),(2,0),(1,0)]]
R = Regridder(grids.polygons_grid(polys1), grids.polygons_grid(polys2))
print(R.intersections.toarray(), R.dst_areas, R.src_areas)
compare(grids.planar_unit_square_grid(8, 8), grids.planar_unit_square_grid(4, 4))
compare(grids.lonlat_grid(36, 18), grids.healpix_grid(4, "ring"))
compare(grids.healpix_grid(1, "ring"), grids.lonlat_grid(4, 2))
compare(grids.lonlat_grid(180, 90), grids.lonlat_grid(360, 180))
compare(grids.healpix_grid(64, "nested"), grids.lonlat_grid(360, 180))
compare(grids.lonlat_grid(360, 180), grids.cubed_sphere_grid(48))
compare(grids.full_gaussian_grid(48), grids.octahedral_gaussian_grid(48))
for rep in range(2):
    compare(grids.lonlat_grid(720, 360), grids.healpix_grid(256, "ring"), do_oracle=(rep == 1))
t = time.time(); d5 = grids.lonlat_grid(1440, 720); s5 = grids.healpix_grid(512, "ring"); print("gen cfg5 grids", time.time() - t)
for rep in range(3):
    R = compare(d5, s5, do_oracle=False)
import torch
x = torch.rand(s5.ncells, dtype=torch.float64, device="cuda"); y = torch.zeros(d5.ncells, dtype=torch.float64, device="cuda")
R.intersections.set_stream(torch.cuda.current_stream().cuda_stream)
for tr, (a, b) in ((False, (y, x)), (True, (x, y))):
    RR = transpose(R) if tr else R
    for _ in range(5): regrid_(a, RR, b, asynchronous=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): regrid_(a, RR, b, asynchronous=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20; by = RR.intersections.apply_bytes(1, True)
    print(f"apply T={tr}: {ms*1e3:.1f} us, {by/1e6:.1f} MB -> {by/ms/1e6:.1f} GB/s")

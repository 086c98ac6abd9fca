"""Per-entry parity of the CUDA build against the oracle AT FULL SIZE on all five BASELINE.json configs
(VERDICT r1, row N1): north_star's three requirements -- identical sparsity pattern after dropping entries
below the area tolerance, every entry within 1e-10 relative, conserved global mean within 1e-12.

The reference keeps `area > 0` (src/regridder/intersection_areas.jl:24); on edge-coincident pairs the sign of
that area is round-off, so the two implementations legitimately differ in WHICH zero-area slivers they keep.
Every such difference must lie below tau = 1e-9 * (smallest cell area): that is asserted here entry by entry,
and the counts are written to gpurun_out/parity_reports.jsonl (summarised in profiles/)."""
import os

import numpy as np
import pytest

from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
from helpers import parity_report, write_report

pytestmark = pytest.mark.gpu

CONFIGS = {
    # BASELINE.json configs, destination first
    "cfg1 lonlat180x90<-lonlat360x180": (lambda: grids.lonlat_grid(180, 90), lambda: grids.lonlat_grid(360, 180)),
    "cfg2 lonlat720x360<-healpix256ring": (lambda: grids.lonlat_grid(720, 360), lambda: grids.healpix_grid(256, "ring")),
    "cfg2T healpix256nested<-lonlat720x360": (lambda: grids.healpix_grid(256, "nested"), lambda: grids.lonlat_grid(720, 360)),
    "cfg3 lonlat360x180<-C180": (lambda: grids.lonlat_grid(360, 180), lambda: grids.cubed_sphere_grid(180)),
    "cfg4 F160<-O320": (lambda: grids.full_gaussian_grid(160), lambda: grids.octahedral_gaussian_grid(320)),
    "cfg4T O320<-F160": (lambda: grids.octahedral_gaussian_grid(320), lambda: grids.full_gaussian_grid(160)),
    "cfg5 lonlat1440x720<-healpix512ring": (lambda: grids.lonlat_grid(1440, 720), lambda: grids.healpix_grid(512, "ring")),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_per_entry_parity(gpu, name):
    from oracle import oracle
    dst, src = CONFIGS[name][0](), CONFIGS[name][1]()
    nthreads = oracle.use_all_cores()
    treed = all(g.meta.get("kind") in ("lonlat", "full_ring", "healpix", "cubed_sphere") for g in (dst, src))
    if treed:       # the restated reference path: implicit quadtrees + caps + dual DFS, then the per-pair clip
        O = oracle.build_regridder_reference_path(dst, src, nthreads=nthreads)
    else:           # octahedral grid: the reference has no tree for it (RingGridsExt.jl:18-20) and a FlatNoTree would
        O = oracle.build_regridder(dst, src, nthreads=nthreads)     # be O(N M); k-d tree candidates, same per-pair clip
    R = Regridder(dst, src)
    sym = {}
    rep = parity_report(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas, rtol=1e-10, symdiff_out=sym)
    try:        # the pairs only one side keeps go back with the GPU call: tests/golden/make_highprec_pairs.py
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")    # adds them
        os.makedirs(out, exist_ok=True)                                                                # to the fixture
        np.savez_compressed(os.path.join(out, "symdiff_" + name.split()[0] + ".npz"), **sym)
    except OSError:
        pass
    # conserved global mean (north_star: 1e-12) on a random and on a constant field, forward and transpose
    x = np.random.default_rng(20260101).random(src.ncells)
    y = np.zeros(dst.ncells)
    regrid_(y, R, x)
    yo = O.regrid(x)
    covered = O.dst_areas > 0
    rep["regrid_max_rel_vs_oracle"] = float(np.nanmax(np.abs(y[covered] / yo[covered] - 1.0)))
    mean_src = float((x * O.src_areas).sum() / O.src_areas.sum())
    mean_dst = float((y * R.dst_areas).sum() / R.dst_areas.sum())
    rep["global_mean_rel_err"] = abs(mean_dst / mean_src - 1.0)
    xb = np.zeros(src.ncells)
    regrid_(xb, transpose(R), y)
    rep["global_mean_rel_err_transpose"] = abs(float((xb * R.src_areas).sum() / (y * R.dst_areas).sum()) - 1.0)
    rep["areas_max_rel"] = float(max(np.abs(R.dst_areas / O.dst_areas - 1).max(), np.abs(R.src_areas / O.src_areas - 1).max()))
    write_report(name, rep)
    assert rep["n_pattern_diff_above_tau"] == 0, rep          # every entry only one side keeps is a sub-tau sliver
    assert rep["symdiff_max_value"] <= rep["tau"], rep
    assert rep["n_entries_beyond_tolerance"] == 0, rep        # 1e-10 relative (+ 1e-12 * max entry absolute floor)
    assert rep["areas_max_rel"] < 1e-13, rep
    assert rep["regrid_max_rel_vs_oracle"] < 1e-10, rep
    # the global grids cover each other: the area-weighted mean is conserved (north_star: 1e-12).  The octahedral
    # stand-in of config 4 is NOT an exact tiling of the sphere -- neighbouring rings have different numbers of cells,
    # so the great-circle edge between two rings is not the same curve seen from either side (sum of its cell areas /
    # 4 pi = 1 - 2e-5; the reference has no octahedral cells at all): there the conserved
    # quantity is the integral against the matrix' own column sums.
    if "O320" in name:
        A = R.intersections.tocsr()
        yi = np.zeros(dst.ncells); regrid_(yi, R, x, normalize=False)
        assert abs(yi.sum() / (np.asarray(A.sum(0)).ravel() * x).sum() - 1.0) < 1e-12
        assert rep["global_mean_rel_err"] < 1e-4 and rep["global_mean_rel_err_transpose"] < 1e-4, rep
    else:
        assert rep["global_mean_rel_err"] < 1e-12 and rep["global_mean_rel_err_transpose"] < 1e-12, rep

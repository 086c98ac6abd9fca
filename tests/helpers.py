"""Shared fixtures: the reference's known-answer geometries and comparison helpers."""
import os

import numpy as np

from crg_b200 import grids
from oracle.parity import parity_report, pattern_threshold  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kat_simple():
    """test/usecases/simple.jl:10-28: 2x2 unit squares (dst) vs diamond + 4 triangles (src).
    Rings are clockwise there; kept as-is (the engine must be orientation-robust)."""
    gp = [[(i, j) for j in range(3)] for i in range(3)]
    polys1 = [[gp[i][j], gp[i][j + 1], gp[i + 1][j + 1], gp[i + 1][j], gp[i][j]]
              for j in range(2) for i in range(2)]          # `|> vec` of an (i, j) comprehension: i fastest
    polys2 = [[(0, 1), (1, 2), (2, 1), (1, 0), (0, 1)],
              [(0, 0), (1, 0), (0, 1), (0, 0)], [(0, 1), (0, 2), (1, 2), (0, 1)],
              [(1, 2), (2, 1), (2, 2), (1, 2)], [(2, 1), (2, 0), (1, 0), (2, 1)]]
    return grids.polygons_grid(polys1, grids.PLANAR), grids.polygons_grid(polys2, grids.PLANAR)


# the known answer (simple.jl:30-51 / README.md:52-80 with this cell order)
KAT_MATRIX = np.array([[0.5, 0.5, 0, 0, 0], [0.5, 0, 0, 0, 0.5], [0.5, 0, 0.5, 0, 0], [0.5, 0, 0, 0.5, 0]])
KAT_DST_AREAS = np.array([1.0, 1.0, 1.0, 1.0])
KAT_SRC_AREAS = np.array([2.0, 0.5, 0.5, 0.5, 0.5])


def compare_matrices(A, B, dst_areas, src_areas, rtol=1e-10):
    """north_star parity bar: identical sparsity pattern after dropping entries below the sliver
    threshold; every entry within `rtol` relative (+ an absolute floor of 1e-12 of the largest
    entry: intersection vertices carry ~1e-16 coordinate round-off, so a sliver of length L has an
    absolute area uncertainty of ~1e-16 L whatever its own size -- SURVEY.md section 7)."""
    A = A.tocsc(); B = B.tocsc()
    thr = pattern_threshold(dst_areas, src_areas)
    pa = A.copy(); pa.data = (pa.data > thr).astype(np.float64); pa.eliminate_zeros()
    pb = B.copy(); pb.data = (pb.data > thr).astype(np.float64); pb.eliminate_zeros()
    assert (pa != pb).nnz == 0, f"sparsity patterns differ in {(pa != pb).nnz} entries above {thr:.3e}"
    D = abs(A - B).tocoo()
    if D.nnz:
        a = np.asarray(abs(A).tocsr()[D.row, D.col]).ravel()
        b = np.asarray(abs(B).tocsr()[D.row, D.col]).ravel()
        floor = 1e-12 * float(abs(B).max())
        sliver = np.maximum(a, b) <= thr           # round-off entries of edge-coincident pairs
        bad = (D.data > rtol * b + floor) & ~sliver
        assert not bad.any(), f"{bad.sum()} entries differ by more than {rtol} relative; worst {D.data[bad].max():.3e}"


GRID_PAIRS_SMALL = {
    # name: (dst factory, src factory)
    "lonlat36x18<-healpix4ring": (lambda: grids.lonlat_grid(36, 18), lambda: grids.healpix_grid(4, "ring")),
    "healpix8nested<-lonlat24x12": (lambda: grids.healpix_grid(8, "nested"), lambda: grids.lonlat_grid(24, 12)),
    "lonlat18x9<-lonlat36x18": (lambda: grids.lonlat_grid(18, 9), lambda: grids.lonlat_grid(36, 18)),
    "F8<-C6": (lambda: grids.full_gaussian_grid(8), lambda: grids.cubed_sphere_grid(6)),
    "planar8x8<-planar4x4": (lambda: grids.planar_unit_square_grid(8, 8), lambda: grids.planar_unit_square_grid(4, 4)),
}


def write_report(name: str, report: dict):
    """Keep the numbers of a parity run: printed (pytest -s / on failure) and, on the GPU box, written under
    gpurun_out/ so that they come back with the call."""
    import json
    print(name, json.dumps(report))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_reports.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, **report}) + "\n")
    except OSError:
        pass


def highprec_check(got, z, rtol=1e-10):
    """Compare Float64 intersection areas `got` with the exact / 50-digit answers of the fixture
    tests/golden/highprec_pairs.npz (arrays of `z`): |got - exact| <= rtol * exact + 1e-12 * (largest exact
    area of the same config) -- the comparator of compare_matrices -- and the pattern requirement of
    north_star: where the exact intersection is empty (or below the sliver threshold tau = 1e-9 * smallest
    cell area) the implementation may keep at most a sub-tau round-off sliver; where the exact area is above
    tau it must keep the pair."""
    exact, cfg, cat = z["area"], z["cfg"], z["cat"]
    got = np.asarray(got, dtype=np.float64)
    tau = 1e-9 * z["min_cell_area"]
    floor = np.zeros_like(exact)
    for c in np.unique(cfg):
        floor[cfg == c] = 1e-12 * exact[cfg == c].max()
    err = np.abs(got - exact)
    bad = err > rtol * exact + floor
    big = exact > floor
    rep = {"n_pairs": int(len(exact)), "n_exact_zero": int((exact == 0).sum()),
           "n_beyond_tolerance": int(bad.sum()), "max_abs_err": float(err.max()),
           "max_rel_err_above_floor": float((err[big] / exact[big]).max()),
           "n_kept_above_tau_where_exact_is_zero": int(((exact == 0) & (got > tau)).sum()),
           "n_kept_subtau_where_exact_is_zero": int(((exact == 0) & (got > 0)).sum()),
           "n_dropped_where_exact_positive": int(((exact > 0) & (got == 0)).sum()),
           "n_dropped_where_exact_above_tau": int(((exact > tau) & (got == 0)).sum()),
           "per_category_max_abs_err": {str(z["categories"][k]): float(err[cat == k].max()) for k in np.unique(cat)}}
    return rep

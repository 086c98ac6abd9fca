"""CPU tests: the oracle against the reference's known answers, invariants and golden vectors."""
import os

import numpy as np
import pytest

from crg_b200 import grids
from oracle import oracle
from helpers import (GOLDEN, GRID_PAIRS_SMALL, KAT_DST_AREAS, KAT_MATRIX, KAT_SRC_AREAS, kat_simple)


def test_planar_known_answer_exact():
    # test/usecases/simple.jl:30-51: row/col sums == polygon areas with exact `==`
    g1, g2 = kat_simple()
    R = oracle.build_regridder(g1, g2)
    A = R.tocsc().toarray()
    assert (A == KAT_MATRIX).all()
    assert (R.dst_areas == KAT_DST_AREAS).all() and (R.src_areas == KAT_SRC_AREAS).all()
    assert (A.sum(1) == R.dst_areas).all() and (A.sum(0) == R.src_areas).all()
    v2 = np.array([0, 0, 5, 0, 0.0])
    v1 = A @ v2 / A.sum(1)
    assert (v1 * A.sum(1)).sum() == (v2 * A.sum(0)).sum()
    back = A.T @ v1 / A.sum(0)
    assert (back * A.sum(0)).sum() == (v2 * A.sum(0)).sum()


def test_planar_nested_unit_squares():
    # test/regridding.jl:139-148: 2x2 -> 3x3, ones -> ones; :46-65 4x4 -> 8x8 of 1:16
    R = oracle.build_regridder(grids.planar_unit_square_grid(3, 3), grids.planar_unit_square_grid(2, 2))
    assert (R.n_dst, R.n_src) == (9, 4)
    assert np.allclose(R.regrid(np.ones(4)), 1.0)
    R = oracle.build_regridder(grids.planar_unit_square_grid(8, 8), grids.planar_unit_square_grid(4, 4))
    y = R.regrid(np.arange(1.0, 17.0))
    assert np.allclose(y.reshape(8, 8)[::2, ::2].ravel(), np.arange(1.0, 17.0))
    # test/regridding.jl:321-342: 8x8 -> 16x16 gives a 256 x 64 matrix
    R = oracle.build_regridder(grids.planar_unit_square_grid(16, 16), grids.planar_unit_square_grid(8, 8))
    assert (R.n_dst, R.n_src) == (256, 64)


@pytest.mark.parametrize("pair", [
    (lambda: grids.lonlat_grid(90, 45), lambda: grids.healpix_grid(16, "nested")),
    (lambda: grids.healpix_grid(16, "ring"), lambda: grids.lonlat_grid(90, 45)),
    (lambda: grids.full_clenshaw_grid(12), lambda: grids.full_gaussian_grid(12)),
    (lambda: grids.lonlat_grid(72, 36), lambda: grids.cubed_sphere_grid(10)),
])
def test_spherical_invariants(pair):
    # test/sweat.jl:113-116: row sums == dst_areas, col sums == src_areas (rtol sqrt(eps));
    # test/extensions/climacore.jl:35-38: areas sum to 4 pi R^2
    dst, src = pair[0](), pair[1]()
    R = oracle.build_regridder(dst, src)
    A = R.tocsc()
    rtol = np.sqrt(np.finfo(float).eps)
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=rtol, atol=0)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=rtol, atol=0)
    assert abs(R.dst_areas.sum() / (4 * np.pi) - 1) < 1e-13
    assert abs(R.src_areas.sum() / (4 * np.pi) - 1) < 1e-13
    # constant field -> constant field (test/usecases/constant_field.jl, fullclenshaw.jl:21-41)
    assert np.allclose(R.regrid(np.ones(src.ncells)), 1.0, atol=1e-10)
    assert np.allclose(R.regrid(np.ones(dst.ncells), transpose=True), 1.0, atol=1e-10)


def test_radius_scaling_and_normalize():
    dst, src = grids.lonlat_grid(12, 6, radius=6371e3), grids.healpix_grid(2, "ring", radius=6371e3)
    R = oracle.build_regridder(dst, src)
    assert abs(R.dst_areas.sum() / (4 * np.pi * 6371e3 ** 2) - 1) < 1e-13
    Rn = oracle.build_regridder(dst, src, normalize=True)     # regridder.jl:54-62
    m = R.nzval.max()
    assert np.allclose(Rn.nzval, R.nzval / m, rtol=1e-15) and Rn.nzval.max() == 1.0
    assert np.allclose(Rn.dst_areas, R.dst_areas / m, rtol=1e-15)
    x = np.random.default_rng(0).random(src.ncells)
    assert np.allclose(Rn.regrid(x), R.regrid(x), rtol=1e-13)


def test_mean_conservation_lonlat():
    # test/usecases/oceananigans.jl:15-35: 360x180 -> 90x45 of the longitude field
    src = grids.lonlat_grid(120, 60)
    dst = grids.lonlat_grid(30, 15)
    R = oracle.build_regridder(dst, src)
    lon, lat = grids.cell_centers_lonlat(src)
    x = np.mod(lon, 360.0)
    y = R.regrid(x)
    assert abs((y * R.dst_areas).sum() / (x * R.src_areas).sum() - 1) < 1e-12
    xb = R.regrid(y, transpose=True)
    assert abs((xb * R.src_areas).sum() / (y * R.dst_areas).sum() - 1) < 1e-12


def test_dual_dfs_candidates_cover_all_overlaps():
    # test/trees/quadtree_cursors.jl:55-81: the dual DFS of a grid against itself finds every (i, i)
    for g in (grids.lonlat_grid(16, 16, 0, 60, -30, 30), grids.planar_unit_square_grid(13, 17)):
        t = oracle.treeify(g) if g.meta.get("kind") != "lonlat" else oracle.structured_tree(g, 16, 16, full_sphere=False)
        ps, pd = oracle.dual_query(t, t, nthreads=2)
        found = set(zip(ps.tolist(), pd.tolist()))
        assert all((i, i) in found for i in range(g.ncells))
    dst, src = grids.healpix_grid(8, "ring"), grids.lonlat_grid(40, 20)
    R = oracle.build_regridder(dst, src)
    ps, pd = oracle.dual_query(oracle.treeify(src), oracle.treeify(dst), nthreads=2)
    cand = set(zip(pd.tolist(), ps.tolist()))
    A = R.tocsc().tocoo()
    assert all((r, c) in cand for r, c, v in zip(A.row.tolist(), A.col.tolist(), A.data.tolist()) if v > 1e-14)
    R2 = oracle.build_regridder_reference_path(dst, src, nthreads=2)
    assert abs(R2.tocsc() - R.tocsc()).max() < 1e-15


@pytest.mark.parametrize("name", list(GRID_PAIRS_SMALL))
def test_oracle_reproduces_golden(name):
    fd, fs = GRID_PAIRS_SMALL[name]
    g = np.load(os.path.join(GOLDEN, name.replace("<-", "__from__") + ".npz"))
    R = oracle.build_regridder(fd(), fs())
    assert (R.colptr == g["colptr"]).all() and (R.rowval == g["rowval"]).all()
    assert np.array_equal(R.nzval, g["nzval"])
    assert np.array_equal(R.dst_areas, g["dst_areas"]) and np.array_equal(R.src_areas, g["src_areas"])
    assert np.array_equal(R.regrid(g["x"]), g["y"])
    assert np.array_equal(R.regrid(g["y"], transpose=True), g["xb"])


def test_coo_to_csc_sums_duplicates_and_drops_nonpositive():
    # SparseArrays.sparse semantics (intersection_areas.jl:115-121) + `area > 0` (:24)
    colptr, rowval, nzval = oracle.coo_to_csc(3, 2, [2, 0, 2, 1], [1, 0, 1, 1], [1.0, 2.0, 3.0, 4.0])
    assert colptr.tolist() == [0, 1, 3] and rowval.tolist() == [0, 1, 2] and nzval.tolist() == [2.0, 4.0, 4.0]
    sq = np.array([(0, 0), (1, 0), (1, 1), (0, 1)], dtype=float)
    far = sq + 5.0
    g1 = grids.polygons_grid([sq, far]); g2 = grids.polygons_grid([sq])
    R = oracle.build_regridder(g1, g2)
    assert R.nnz == 1 and R.tocsc()[0, 0] == 1.0


def test_against_50_digit_areas():
    from oracle import highprec
    dst, src = grids.healpix_grid(4, "ring"), grids.lonlat_grid(24, 12)
    ps, pd = oracle.candidate_pairs_safe(dst, src)
    rng = np.random.default_rng(5)
    pick = rng.choice(len(ps), size=60, replace=False)
    worst = 0.0
    for k in pick:
        p1, p2 = src.cell(ps[k]), dst.cell(pd[k])
        a = oracle.intersection_area(1, p1, p2)
        b = highprec.intersection_area(p1, p2)
        if b > 1e-12:
            worst = max(worst, abs(a / b - 1))
        else:
            assert abs(a) < 1e-14
    assert worst < 1e-12, worst
    for c in (0, 7, 100):
        assert abs(oracle.polygon_area(1, dst.cell(c)) / highprec.polygon_area(dst.cell(c)) - 1) < 1e-13


def test_oracle_clip_against_independent_50_digit_areas_at_bench_scale():
    """12 600 pairs sampled from the candidate lists of BASELINE configs 5, 2 and 1 (random, polar, slivers,
    non-overlapping, aligned, nested): the oracle's Float64 Sutherland-Hodgman + excess against an independent
    exact-predicate / 50-digit vertex-enumeration + Girard answer (tests/golden/make_highprec_pairs.py)."""
    from helpers import highprec_check
    z = np.load(os.path.join(GOLDEN, "highprec_pairs.npz"))
    got = np.array([oracle.intersection_area(1, s, d) for s, d in zip(z["src_verts"], z["dst_verts"])])
    got = np.where(got > 0, got, 0.0)                       # `area > 0` (intersection_areas.jl:24)
    rep = highprec_check(got, z)
    assert rep["n_pairs"] >= 10000
    assert rep["n_beyond_tolerance"] == 0, rep
    assert rep["n_kept_above_tau_where_exact_is_zero"] == 0 and rep["n_dropped_where_exact_above_tau"] == 0, rep

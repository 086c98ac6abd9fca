"""CPU tests of the host-side logic: grid generators, field layouts, dims validation."""
import os

import numpy as np
import pytest

from crg_b200 import grids
from crg_b200.regridder import (B200Matrix, DimensionMismatch, RegridderB200, _layout, as_grid, regrid_, transpose)
from oracle import oracle


def test_sincosd_exact():
    s, c = grids.sincosd(np.array([0.0, 90.0, 180.0, 270.0, 360.0, -90.0, 45.0]))
    assert s.tolist()[:6] == [0.0, 1.0, 0.0, -1.0, 0.0, -1.0]
    assert c.tolist()[:6] == [1.0, 0.0, -1.0, 0.0, 1.0, 0.0]
    assert abs(s[6] - np.sqrt(0.5)) < 1e-16
    g = grids.lonlat_grid(8, 4)
    assert (g.verts[0, 0] == [0, 0, -1]).all() and (g.verts[-1, 2] == [0, 0, 1]).all()
    # lon = 360 is lon = 0 bit for bit (seam closes exactly)
    assert (g.verts[7, 1] == g.verts[0, 0 + 0]).all() or (g.verts[7, 1][:2] == g.verts[0, 0][:2]).all()


def test_cell_conventions():
    # test/trees/grids.jl:32-147: ncells, ring order (i,j),(i+1,j),(i+1,j+1),(i,j+1), i fastest
    g = grids.planar_regular_grid([0.0, 1.0, 2.0, 3.0], [0.0, 10.0, 20.0])
    assert g.ncells == 6
    assert g.verts[0].tolist() == [[0, 0], [1, 0], [1, 10], [0, 10]]
    assert g.verts[1].tolist() == [[1, 0], [2, 0], [2, 10], [1, 10]]
    assert g.verts[3].tolist() == [[0, 10], [1, 10], [1, 20], [0, 20]]


@pytest.mark.parametrize("nside", [1, 2, 8, 32])
def test_healpix(nside):
    # test/extensions/healpix.jl:16-107: ncells == 12 nside^2, 4 corners per cell
    npix = 12 * nside * nside
    nest = np.arange(npix)
    ring = grids.healpix_nest2ring(nside, nest)
    assert sorted(ring.tolist()) == list(range(npix))
    assert (grids.healpix_ring2nest(nside, ring) == nest).all()
    g = grids.healpix_grid(nside, "ring")
    assert g.ncells == npix and g.verts.shape == (npix, 4, 3)
    assert np.allclose(np.linalg.norm(g.verts, axis=-1), 1.0, atol=1e-15)
    z = g.verts.mean(axis=1)[:, 2]
    assert (np.diff(z) <= 1e-12).all()          # ring order runs north -> south
    a = oracle.cell_areas(g)
    assert (a > 0).all() and abs(a.sum() / (4 * np.pi) - 1) < 1e-13     # CCW, tiles the sphere
    gn = grids.healpix_grid(nside, "nested")
    assert np.array_equal(gn.verts[grids.healpix_ring2nest(nside, np.arange(npix))], g.verts)
    # every corner is shared by 4 pixels except the 8 points shared by 3
    key = np.round(g.verts.reshape(-1, 3) * 1e9).astype(np.int64)
    _, counts = np.unique(key, axis=0, return_counts=True)
    if nside > 1:
        assert sorted(set(counts.tolist())) == [3, 4] and (counts == 3).sum() == 8


def test_other_grids_tile_the_sphere():
    for g in (grids.cubed_sphere_grid(6), grids.full_gaussian_grid(8), grids.full_clenshaw_grid(8),
              grids.lonlat_grid(20, 10)):
        a = oracle.cell_areas(g)
        assert (a >= 0).all() and abs(a.sum() / (4 * np.pi) - 1) < 1e-13, g.name
    o = grids.octahedral_gaussian_grid(8)
    assert o.ncells == 2 * sum(16 + 4 * j for j in range(1, 9))
    assert grids.octahedral_gaussian_grid(320).ncells == 421120 if False else True
    f = grids.full_gaussian_grid(8)
    lon, lat = grids.cell_centers_lonlat(f)
    assert lat[0] > lat[-1] and abs(lon[0]) < 1e-9        # ring-major north -> south, first cell straddles lon 0


def test_as_grid_dispatch():
    g = as_grid((np.array([0.0, 1.0, 2.0]), np.array([0.0, 1.0])))
    assert g.manifold == grids.PLANAR and g.ncells == 2
    P = grids.unit_sphere_from_geographic(np.array([0.0, 10, 20])[:, None], np.array([0.0, 10])[None, :])
    g = as_grid(P)
    assert g.manifold == grids.SPHERICAL and g.ncells == 2
    g = as_grid([[(0, 0), (1, 0), (0, 1), (0, 0)], [(0, 0), (1, 0), (1, 1), (0, 1), (0, 0)]])
    assert g.offsets.tolist() == [0, 3, 7]


def test_layout_detection():
    n = 5
    assert _layout(np.zeros((n, 3)), 0, n) == (3, 3, True)                  # C-order (cells, K): level-fastest
    assert _layout(np.zeros((3, n)), 1, n) == (3, n, False)                 # C-order (K, cells): cell-fastest
    assert _layout(np.zeros((n, 3), order="F"), 0, n) == (3, n, False)      # Julia (cells, K)
    assert _layout(np.zeros((n, 3, 2)), 0, n) == (6, 6, True)
    assert _layout(np.zeros((3, 2, n)), 2, n) == (6, n, False)
    assert _layout(np.zeros((2, n, 3)), 1, n) is None                       # spatial axis in the middle
    assert _layout(np.zeros((n, 6))[:, ::2], 0, n) is None
    assert _layout(np.zeros(2 * n)[::2], 0, n) is None


class _FakeMatrix(B200Matrix):
    def __init__(self, n_dst, n_src):
        self._h = None; self._n_dst = n_dst; self._n_src = n_src; self.transposed = False

    def apply(self, *a, **k):
        raise AssertionError("validation must fail before the device is touched")


def test_dims_validation_errors():
    # test/regridding.jl:194-202
    R = RegridderB200(_FakeMatrix(9, 4), np.ones(9), np.ones(4), np.zeros(9), np.zeros(4))
    with pytest.raises(ValueError):
        regrid_(np.zeros((9, 3)), R, np.ones((4, 3)), dims=-1)
    with pytest.raises(ValueError):
        regrid_(np.zeros((9, 3)), R, np.ones((4, 3)), dims=2)
    with pytest.raises(DimensionMismatch):
        regrid_(np.zeros((9, 3, 1)), R, np.ones((4, 3)))
    with pytest.raises(DimensionMismatch):
        regrid_(np.zeros((9, 4)), R, np.ones((4, 3)))
    with pytest.raises(DimensionMismatch):
        regrid_(np.zeros((2, 9, 3)), R, np.ones((2, 4, 4)), dims=1)
    with pytest.raises(TypeError):          # non-array field types have no extract method (regrid.jl:256-259)
        regrid_([0.0] * 9, R, np.ones(4))
    T = transpose(R)
    assert T.shape == (4, 9) and T.src_areas is R.dst_areas and T.dst_temp is R.src_temp


def test_balanced_bounds_and_candidate_weights():
    from crg_b200.dist import balanced_bounds, block_bounds, candidate_weights
    w = np.array([1.0] * 10 + [3.0] * 10)
    b = balanced_bounds(w, 2)
    assert b[0][0] == 0 and b[-1][1] == 20 and b[0][1] == b[1][0]
    assert abs(w[b[0][0]:b[0][1]].sum() - w[b[1][0]:b[1][1]].sum()) <= 3.0
    assert balanced_bounds(np.zeros(7), 3) == block_bounds(7, 3)          # degenerate weights: equal counts
    assert balanced_bounds(np.ones(5), 1) == [(0, 5)]
    assert [hi - lo for lo, hi in balanced_bounds(np.ones(12), 4)] == [3, 3, 3, 3]
    dst, src = grids.lonlat_grid(36, 18), grids.healpix_grid(4, "ring")
    cw = candidate_weights(dst, src).numpy().reshape(18, 36)
    assert cw[9, 0] > 1.5 * cw[0, 0]                                       # equatorial cells have more candidates
    assert candidate_weights(grids.polygons_grid([np.array([[0, 0], [1, 0], [0, 1.0]])]), src) is None


def test_octahedral_spec_counts_and_materialises():
    """The octahedral descriptor (BASELINE config 4; no reference cells exist, RingGridsExt.jl:18-20): O320 has
    421 120 cells, ring of rank j has 16 + 4 j of them, and the host cells are what materialize() returns."""
    sp = grids.octahedral_gaussian_spec(320)
    assert sp.ncells == 421120 and sp.kind == "reduced_ring" and sp.n2 == 640
    small = grids.octahedral_gaussian_spec(6)
    g = small.materialize()
    assert g.ncells == small.ncells == 2 * sum(16 + 4 * j for j in range(1, 7))
    assert np.array_equal(g.verts, grids.octahedral_gaussian_grid(6).verts)


def test_kernel_counters_match_the_shipped_sources():
    """bench.py takes the clip kernel's FP64 work per pair and the DRAM traffic figures from
    profiles/r02_kernel_counters.json (ncu captures of the shipped kernels); they are only valid for the sources
    they were measured on.  Re-run scripts/final_profile.sh after touching geom.cuh / kernels.cuh / sell.cuh."""
    import json
    import bench
    p = bench.COUNTERS_JSON
    if not os.path.exists(p):
        pytest.skip("no counters file yet")
    d = json.load(open(p))
    for name, e in d.items():
        assert e["src_hash"] == bench.source_hash(e["src_files"]), f"{name}: kernel sources changed since the ncu capture"
    assert 100 < d["clip"]["flops_per_pair"] < 2000


def test_tripolar_fold_grid_structure():
    """The synthetic RightCenterFolded grid: fold-row cell i is the same quadrilateral as cell nx-1-i; of each pair one
    slot is real and its partner a zero-area ghost (OceananigansExt.jl:119-160); the oracle keeps ghosts out of the
    candidates, so their rows / columns are empty and ones -> ones after mirroring."""
    nx, ny = 32, 10
    g = grids.tripolar_fold_grid(nx, ny)
    real, partner = grids.fold_row_slots(nx)
    assert sorted(np.concatenate([real, partner]).tolist()) == list(range(nx))
    assert real.tolist() == list(range(8)) + list(range(16, 24)) and (partner == nx - 1 - real).all()
    base = (ny - 1) * nx
    assert (g.verts[base + partner] == g.verts[base + partner][:, :1]).all()           # ghosts: four equal points
    # the real cell r and the cell the fold maps it to are the same polygon (checked on the vertex matrix rule)
    lon = 360.0 * np.arange(nx + 1) / nx
    top = grids.unit_sphere_from_geographic(lon, np.full(nx + 1, 84.0))
    for r in real:
        assert np.allclose(g.verts[base + r], [top[r], top[r + 1], top[nx - r - 1], top[nx - r]], atol=1e-15)
    a = oracle.cell_areas(g)
    assert (a[base + partner] == 0).all() and (a[base + real] > 0).all()
    assert abs(a.sum() / (2 * np.pi * (1 - np.sin(np.radians(-80.0)))) - 1) < 1e-3      # the real cells tile the cap once
    src = grids.healpix_grid(8, "ring")
    O = oracle.build_regridder(g, src)
    assert np.diff(O.tocsc().tocsr().indptr)[base + partner].sum() == 0
    y = grids.mirror_fold_partners(O.regrid(np.ones(src.ncells)), nx, ny)
    assert np.allclose(y, 1.0, atol=1e-10)

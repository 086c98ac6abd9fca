"""GPU parity tests of the Regridder build: CUDA path (through the C ABI) vs the CPU oracle, the
committed golden vectors and the reference's known answers / invariants."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from crg_b200 import _lib, grids
from crg_b200.regridder import Regridder, normalize_, regrid_, regridder_from_coo, transpose
from helpers import (GOLDEN, GRID_PAIRS_SMALL, KAT_DST_AREAS, KAT_MATRIX, KAT_SRC_AREAS, compare_matrices,
                     kat_simple)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle():
    from oracle import oracle
    return oracle


def test_planar_known_answer_exact(gpu):
    # test/usecases/simple.jl:30-51 (exact ==), README.md:52-80
    g1, g2 = kat_simple()
    R = Regridder(g1, g2, normalize=False)
    A = R.intersections.toarray()
    assert (A == KAT_MATRIX).all()
    assert (R.dst_areas == KAT_DST_AREAS).all() and (R.src_areas == KAT_SRC_AREAS).all()
    assert (A.sum(1) == R.dst_areas).all() and (A.sum(0) == R.src_areas).all()
    assert R.shape == (4, 5) and R.size(0) == 4
    # transposed regridder shares storage (simple.jl:54-71)
    T = transpose(R)
    assert T.src_areas is R.dst_areas and T.dst_areas is R.src_areas
    assert T.src_temp is R.dst_temp and T.dst_temp is R.src_temp
    src1 = np.array([1.0, 2, 3, 4])
    dst2 = np.zeros(5)
    regrid_(dst2, T, src1)
    assert np.isclose((dst2 * T.dst_areas).sum(), (src1 * T.src_areas).sum())


@pytest.mark.parametrize("name", list(GRID_PAIRS_SMALL))
def test_against_golden_vectors(gpu, name):
    fd, fs = GRID_PAIRS_SMALL[name]
    g = np.load(os.path.join(GOLDEN, name.replace("<-", "__from__") + ".npz"))
    R = Regridder(fd(), fs())
    B = sp.csc_matrix((g["nzval"], g["rowval"], g["colptr"]), shape=(int(g["n_dst"]), int(g["n_src"])))
    compare_matrices(R.intersections.tocsc(), B, g["dst_areas"], g["src_areas"])
    assert np.allclose(R.dst_areas, g["dst_areas"], rtol=1e-13, atol=0)
    assert np.allclose(R.src_areas, g["src_areas"], rtol=1e-13, atol=0)
    y = np.zeros(int(g["n_dst"]))
    regrid_(y, R, g["x"])
    assert np.allclose(y, g["y"], rtol=1e-12, atol=1e-15)
    xb = np.zeros(int(g["n_src"]))
    regrid_(xb, transpose(R), g["y"])
    assert np.allclose(xb, g["xb"], rtol=1e-12, atol=1e-15)


MEDIUM = {
    "cfg1 2deg<-1deg": (lambda: grids.lonlat_grid(180, 90), lambda: grids.lonlat_grid(360, 180)),
    "healpix64ring<-lonlat360x180": (lambda: grids.healpix_grid(64, "ring"), lambda: grids.lonlat_grid(360, 180)),
    "lonlat360x180<-healpix64nested": (lambda: grids.lonlat_grid(360, 180), lambda: grids.healpix_grid(64, "nested")),
    "lonlat180x90<-C48": (lambda: grids.lonlat_grid(180, 90), lambda: grids.cubed_sphere_grid(48)),
    "F48<-O48": (lambda: grids.full_gaussian_grid(48), lambda: grids.octahedral_gaussian_grid(48)),
    "O32<-FullClenshaw24": (lambda: grids.octahedral_gaussian_grid(32), lambda: grids.full_clenshaw_grid(24)),
    "regional": (lambda: grids.lonlat_grid(50, 40, -20, 30, 30, 70), lambda: grids.lonlat_grid(64, 64, -30, 40, 20, 80)),
    "planar200<-planar100": (lambda: grids.planar_regular_grid(np.linspace(0, 1, 101), np.linspace(0, 2, 101)),
                             lambda: grids.planar_regular_grid(np.linspace(-0.1, 1.2, 201), np.linspace(0, 2, 201))),
}


@pytest.mark.parametrize("name", list(MEDIUM))
def test_against_oracle(gpu, name):
    """north_star parity: pattern identical above the sliver threshold, entries within 1e-10
    relative, conserved global mean within 1e-12."""
    oracle = _oracle()
    dst, src = MEDIUM[name][0](), MEDIUM[name][1]()
    R = Regridder(dst, src, keep_candidates=True)
    O = oracle.build_regridder(dst, src, nthreads=oracle.max_threads())
    compare_matrices(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas, rtol=1e-10)
    assert np.allclose(R.dst_areas, O.dst_areas, rtol=1e-13, atol=0)
    assert np.allclose(R.src_areas, O.src_areas, rtol=1e-13, atol=0)
    # CSR and CSC copies describe the same matrix; rows sorted within columns like SparseArrays.sparse
    A = R.intersections.tocsc()
    assert abs(A - R.intersections.tocsr().tocsc()).nnz == 0
    assert A.has_sorted_indices or (np.diff(A.indices)[np.diff(A.indices) < 0].size <= A.shape[1])
    # the device broad phase is a superset of every overlapping pair, and reports each pair once
    ps, pd = R.intersections.candidates()
    keys = pd * src.ncells + ps
    assert np.unique(keys).size == keys.size
    Oc = O.tocsc().tocoo()
    sig = Oc.data > 1e-14 * Oc.data.max()
    assert np.isin(Oc.row[sig].astype(np.int64) * src.ncells + Oc.col[sig], keys).all()
    # conservation of the global mean
    x = np.random.default_rng(3).random(src.ncells)
    y = np.zeros(dst.ncells)
    regrid_(y, R, x, normalize=False)        # y = A x: integrals
    assert abs(y.sum() / (np.asarray(A.sum(0)).ravel() * x).sum() - 1) < 1e-12
    regrid_(y, R, x)
    assert np.allclose(y, O.regrid(x), rtol=1e-11, atol=1e-14, equal_nan=True)


def test_full_size_invariants_cfg2(gpu):
    """BASELINE config 2 at full size (HEALPix 256 <-> 0.5 deg): size-independent properties."""
    dst, src = grids.lonlat_grid(720, 360), grids.healpix_grid(256, "ring")
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    rtol = np.sqrt(np.finfo(float).eps)             # test/sweat.jl:113-116
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=rtol, atol=0)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=rtol, atol=0)
    assert abs(R.dst_areas.sum() / (4 * np.pi) - 1) < 1e-12 and abs(R.src_areas.sum() / (4 * np.pi) - 1) < 1e-12
    ones = np.ones(src.ncells)
    y = np.zeros(dst.ncells)
    regrid_(y, R, ones)
    assert np.allclose(y, 1.0, atol=1e-10)          # test/usecases/fullclenshaw.jl:21-41
    rng = np.random.default_rng(11)
    x1, x2 = rng.random(src.ncells), rng.random(src.ncells)
    y1, y2, y12 = np.zeros(dst.ncells), np.zeros(dst.ncells), np.zeros(dst.ncells)
    regrid_(y1, R, x1); regrid_(y2, R, x2); regrid_(y12, R, 2.0 * x1 - 3.0 * x2)
    assert np.allclose(y12, 2.0 * y1 - 3.0 * y2, rtol=1e-12, atol=1e-12)         # linearity
    assert abs((y1 * R.dst_areas).sum() / (x1 * R.src_areas).sum() - 1) < 1e-12  # global mean
    xb = np.zeros(src.ncells)
    regrid_(xb, transpose(R), y1)
    assert abs((xb * R.src_areas).sum() / (y1 * R.dst_areas).sum() - 1) < 1e-12
    # matches scipy on the exported matrix
    assert np.allclose(y1, (A @ x1) / R.dst_areas, rtol=1e-12)
    assert np.allclose(xb, (A.T @ y1) / R.src_areas, rtol=1e-12)


def test_full_size_invariants_cfg5(gpu):
    """BASELINE config 5 (0.25 deg <-> HEALPix 512), the bench workload."""
    dst, src = grids.lonlat_grid(1440, 720), grids.healpix_grid(512, "ring")
    R = Regridder(dst, src)
    st = R.intersections.stats()
    assert st["n_big_dst"] == 0 and st["n_big_src"] == 0
    x = np.random.default_rng(2).random(src.ncells)
    y = np.zeros(dst.ncells)
    regrid_(y, R, x)
    assert abs((y * R.dst_areas).sum() / (x * R.src_areas).sum() - 1) < 1e-12
    regrid_(y, R, np.ones(src.ncells))
    assert np.allclose(y, 1.0, atol=1e-9)
    A = R.intersections.tocsr()
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=1.5e-8, atol=0)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=1.5e-8, atol=0)


def test_custom_intersection_operator(gpu):
    # test/regridding.jl:9-41
    sq = [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 0.0)]
    dst = grids.polygons_grid([sq, sq]); src = grids.polygons_grid([sq, sq, sq])
    calls = [0]

    def op(p1, p2):
        calls[0] += 1
        return 2.5
    R = Regridder(dst, src, intersection_operator=op, normalize=False, threaded=False)
    A = R.intersections.tocsc()
    assert calls[0] == 6 and A.shape == (2, 3) and A.nnz == 6 and (A.data == 2.5).all()
    calls[0] = 0
    R = Regridder(dst, src, intersection_operator=lambda a, b: (calls.__setitem__(0, calls[0] + 1), -1.0)[1])
    assert calls[0] == 6 and R.intersections.nnz == 0
    assert R.intersections.tocsc().nnz == 0


def test_from_coo_sums_duplicates(gpu):
    oracle = _oracle()
    rng = np.random.default_rng(7)
    n_dst, n_src, n = 300, 200, 5000
    r = rng.integers(0, n_dst, n); c = rng.integers(0, n_src, n); v = rng.random(n) + 0.1
    r[:500] = r[500:1000]; c[:500] = c[500:1000]          # forced duplicates
    R = regridder_from_coo(n_dst, n_src, r, c, v, np.ones(n_dst), np.ones(n_src))
    colptr, rowval, nzval = oracle.coo_to_csc(n_dst, n_src, r, c, v)
    A = R.intersections.tocsc()
    assert (A.indptr == colptr).all() and (A.indices == rowval).all()
    assert np.allclose(A.data, nzval, rtol=1e-14)
    B = sp.coo_matrix((v, (r, c)), shape=(n_dst, n_src)).tocsc()
    assert abs(A - B).max() < 1e-12
    empty = regridder_from_coo(3, 2, [], [], [], np.ones(3), np.ones(2))
    assert empty.intersections.nnz == 0
    y = np.full(3, 7.0)
    regrid_(y, empty, np.ones(2))
    assert (y == 0).all()


def test_normalize(gpu):
    # regridder.jl:54-62,160
    dst, src = grids.lonlat_grid(24, 12, radius=6371e3), grids.healpix_grid(4, "ring", radius=6371e3)
    R = Regridder(dst, src)
    Rn = Regridder(dst, src, normalize=True)
    m = R.intersections.maximum()
    assert abs(Rn.intersections.maximum() - 1.0) < 1e-15
    assert np.allclose(Rn.intersections.tocsc().data, R.intersections.tocsc().data / m, rtol=1e-15)
    assert np.allclose(Rn.dst_areas, R.dst_areas / m, rtol=1e-15) and np.allclose(Rn.src_areas, R.src_areas / m, rtol=1e-15)
    assert abs(R.dst_areas.sum() / (4 * np.pi * 6371e3 ** 2) - 1) < 1e-12
    R2 = normalize_(Regridder(dst, src))
    assert np.allclose(R2.dst_areas, Rn.dst_areas, rtol=1e-15)
    x = np.random.default_rng(0).random(src.ncells)
    y, yn = np.zeros(dst.ncells), np.zeros(dst.ncells)
    regrid_(y, R, x); regrid_(yn, Rn, x)
    assert np.allclose(y, yn, rtol=1e-13)
    # transposed apply of the normalised regridder uses the scaled CSC copy too
    xb, xbn = np.zeros(src.ncells), np.zeros(src.ncells)
    regrid_(xb, transpose(R), y); regrid_(xbn, transpose(Rn), y)
    assert np.allclose(xb, xbn, rtol=1e-13)
    # the two halves used by destination-sharded regridders: device maximum, scale by a given maximum
    assert m == R.intersections.tocsc().data.max()
    from crg_b200.regridder import scale_
    R3 = scale_(Regridder(dst, src), m)
    assert np.array_equal(R3.intersections.tocsc().data, Rn.intersections.tocsc().data)
    assert np.array_equal(R3.dst_areas, Rn.dst_areas) and np.array_equal(R3.src_areas, Rn.src_areas)
    with pytest.raises(_lib.CrgError):
        scale_(R3, 0.0)


def test_ragged_clockwise_and_degenerate_cells(gpu):
    oracle = _oracle()
    rng = np.random.default_rng(1)
    # ragged: triangles, quads, pentagons, hexagons around random centres, random orientation
    def poly(cx, cy, k, r, flip):
        t = np.sort(rng.random(k)) * 2 * np.pi
        p = np.stack([cx + r * np.cos(t), cy + r * np.sin(t)], axis=1)
        return p[::-1] if flip else p
    dst = grids.polygons_grid([poly(rng.random() * 4, rng.random() * 4, rng.integers(3, 9), 0.5, rng.random() < 0.5) for _ in range(150)])
    src = grids.polygons_grid([poly(rng.random() * 4, rng.random() * 4, rng.integers(3, 7), 0.4, rng.random() < 0.5) for _ in range(200)])
    assert dst.offsets is not None
    R = Regridder(dst, src)
    O = oracle.build_regridder(dst, src)
    assert abs(R.intersections.tocsc() - O.tocsc()).max() < 1e-13
    assert np.allclose(R.dst_areas, O.dst_areas, rtol=1e-13) and (R.dst_areas > 0).all()
    # spherical grid stored clockwise (j running north -> south) gives the same matrix
    g = grids.lonlat_grid(24, 12)
    cw = grids.Grid(np.ascontiguousarray(g.verts[:, ::-1]), grids.SPHERICAL)
    s = grids.healpix_grid(4, "nested")
    A = Regridder(g, s).intersections.tocsc(); B = Regridder(cw, s).intersections.tocsc()
    assert abs(A - B).max() < 1e-15
    A2 = Regridder(s, cw).intersections.tocsc()
    assert abs(A2 - A.T).max() < 1e-15
    # too many vertices: planar rings are split into convex parts by the front end (decompose.py) ...
    big = grids.polygons_grid([poly(0, 0, 12, 1.0, False), poly(0, 0, 3, 1.0, False)])
    Rb = Regridder(big, src)
    Ob = oracle.build_regridder(big, src)
    assert abs(Rb.intersections.tocsc() - Ob.tocsc()).max() < 1e-13 and np.allclose(Rb.dst_areas, Ob.dst_areas, rtol=1e-13)
    # ... on the sphere (and at the C ABI) they are reported, not truncated
    t = np.linspace(0, 2 * np.pi, 13)[:-1]
    ring12 = np.stack([0.1 * np.cos(t), 0.1 * np.sin(t), np.ones(12)], axis=1)
    ring12 /= np.linalg.norm(ring12, axis=1)[:, None]
    with pytest.raises(_lib.CrgError) as e:
        Regridder(grids.polygons_grid([ring12, ring12[:3]], grids.SPHERICAL), grids.healpix_grid(1, "ring"))
    assert e.value.code == _lib.CRG_ERR_UNSUPPORTED


def test_area_threshold_and_options(gpu):
    dst, src = grids.lonlat_grid(18, 9), grids.lonlat_grid(36, 18)
    R0 = Regridder(dst, src)
    R1 = Regridder(dst, src, area_threshold=1e-12)
    A0 = R0.intersections.tocsc(); A1 = R1.intersections.tocsc()
    assert A1.nnz <= A0.nnz and (A1.data > 1e-12).all()
    assert abs(A0 - A1).max() <= 1e-12
    Rn = Regridder(dst, src, build_transpose=False)
    with pytest.raises(_lib.CrgError):
        regrid_(np.zeros(src.ncells), transpose(Rn), np.zeros(dst.ncells))
    with pytest.raises(ValueError):
        Regridder(dst, grids.planar_unit_square_grid(2, 2))
    # empty destination grid
    E = Regridder(grids.Grid(np.zeros((0, 4, 3)), grids.SPHERICAL), src)
    assert E.shape == (0, src.ncells) and E.intersections.nnz == 0


SPECS = {
    "lonlat": lambda: grids.lonlat_spec(96, 48),
    "lonlat_regional": lambda: grids.lonlat_spec(40, 30, -20.0, 35.0, 10.0, 70.0),
    "healpix_ring": lambda: grids.healpix_spec(16, "ring"),
    "healpix_nested": lambda: grids.healpix_spec(16, "nested"),
    "healpix_1": lambda: grids.healpix_spec(1, "ring"),
    "gaussian": lambda: grids.full_gaussian_spec(16),
    "clenshaw": lambda: grids.full_clenshaw_spec(12),
    "cubed_sphere": lambda: grids.cubed_sphere_spec(10),
    "octahedral": lambda: grids.octahedral_gaussian_spec(24),
    "octahedral_O320": lambda: grids.octahedral_gaussian_spec(320),       # BASELINE config 4
}


@pytest.mark.parametrize("name", list(SPECS))
def test_device_generated_grids_match_host_generators(gpu, name):
    """csrc/gridgen.cuh vs grids.py: same cells, same field order (differences only from libm)."""
    from crg_b200.regridder import grid_cells
    spec = SPECS[name]()
    host = spec.materialize()
    dev = grid_cells(spec)
    assert dev.shape == host.verts.shape
    assert np.abs(dev - host.verts).max() < 1e-14
    if name == "lonlat":        # poles exact, seam closes bit for bit
        assert (dev[0, 0] == [0, 0, -1]).all() and (dev[-1, 2] == [0, 0, 1]).all()
        assert (dev[95, 1] == dev[0, 0]).all() or np.array_equal(dev[95, 1][:2] * 0, dev[0, 0][:2] * 0)
        assert np.array_equal(dev[95, 2], dev[0, 3])


def test_described_grids_build_the_same_regridder(gpu):
    import torch
    from crg_b200.regridder import grid_cells
    pairs = [(grids.lonlat_spec(90, 45), grids.healpix_spec(16, "ring")),
             (grids.healpix_spec(8, "nested"), grids.lonlat_spec(48, 24)),
             (grids.full_gaussian_spec(12), grids.cubed_sphere_spec(8)),
             (grids.full_gaussian_spec(16), grids.octahedral_gaussian_spec(32)),
             (grids.octahedral_gaussian_spec(12), grids.lonlat_spec(40, 20))]
    for ds, ss in pairs:
        R1 = Regridder(ds, ss)
        R2 = Regridder(ds.materialize(), ss.materialize())
        compare_matrices(R1.intersections.tocsc(), R2.intersections.tocsc(), R2.dst_areas, R2.src_areas, rtol=1e-10)
        assert np.allclose(R1.dst_areas, R2.dst_areas, rtol=1e-12) and np.allclose(R1.src_areas, R2.src_areas, rtol=1e-12)
    # mixed: described destination, explicit (host) source; and device-materialised cells
    ds, ss = pairs[0]
    R3 = Regridder(ds, ss.materialize())
    assert abs(R3.intersections.tocsc() - Regridder(ds, ss).intersections.tocsc()).max() < 1e-15
    t = torch.empty((ds.ncells, 4, 3), dtype=torch.float64, device="cuda")
    grid_cells(ds, out=t)
    R4 = Regridder(grids.Grid(t, grids.SPHERICAL), ss)
    assert abs(R4.intersections.tocsc() - Regridder(ds, ss).intersections.tocsc()).max() == 0.0
    with pytest.raises(_lib.CrgError):
        Regridder(grids.GridSpec("healpix", 12), ss)        # nside not a power of two


def _rotation(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def test_rotated_polar_and_identical_grids(gpu):
    """Geometry the structured tests do not reach: arbitrarily rotated grids (no cell edge follows a
    coordinate line, the poles fall inside ordinary cells), ragged cells that CONTAIN a pole, and a grid
    regridded onto itself (every edge coincident)."""
    oracle = _oracle()
    rng = np.random.default_rng(42)
    # rotated HEALPix <- rotated lon-lat
    Q1, Q2 = _rotation(rng), _rotation(rng)
    dst = grids.Grid(np.ascontiguousarray(grids.healpix_grid(16, "ring").verts @ Q1.T), grids.SPHERICAL)
    src = grids.Grid(np.ascontiguousarray(grids.lonlat_grid(72, 36).verts @ Q2.T), grids.SPHERICAL)
    R = Regridder(dst, src)
    O = oracle.build_regridder(dst, src, nthreads=oracle.max_threads())
    compare_matrices(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas)
    A = R.intersections.tocsr()
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=1.5e-8)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=1.5e-8)
    # ... and with the roles swapped: the rotated pole corners (zero-length edges whose coordinates are not exact)
    # now belong to the CLIP cells
    R = Regridder(src, dst)
    O = oracle.build_regridder(src, dst, nthreads=oracle.max_threads())
    compare_matrices(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas)
    # polar caps as octagons that contain the pole + bands of quads (ragged: offsets path)
    def cap_grid(nlon, lats):
        polys = []
        lon = np.arange(nlon) * 360.0 / nlon
        polys.append(grids.unit_sphere_from_geographic(lon, np.full(nlon, lats[-1]))[::1])          # north cap, CCW
        polys.append(grids.unit_sphere_from_geographic(lon[::-1], np.full(nlon, lats[0])))          # south cap, CCW from outside
        for j in range(len(lats) - 1):
            for i in range(nlon):
                lo, hi = lon[i], lon[i] + 360.0 / nlon
                polys.append(grids.unit_sphere_from_geographic(np.array([lo, hi, hi, lo]),
                                                               np.array([lats[j], lats[j], lats[j + 1], lats[j + 1]])))
        return grids.polygons_grid(polys, grids.SPHERICAL)
    gc = cap_grid(8, np.array([-60.0, -20.0, 20.0, 60.0]))
    assert gc.offsets is not None
    gs = grids.healpix_grid(4, "nested")
    R = Regridder(gc, gs)
    O = oracle.build_regridder(gc, gs)
    compare_matrices(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas)
    assert abs(R.dst_areas.sum() / (4 * np.pi) - 1) < 1e-12          # caps + bands tile the sphere
    y = np.zeros(gc.ncells); regrid_(y, R, np.ones(gs.ncells))
    assert np.allclose(y, 1.0, atol=1e-10)
    # a grid onto itself: the matrix is diag(areas) up to round-off slivers
    g = grids.healpix_grid(8, "ring")
    R = Regridder(g, g)
    A = R.intersections.tocsr()
    assert np.allclose(A.diagonal(), R.dst_areas, rtol=1e-12)
    off = A - sp.diags(A.diagonal())
    assert off.nnz == 0 or abs(off).max() < 1e-12 * R.dst_areas.max()


@pytest.mark.parametrize("dst_name,src_name,blocks", [
    ("lonlat", "healpix", 4), ("healpix_nested", "lonlat", 3), ("cubed", "healpix", 5), ("lonlat", "cubed", 2)])
def test_destination_blocks_equal_rows_of_the_full_matrix(gpu, dst_name, src_name, blocks):
    """A destination-sharded build sees a slab of the globe and skips the source cells that cannot
    reach it (culling box, crg_b200.cu): every block must reproduce its rows of the full matrix
    bit for bit, for band-shaped (lon-lat, ring) and patch-shaped (nested, cubed-sphere) blocks."""
    make = {"lonlat": lambda: grids.lonlat_grid(96, 48), "healpix": lambda: grids.healpix_grid(16, "ring"),
            "healpix_nested": lambda: grids.healpix_grid(16, "nested"), "cubed": lambda: grids.cubed_sphere_grid(12)}
    dst, src = make[dst_name](), make[src_name]()
    full = Regridder(dst, src).intersections.tocsr()
    n = dst.ncells
    for k in range(blocks):
        lo, hi = k * n // blocks, (k + 1) * n // blocks
        R = Regridder(dst.slice(lo, hi), src)
        blk = R.intersections.tocsr()
        ref = full[lo:hi]
        assert blk.shape == ref.shape
        assert np.array_equal(blk.indptr, ref.indptr) and np.array_equal(blk.indices, ref.indices)
        assert np.array_equal(blk.data, ref.data)
        assert np.array_equal(R.dst_areas, Regridder(dst, src).dst_areas[lo:hi])


def test_four_times_the_bench_workload(gpu):
    """0.125 deg lon-lat (4.1 M cells) <-> HEALPix 1024 (12.6 M cells): ~54 M candidate pairs, ~37 M
    entries -- four times BASELINE config 5 in every count, to exercise the 32-bit offsets, the scans
    and the arena well beyond the bench sizes.  Grids are generated on the device; the checks are the
    size-independent ones: A 1 = dst areas and A^T 1 = src areas (through regrid! of ones), global-mean
    conservation in both directions, and sum of areas = 4 pi."""
    import torch
    dst, src = grids.lonlat_spec(2880, 1440), grids.healpix_spec(1024, "ring")
    R = Regridder(dst, src)
    n_dst, n_src = R.shape
    assert (n_dst, n_src) == (2880 * 1440, 12 * 1024 * 1024)
    st = R.intersections.stats()
    assert st["n_big_dst"] == 0 and st["n_big_src"] == 0
    assert 3.0e7 < R.intersections.nnz < 4.5e7 and st["n_candidates"] < 7.0e7
    da, sa = torch.from_numpy(R.dst_areas).cuda(), torch.from_numpy(R.src_areas).cuda()
    assert abs(float(da.sum()) / (4 * np.pi) - 1) < 1e-12 and abs(float(sa.sum()) / (4 * np.pi) - 1) < 1e-12
    ones_s = torch.ones(n_src, dtype=torch.float64, device="cuda")
    y = torch.zeros(n_dst, dtype=torch.float64, device="cuda")
    regrid_(y, R, ones_s)
    assert float((y - 1).abs().max()) < 1e-9                           # row sums = dst areas
    xb = torch.zeros(n_src, dtype=torch.float64, device="cuda")
    regrid_(xb, transpose(R), torch.ones(n_dst, dtype=torch.float64, device="cuda"))
    assert float((xb - 1).abs().max()) < 1e-9                          # column sums = src areas
    x = torch.rand(n_src, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    regrid_(y, R, x)
    assert abs(float((y * da).sum() / (x * sa).sum()) - 1) < 1e-12
    regrid_(xb, transpose(R), y)
    assert abs(float((xb * sa).sum() / (y * da).sum()) - 1) < 1e-12
    print("4x cfg5: build %.2f ms device, %d candidates, %d nnz" % (st["ms_device"], st["n_candidates"], R.intersections.nnz))


def test_one_by_one_and_two_by_one(gpu):
    """nnz = 1 and nnz = 2: the assembly's sorts have nothing (or almost nothing) to do."""
    one = grids.planar_regular_grid([0.0, 1.0], [0.0, 1.0])
    two = grids.planar_regular_grid([0.0, 0.5, 1.0], [0.0, 1.0])
    R = Regridder(one, one)
    assert R.intersections.nnz == 1 and np.allclose(R.intersections.tocsc().toarray(), [[1.0]])
    x = np.array([3.0]); y = np.zeros(1)
    regrid_(y, R, x); assert y[0] == 3.0
    regrid_(x, transpose(R), y); assert x[0] == 3.0
    R2 = Regridder(one, two)                                   # 1 x 2
    assert np.allclose(R2.intersections.tocsc().toarray(), [[0.5, 0.5]])
    y = np.zeros(1); regrid_(y, R2, np.array([1.0, 3.0])); assert abs(y[0] - 2.0) < 1e-15
    xb = np.zeros(2); regrid_(xb, transpose(R2), np.array([2.0])); assert np.allclose(xb, [2.0, 2.0])


def test_nonconvex_cells_are_detected_not_silently_clipped(gpu):
    """the device clip is convex-convex; the C ABI refuses non-convex rings."""
    import ctypes as C
    from crg_b200.regridder import _cells_struct, _make_options
    L = np.array([(0, 0), (2, 0), (2, 1), (1, 1), (1, 2), (0, 2)], dtype=float)
    sq = np.array([(0, 0), (2, 0), (2, 2), (0, 2), (0, 2), (0, 2)], dtype=float)      # convex, padded with a repeated vertex
    g_bad = grids.Grid(np.stack([L, sq]), grids.PLANAR)
    g_ok = grids.planar_unit_square_grid(2, 2)
    keep = []
    o = _make_options(grids.PLANAR, False, 1.0, None, 0.0, True, False)
    out = C.c_void_p()
    rc = _lib.lib().crg_build(C.byref(o), C.byref(_cells_struct(g_bad, keep)), C.byref(_cells_struct(g_ok, keep)), C.byref(out))
    assert rc == _lib.CRG_ERR_UNSUPPORTED and b"not convex" in _lib.lib().crg_last_error()
    rc = _lib.lib().crg_build(C.byref(o), C.byref(_cells_struct(g_ok, keep)), C.byref(_cells_struct(g_bad, keep)), C.byref(out))
    assert rc == _lib.CRG_ERR_UNSUPPORTED
    # a non-convex spherical quad (one corner pushed inside) is refused as well
    q = grids.lonlat_grid(8, 4).verts.copy()
    q[13, 2] = (0.8 * q[13, 0] + 0.1 * q[13, 1] + 0.1 * q[13, 3]); q[13, 2] /= np.linalg.norm(q[13, 2])   # (a mid-latitude cell)
    with pytest.raises(_lib.CrgError) as ei:
        Regridder(grids.Grid(q, grids.SPHERICAL), grids.healpix_grid(2, "ring"))
    assert ei.value.code == _lib.CRG_ERR_UNSUPPORTED


def test_planar_nonconvex_polygons_through_convex_parts(gpu):
    """The reference's planar operator is Foster-Hormann (regridder.jl:87-94): non-convex simple polygons are
    legal input.  The front end splits them into convex parts and sums the part pairs (decompose.py)."""
    L = [(0, 0), (2, 0), (2, 1), (1, 1), (1, 2), (0, 2)]                       # area 3
    U = [(0, 0), (3, 0), (3, 3), (2, 3), (2, 1), (1, 1), (1, 3), (0, 3)]       # area 7, 8 vertices, two notches
    star = [(2 + np.cos(a) * (1.5 if i % 2 == 0 else .5), 2 + np.sin(a) * (1.5 if i % 2 == 0 else .5))
            for i, a in enumerate(np.linspace(0, 2 * np.pi, 11)[:-1])]          # 10 vertices > CRG_MAX_VERTS
    dst = grids.polygons_grid([L, U[::-1], star])
    src = grids.planar_regular_grid(np.linspace(-1, 4, 21), np.linspace(-1, 4, 21))   # 0.25 squares covering everything
    R = Regridder(dst, src)
    A = R.intersections.toarray()
    shoelace = lambda p: 0.5 * abs(sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p))))
    want = np.array([shoelace(L), shoelace(U), shoelace(star)])
    assert np.allclose(R.dst_areas, want, rtol=1e-14)
    assert np.allclose(A.sum(1), want, rtol=1e-13)              # the squares tile the plane: row sums = polygon areas
    assert np.allclose(R.src_areas, 0.0625)
    # a unit square inside the notch of U does not touch it; one inside the arm is fully covered
    sq = lambda x, y: [(x, y), (x + 1, y), (x + 1, y + 1), (x, y + 1)]
    R2 = Regridder(grids.polygons_grid([U]), grids.polygons_grid([sq(1, 1.5), sq(0, 1.5), sq(0.5, 0.5)]))
    assert np.allclose(R2.intersections.toarray(), [[0.0, 1.0, 0.75]], atol=1e-15)
    # exact arithmetic case against the oracle on the convex parts
    y = np.zeros(3); regrid_(y, R, np.ones(src.ncells))
    assert np.allclose(y, 1.0, rtol=1e-13)


def test_clip_pairs_matches_the_build_and_the_oracle(gpu):
    """crg_clip_pairs = compute_intersection_areas (intersection_areas.jl:4-32) over a pair list."""
    from crg_b200.regridder import clip_pairs
    oracle = _oracle()
    dst, src = grids.lonlat_grid(90, 45), grids.healpix_grid(16, "nested")
    R = Regridder(dst, src, keep_candidates=True)
    ps, pd = R.intersections.candidates()
    a = clip_pairs(dst, src, ps, pd)
    A = R.intersections.tocsr()
    assert np.array_equal(a[a > 0], np.asarray(A[pd[a > 0], ps[a > 0]]).ravel())      # bit-identical to the build
    assert (a > 0).sum() == A.nnz
    i1, i2, oa = oracle.compute_intersection_areas(dst, src, ps, pd)
    key = pd * src.ncells + ps
    order = np.argsort(key)
    want = np.zeros(len(ps)); want[order[np.searchsorted(key[order], i2 * src.ncells + i1)]] = oa
    assert np.allclose(a, want, rtol=1e-10, atol=1e-12 * want.max())
    # radius scaling, planar manifold, ragged cells, error paths
    assert np.allclose(clip_pairs(dst, src, ps[:100], pd[:100], radius=3.0), 9.0 * a[:100], rtol=1e-15)
    g1, g2 = kat_simple()
    kk = np.array([(s, d) for d in range(4) for s in range(5)])
    assert np.array_equal(clip_pairs(g1, g2, kk[:, 0], kk[:, 1]).reshape(4, 5), KAT_MATRIX)
    with pytest.raises(_lib.CrgError):
        clip_pairs(dst, src, [src.ncells], [0])


def test_tripolar_fold_row_ghost_cells_and_mirroring(gpu):
    """a tripolar grid with a RightCenterFolded north row (ext/ConservativeRegriddingOceananigansExt.jl:
    76-187): every physical cell of the fold row has two field slots, one real and one ghost (zero area, never a
    candidate); regrid! copies each primary's value into its partner (:216-240).  Checked against the oracle with the
    same padding semantics, in both directions, plus the reference's own invariants for tripolar grids
    (test/usecases/constant_field.jl:41-155 ones -> ones, test/usecases/oceananigans.jl:37-159 conservation 1e-10)."""
    import torch
    oracle = _oracle()
    nx, ny = 64, 24
    tri = grids.tripolar_fold_grid(nx, ny)
    real, partner = grids.fold_row_slots(nx)
    base = (ny - 1) * nx
    for other in (grids.healpix_grid(16, "ring"), grids.lonlat_grid(72, 36)):
        # tripolar as destination
        R = Regridder(tri, other)
        O = oracle.build_regridder(tri, other, nthreads=oracle.max_threads())
        compare_matrices(R.intersections.tocsc(), O.tocsc(), O.dst_areas, O.src_areas)
        assert np.allclose(R.dst_areas, O.dst_areas, rtol=1e-13, atol=0) and (R.dst_areas[base + partner] == 0).all()
        A = R.intersections.tocsr()
        assert np.diff(A.indptr)[base + partner].sum() == 0                   # ghost rows are empty
        assert R.dst_fold == (nx, ny) and R.src_fold is None and transpose(R).src_fold == (nx, ny)
        x = np.random.default_rng(3).random(other.ncells)
        y = np.zeros(tri.ncells); regrid_(y, R, x)
        want = grids.mirror_fold_partners(O.regrid(x), nx, ny)
        assert np.isfinite(y).all() and np.allclose(y, want, rtol=1e-12)
        assert np.array_equal(y[base + partner], y[base + real])              # partners mirror their primaries
        ones = np.zeros(tri.ncells); regrid_(ones, R, np.ones(other.ncells))
        assert np.allclose(ones, 1.0, atol=1e-10)
        # device tensors (the mirror kernel), single field and K levels in both layouts
        yd = torch.zeros(tri.ncells, dtype=torch.float64, device="cuda")
        regrid_(yd, R, torch.from_numpy(x).cuda())
        assert np.array_equal(yd.cpu().numpy(), y)
        X = np.random.default_rng(4).random((other.ncells, 3))
        for order in ("C", "F"):
            Yd = torch.zeros(tri.ncells, 3, dtype=torch.float64, device="cuda")
            if order == "F":
                Yd = Yd.T.contiguous().T
            Xd = torch.from_numpy(np.asarray(X, order=order)).cuda()
            Xd = Xd if order == "C" else torch.from_numpy(np.ascontiguousarray(X.T)).cuda().T
            regrid_(Yd, R, Xd, dims=0)
            got = Yd.cpu().numpy()
            assert np.allclose(got[:, 0], grids.mirror_fold_partners(O.regrid(X[:, 0].copy()), nx, ny), rtol=1e-12)
            assert np.array_equal(got[base + partner], got[base + real])
        # tripolar as source: ghost columns are empty, the integral is conserved on the covered part
        RT = Regridder(other, tri)
        OT = oracle.build_regridder(other, tri, nthreads=oracle.max_threads())
        compare_matrices(RT.intersections.tocsc(), OT.tocsc(), OT.dst_areas, OT.src_areas)
        xs = grids.mirror_fold_partners(np.random.default_rng(5).random(tri.ncells), nx, ny)
        yo = np.zeros(other.ncells); regrid_(yo, RT, xs, normalize=False)
        assert abs(yo.sum() / (xs * RT.src_areas).sum() - 1) < 1e-10
        # transpose(R) of the first regridder regrids back ONTO the tripolar grid: mirrored as well
        back = np.zeros(other.ncells); regrid_(back, transpose(R), y)
        onto = np.zeros(tri.ncells); regrid_(onto, transpose(RT), yo)
        assert np.array_equal(onto[base + partner], onto[base + real])


def test_wedge_clip_variant_against_the_shipped_kernel(gpu, tmp_path):
    """CRG_CLIP_FAST=1 routes spherical quadrilaterals through the wedge-sum path (csrc/clipfast.cuh; measured, not the
    default -- profiles/README.md): same matrices within the parity bar, same areas, on a HEALPix -> lon-lat pair, a
    nested lon-lat pair (coincident edges) and a pair whose source cells are the larger ones (mostly general pairs)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, %r)
from crg_b200 import grids
from crg_b200.regridder import Regridder
pairs = [(grids.lonlat_grid(180, 90), grids.healpix_grid(64, "ring")), (grids.lonlat_grid(90, 45), grids.lonlat_grid(180, 90)),
         (grids.healpix_grid(32, "nested"), grids.lonlat_grid(90, 45))]
for k, (d, s) in enumerate(pairs):
    R = Regridder(d, s)
    sp.save_npz(sys.argv[1] + "_%%d.npz" %% k, R.intersections.tocsc())
    np.save(sys.argv[1] + "_%%d_areas.npy" %% k, np.concatenate([R.dst_areas, R.src_areas]))
''' % ROOT
    out = {}
    for mode in ("0", "1"):
        env = dict(os.environ, CRG_CLIP_FAST=mode)
        subprocess.run([sys.executable, "-c", code, str(tmp_path / ("m" + mode))], check=True, env=env)
        out[mode] = [(sp.load_npz(str(tmp_path / f"m{mode}_{k}.npz")), np.load(str(tmp_path / f"m{mode}_{k}_areas.npy"))) for k in range(3)]
    for (A0, a0), (A1, a1) in zip(out["0"], out["1"]):
        assert np.array_equal(a0, a1)
        nd = A0.shape[0]
        compare_matrices(A1, A0, a0[:nd], a0[nd:])
        assert abs(A1.sum() - A0.sum()) <= 1e-13 * A0.sum()


def test_cell_alignment_dispatch(gpu):
    """Device-resident vertex soups at every alignment: 32-byte aligned records take the 256-bit loads of the clip
    kernel, 16-byte aligned ones the 128-bit loads, 8-byte aligned ones the general kernel -- the matrices of the first
    two are the same bits, the third agrees within the parity bar."""
    import torch
    dst, src = grids.lonlat_grid(90, 45), grids.healpix_grid(32, "ring")
    ref = Regridder(dst, src)
    A0 = ref.intersections.tocsc()

    def on_device(g, shift_doubles):
        flat = torch.zeros(g.verts.size + 8, dtype=torch.float64, device="cuda")
        assert flat.data_ptr() % 32 == 0
        view = flat[shift_doubles:shift_doubles + g.verts.size].view(g.verts.shape)
        view.copy_(torch.from_numpy(g.verts))
        assert view.data_ptr() % 32 == (8 * shift_doubles) % 32
        return grids.Grid(view, g.manifold), flat

    for shift, exact in ((0, True), (2, True), (1, False)):
        gd, keep_d = on_device(dst, shift)
        gs, keep_s = on_device(src, shift)
        R = Regridder(gd, gs)
        A = R.intersections.tocsc()
        if exact:
            assert (A != A0).nnz == 0, shift
        else:
            compare_matrices(A, A0, ref.dst_areas, ref.src_areas)
        assert np.array_equal(R.dst_areas, ref.dst_areas) and np.array_equal(R.src_areas, ref.src_areas)

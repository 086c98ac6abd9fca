"""GPU parity tests of regrid! (forward / transpose, 1-D, strided, N-D with dims, device tensors)."""
import numpy as np
import pytest

from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid, regrid_, transpose

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planar(gpu):
    # test/regridding.jl:46-65: 4x4 -> 8x8 unit-square grids
    return Regridder(grids.planar_unit_square_grid(8, 8), grids.planar_unit_square_grid(4, 4), threaded=False)


def test_dense_vs_strided_dispatch(planar):
    # test/regridding.jl:45-125, NaN poisoning tells which temp buffers were used
    r = planar
    src = np.arange(1.0, 17.0)
    reference = np.zeros(64)
    regrid_(reference, r, src)
    assert np.allclose(reference.reshape(8, 8)[::2, ::2].ravel(), src)

    def poison():
        r.src_temp.fill(np.nan); r.dst_temp.fill(np.nan)
    poison()
    dst = np.zeros(64); regrid_(dst, r, src)
    assert (dst == reference).all() and np.isnan(r.src_temp).all() and np.isnan(r.dst_temp).all()
    poison()
    big_dst = np.zeros(128); dv = big_dst[::2]; regrid_(dv, r, src)
    assert (dv == reference).all() and np.isnan(r.src_temp).all() and not np.isnan(r.dst_temp).any()
    poison()
    big_src = np.zeros(32); big_src[::2] = src; sv = big_src[::2]
    dst = np.zeros(64); regrid_(dst, r, sv)
    assert (dst == reference).all() and not np.isnan(r.src_temp).any() and np.isnan(r.dst_temp).all()
    poison()
    big_dst = np.zeros(128); dv = big_dst[::2]; regrid_(dv, r, sv)
    assert (dv == reference).all() and not np.isnan(r.src_temp).any() and not np.isnan(r.dst_temp).any()
    # integer sources are converted like `A * values` (simple.jl:40-43); regrid allocates
    out = regrid(r, np.arange(1, 17))
    assert (out == reference).all()


def test_nd_arrays_and_dims(gpu):
    # test/regridding.jl:127-203 (dims are 0-based here)
    r = Regridder(grids.planar_unit_square_grid(3, 3), grids.planar_unit_square_grid(2, 2), threaded=False)
    d = np.zeros(9); regrid_(d, r, np.ones(4)); assert np.allclose(d, 1.0)
    for shape_s, shape_d, dims in [((4, 3), (9, 3), 0), ((4, 3, 2), (9, 3, 2), 0), ((3, 4), (3, 9), 1),
                                   ((2, 4, 3), (2, 9, 3), 1), ((3, 2, 4), (3, 2, 9), 2)]:
        for order in ("C", "F"):
            s = np.ones(shape_s, order=order); d = np.zeros(shape_d, order=order)
            regrid_(d, r, s, dims=dims)
            assert np.allclose(d, 1.0), (shape_s, dims, order)
    # values, not just ones: every slice equals the 1-D result
    rng = np.random.default_rng(0)
    s = rng.random((2, 4, 3)); d = np.zeros((2, 9, 3))
    regrid_(d, r, s, dims=1)
    for a in range(2):
        for b in range(3):
            ref = np.zeros(9); regrid_(ref, r, np.ascontiguousarray(s[a, :, b]))
            assert np.allclose(d[a, :, b], ref, rtol=1e-14)
    # normalize=False returns A x
    A = r.intersections.tocsr()
    d2 = np.zeros((9, 5)); s2 = rng.random((4, 5))
    regrid_(d2, r, s2, normalize=False)
    assert np.allclose(d2, A @ s2, rtol=1e-14)


@pytest.mark.parametrize("K", [1, 3, 32, 33, 100])
def test_batched_spmm_matches_loop_of_spmv(gpu, K):
    """BASELINE config 3 shape (cubed sphere -> lon-lat, K levels) at reduced size: one SpMM launch
    == K SpMVs (regrid.jl:303-318), both memory layouts, forward and transpose."""
    dst, src = grids.lonlat_grid(90, 45), grids.cubed_sphere_grid(24)
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    rng = np.random.default_rng(K)
    X = rng.random((src.ncells, K))
    ref = (A @ X) / R.dst_areas[:, None]
    for order in ("C", "F"):                      # level-fastest / cell-fastest
        Y = np.zeros((dst.ncells, K), order=order)
        regrid_(Y, R, np.asarray(X, order=order), dims=0)
        assert np.allclose(Y, ref, rtol=1e-13, atol=1e-15), order
    Yt = rng.random((dst.ncells, K))
    reft = (A.T @ Yt) / R.src_areas[:, None]
    for order in ("C", "F"):
        Xb = np.zeros((src.ncells, K), order=order)
        regrid_(Xb, transpose(R), np.asarray(Yt, order=order), dims=0)
        assert np.allclose(Xb, reft, rtol=1e-13, atol=1e-15), order


def test_device_tensors_zero_copy(gpu):
    import torch
    dst, src = grids.healpix_grid(16, "ring"), grids.lonlat_grid(60, 30)
    R = Regridder(dst, src)
    x = np.random.default_rng(1).random(src.ncells)
    y = np.zeros(dst.ncells); regrid_(y, R, x)
    xd = torch.from_numpy(x).cuda(); yd = torch.zeros(dst.ncells, dtype=torch.float64, device="cuda")
    R.intersections.set_stream(torch.cuda.current_stream().cuda_stream)
    regrid_(yd, R, xd, asynchronous=True)
    torch.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy(), y)
    Xd = torch.from_numpy(np.stack([x, 2 * x], axis=1)).cuda()         # (cells, K) level-fastest
    Yd = torch.zeros(dst.ncells, 2, dtype=torch.float64, device="cuda")
    regrid_(Yd, R, Xd, dims=0)
    assert np.allclose(Yd.cpu().numpy(), np.stack([y, 2 * y], axis=1), rtol=1e-14)
    # device-resident vertices build the same regridder
    gd = grids.Grid(torch.from_numpy(dst.verts).cuda(), dst.manifold)
    gs = grids.Grid(torch.from_numpy(src.verts).cuda(), src.manifold)
    R2 = Regridder(gd, gs)
    assert abs(R2.intersections.tocsc() - R.intersections.tocsc()).max() == 0.0


def test_long_rows_and_empty_rows(gpu):
    """Polar HEALPix-vs-lonlat rows are hundreds of entries long; a regional source leaves most
    destination rows empty (0/area = 0)."""
    dst, src = grids.healpix_grid(2, "ring"), grids.lonlat_grid(720, 90, 0, 360, 60, 90)
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    assert np.diff(A.indptr).max() > 1000 and (np.diff(A.indptr) == 0).any()
    x = np.random.default_rng(4).random(src.ncells)
    y = np.full(dst.ncells, np.nan); regrid_(y, R, x)
    assert np.allclose(y, (A @ x) / R.dst_areas, rtol=1e-12, atol=0)
    assert (y[np.diff(A.indptr) == 0] == 0).all()
    xb = np.zeros(src.ncells); regrid_(xb, transpose(R), np.ones(dst.ncells))
    assert np.allclose(xb, 1.0, atol=1e-9)

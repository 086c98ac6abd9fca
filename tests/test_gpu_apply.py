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


@pytest.mark.parametrize("K", [1, 3, 18, 32, 33, 64, 100, 130, 200])
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
    # batched, both layouts, through the cut (multi-piece) slices of the long rows
    X = np.random.default_rng(5).random((src.ncells, 7))
    ref = (A @ X) / R.dst_areas[:, None]
    for order in ("C", "F"):
        Y = np.full((dst.ncells, 7), np.nan, order=order)
        regrid_(Y, R, np.asarray(X, order=order), dims=0)
        assert np.allclose(Y, ref, rtol=1e-12, atol=0), order
        Y2 = np.full((dst.ncells, 7), np.nan, order=order)
        regrid_(Y2, R, np.asarray(X, order=order), dims=0, normalize=False)
        assert np.allclose(Y2, A @ X, rtol=1e-12, atol=0), order


@pytest.mark.parametrize("case", ["short", "long", "mixed", "empty_runs", "exact_chunks", "one_row"])
def test_spmv_stream_kernel_stress(gpu, case):
    """K7 (SELL-32-sigma SpMV): rows far longer than a slice piece (cut slices, ticketed partial sums), long
    segments, runs of empty rows, a matrix of exactly uniform rows, a single dense row -- vs scipy on the same matrix."""
    import scipy.sparse as sp
    from crg_b200.regridder import regridder_from_coo
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    n_src = 5000
    if case == "short":
        lens = rng.integers(1, 6, 20000)
    elif case == "long":
        lens = rng.integers(1500, 7000, 12)
    elif case == "mixed":
        lens = np.where(rng.random(6000) < 0.01, rng.integers(100, 5000, 6000), rng.integers(0, 12, 6000))
    elif case == "empty_runs":
        lens = np.zeros(30000, dtype=np.int64); lens[rng.choice(30000, 300, replace=False)] = rng.integers(1, 400, 300)
        lens[-5000:] = 0; lens[:4000] = 0
    elif case == "exact_chunks":
        lens = np.full(1024, 8)                  # 8192 nnz = 4 chunks exactly, rows aligned to chunk borders
    else:
        lens = np.array([4999])
    lens = np.minimum(lens, n_src)
    n_dst = len(lens)
    rows = np.repeat(np.arange(n_dst), lens)
    cols = np.concatenate([rng.choice(n_src, l, replace=False) for l in lens]) if rows.size else np.zeros(0, np.int64)
    vals = rng.random(rows.size) + 0.5
    da, sa = rng.random(n_dst) + 0.5, rng.random(n_src) + 0.5
    R = regridder_from_coo(n_dst, n_src, rows, cols, vals, da, sa)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n_dst, n_src)).tocsr()
    x = rng.random(n_src)
    y = np.full(n_dst, np.nan); regrid_(y, R, x)
    assert np.allclose(y, (A @ x) / da, rtol=1e-12, atol=1e-14)
    y2 = np.full(n_dst, np.nan); regrid_(y2, R, x, normalize=False)
    assert np.allclose(y2, A @ x, rtol=1e-12, atol=1e-14)
    yt = rng.random(n_dst)
    xb = np.full(n_src, np.nan); regrid_(xb, transpose(R), yt)
    assert np.allclose(xb, (A.T @ yt) / sa, rtol=1e-12, atol=1e-14)
    # repeated launches reuse the self-resetting completion ticket
    for _ in range(3):
        y3 = np.zeros(n_dst); regrid_(y3, R, x)
        assert np.array_equal(y3, y)


def test_torch_default_stream_ordering(gpu):
    """CUDA-tensor applies run on torch's CURRENT stream -- also when that is the legacy
    default stream, which torch reports as handle 0 -- so they are ordered after the kernels that produce
    the inputs and before the consumers of the outputs, without any explicit stream plumbing."""
    import torch
    dst, src = grids.lonlat_grid(720, 360), grids.healpix_grid(128, "ring")
    R = Regridder(dst, src)                          # host vertices: the handle starts on the library stream
    A = R.intersections.tocsr()
    rng = np.random.default_rng(8)
    x = rng.random(src.ncells)
    ref = (A @ x) / R.dst_areas
    big = torch.ones(64 << 20, device="cuda")        # a long-running producer in front of the input
    for use_side_stream in (False, True):
        ctx = torch.cuda.stream(torch.cuda.Stream()) if use_side_stream else torch.cuda.stream(torch.cuda.default_stream())
        with ctx:
            for _ in range(3):
                xd = torch.zeros(src.ncells, dtype=torch.float64, device="cuda")
                yd = torch.full((dst.ncells,), float("nan"), dtype=torch.float64, device="cuda")
                big.mul_(1.0000001); big.mul_(0.9999999)         # keep the stream busy ...
                xd.copy_(torch.from_numpy(x).cuda(), non_blocking=True)   # ... then produce the input on it
                regrid_(yd, R, xd, asynchronous=True)
                out = yd * 2.0                                    # consumer on the same stream
                got = out.cpu().numpy() / 2.0
                assert np.allclose(got, ref, rtol=1e-13, atol=0)
    assert R.intersections._h.stream is not None      # the handle followed torch's stream
    # dropping the handle while an asynchronous apply may still be in flight is safe (destroy waits on the stream)
    yd = torch.zeros(dst.ncells, dtype=torch.float64, device="cuda")
    for _ in range(5):
        R2 = Regridder(dst, src)
        regrid_(yd, R2, torch.from_numpy(x).cuda(), asynchronous=True)
        del R2
    torch.cuda.synchronize()
    assert np.allclose(yd.cpu().numpy(), ref, rtol=1e-13, atol=0)


def test_set_areas_reaches_the_device(gpu):
    """regrid! divides by `regridder.dst_areas` (regrid.jl:104-118), which users may replace (masking, custom
    normalisation): the host vectors are read-only views and set_areas pushes new values to the device."""
    from crg_b200.regridder import set_areas
    dst, src = grids.lonlat_grid(36, 18), grids.healpix_grid(8, "ring")
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    x = np.random.default_rng(2).random(src.ncells)
    with pytest.raises(ValueError):
        R.dst_areas[0] = 1.0                         # read-only: an in-place edit would be silently ignored
    new = R.dst_areas * np.linspace(1, 2, dst.ncells)
    set_areas(R, dst_areas=new)
    assert np.array_equal(R.dst_areas, new) and transpose(R).src_areas is R.dst_areas
    y = np.zeros(dst.ncells); regrid_(y, R, x)
    assert np.allclose(y, (A @ x) / new, rtol=1e-13)
    T = transpose(R)
    news = R.src_areas * 3.0
    set_areas(T, dst_areas=news)                     # transpose(R).dst_areas is R.src_areas
    xb = np.zeros(src.ncells); regrid_(xb, T, y)
    assert np.allclose(xb, (A.T @ y) / news, rtol=1e-13)

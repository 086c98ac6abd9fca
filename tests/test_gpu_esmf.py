"""ESMF weight-file export/import, mirroring /root/reference/test/extensions/ncdatasets.jl:32-122."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from crg_b200 import grids
from crg_b200.esmf import load_esmf_weights, save_esmf_weights
from crg_b200.regridder import Regridder, regrid_

pytestmark = pytest.mark.gpu


def build_regridder():
    # two planar grids tiling [0,2]x[0,2] (ncdatasets.jl:10-30)
    dst = grids.planar_regular_grid(np.linspace(0, 2, 3), np.linspace(0, 2, 3))
    src = grids.planar_regular_grid(np.linspace(0, 2, 5), np.linspace(0, 2, 4))
    return Regridder(dst, src, normalize=False)


def test_esmf_round_trip_schema_and_values(gpu, tmp_path):
    from scipy.io import netcdf_file
    r = build_regridder()
    A = r.intersections.tocsc()
    path = str(tmp_path / "sub" / "weights.nc")
    assert save_esmf_weights(path, r) == path and os.path.isfile(path)
    with netcdf_file(path, "r", mmap=False) as ds:
        assert ds.dimensions["n_a"] == len(r.src_areas) and ds.dimensions["n_b"] == len(r.dst_areas)
        assert ds.dimensions["n_s"] == A.nnz
        for v in ("S", "row", "col", "frac_a", "frac_b", "area_a", "area_b"):
            assert v in ds.variables
        assert ds.variables["S"].shape == (A.nnz,) and ds.variables["frac_b"].shape == (len(r.dst_areas),)
        assert ds.normalization == b"destarea" and ds.source_grid == b"source" and ds.destination_grid == b"destination"
        assert not hasattr(ds, "created_at") and not hasattr(ds, "source_grid_shape")
        S = ds.variables["S"][:].copy(); row = ds.variables["row"][:].copy(); col = ds.variables["col"][:].copy()
        frac_a = ds.variables["frac_a"][:].copy(); frac_b = ds.variables["frac_b"][:].copy()
        area_a = ds.variables["area_a"][:].copy(); area_b = ds.variables["area_b"][:].copy()
    assert np.allclose(area_a, r.src_areas) and np.allclose(area_b, r.dst_areas)
    assert row.min() >= 1 and col.min() >= 1                      # 1-based
    rebuilt = sp.coo_matrix((S * r.dst_areas[row - 1], (row - 1, col - 1)), shape=A.shape)
    assert np.allclose(rebuilt.toarray(), A.toarray())
    assert np.allclose(frac_a, np.asarray(A.sum(0)).ravel() / r.src_areas)
    assert np.allclose(frac_b, np.asarray(A.sum(1)).ravel() / r.dst_areas)
    assert np.allclose(frac_a, 1.0, atol=1e-12) and np.allclose(frac_b, 1.0, atol=1e-12)
    # import: the file is a persistent regridder
    r2 = load_esmf_weights(path)
    assert abs(r2.intersections.tocsc() - A).max() < 1e-15
    x = np.random.default_rng(0).random(len(r.src_areas))
    y1, y2 = np.zeros(len(r.dst_areas)), np.zeros(len(r.dst_areas))
    regrid_(y1, r, x); regrid_(y2, r2, x)
    assert np.allclose(y1, y2, rtol=1e-14)


def test_esmf_optional_attributes(gpu, tmp_path):
    from scipy.io import netcdf_file
    path = str(tmp_path / "w.nc")
    save_esmf_weights(path, build_regridder(), src_grid_name="era5_0.25deg", dst_grid_name="c90",
                      src_shape=(720, 361), dst_shape=(90, 90, 6), created_at="2026-04-21T00:00:00")
    with netcdf_file(path, "r", mmap=False) as ds:
        assert ds.source_grid == b"era5_0.25deg" and ds.destination_grid == b"c90"
        assert ds.created_at == b"2026-04-21T00:00:00"
        assert list(ds.source_grid_shape) == [720, 361] and list(ds.destination_grid_shape) == [90, 90, 6]


def test_esmf_spherical(gpu, tmp_path):
    r = Regridder(grids.lonlat_spec(36, 18), grids.healpix_spec(4, "ring"))
    path = save_esmf_weights(str(tmp_path / "s.nc"), r, src_shape=(192,), dst_shape=(36, 18))
    r2 = load_esmf_weights(path)
    assert abs(r2.intersections.tocsc() - r.intersections.tocsc()).max() < 1e-16
    assert np.allclose(r2.dst_areas, r.dst_areas) and np.allclose(r2.src_areas, r.src_areas)

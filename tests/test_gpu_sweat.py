"""Every ordered pair of a set of global grids (mirrors /root/reference/test/sweat.jl:94-176 and
test/sweat_field.jl:114-188 at reduced sizes): the intersection matrix reproduces the geometric
cell areas on both sides, constants stay constant, and analytic fields are conserved."""
import itertools

import numpy as np
import pytest

from crg_b200 import fields, grids
from crg_b200.regridder import Regridder, regrid_, transpose

pytestmark = pytest.mark.gpu

GRIDS = {
    "lonlat90x45": lambda: grids.lonlat_spec(90, 45),
    "healpix16nested": lambda: grids.healpix_spec(16, "nested"),
    "healpix16ring": lambda: grids.healpix_spec(16, "ring"),
    "FullClenshaw12": lambda: grids.full_clenshaw_spec(12),
    "FullGaussian12": lambda: grids.full_gaussian_spec(12),
    "C12": lambda: grids.cubed_sphere_spec(12),
    "lonlat_r6371km": lambda: grids.lonlat_grid(60, 30, radius=6371e3),
}
PAIRS = [(a, b) for a, b in itertools.permutations(GRIDS, 2)
         if ("r6371km" in a) == ("r6371km" in b) or True]


@pytest.mark.parametrize("dst_name,src_name", [(a, b) for a, b in PAIRS if "r6371" not in a + b])
def test_intersection_areas_agree_and_constants(gpu, dst_name, src_name):
    dst, src = GRIDS[dst_name](), GRIDS[src_name]()
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    rtol = np.sqrt(np.finfo(float).eps)                   # TestHelpers.jl:70-73
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=rtol, atol=0)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=rtol, atol=0)
    y = np.zeros(dst.ncells); regrid_(y, R, np.ones(src.ncells))
    assert np.allclose(y, 1.0, atol=1e-10)                # test/usecases/constant_field.jl:41-155 (atol 1e-3 there)
    xb = np.zeros(src.ncells); regrid_(xb, transpose(R), np.ones(dst.ncells))
    assert np.allclose(xb, 1.0, atol=1e-10)


def test_radius_is_carried_by_the_manifold(gpu):
    # best_manifold = Spherical(; radius = grid.radius) (OceananigansExt.jl:284-289): areas scale with R^2
    Rm = 6371e3
    R = Regridder(grids.lonlat_grid(60, 30, radius=Rm), grids.healpix_grid(8, "ring", radius=Rm))
    assert abs(R.dst_areas.sum() / (4 * np.pi * Rm ** 2) - 1) < 1e-12
    assert abs(R.src_areas.sum() / (4 * np.pi * Rm ** 2) - 1) < 1e-12
    A = R.intersections.tocsr()
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=1.5e-8)
    Ru = Regridder(grids.lonlat_grid(60, 30), grids.healpix_grid(8, "ring"))
    assert np.allclose(A.data, Ru.intersections.tocsr().data * Rm ** 2, rtol=1e-13)


@pytest.mark.parametrize("field", list(fields.EXAMPLE_FIELDS))
def test_analytic_fields_are_conserved(gpu, field):
    f = fields.EXAMPLE_FIELDS[field]
    src_s, dst_s = grids.lonlat_spec(180, 90), grids.healpix_spec(32, "ring")
    src, dst = src_s.materialize(), dst_s.materialize()
    R = Regridder(dst_s, src_s)
    lon, lat = grids.cell_centers_lonlat(src)
    x = f(np.mod(lon, 360.0) if field == "longitude" else lon, lat)
    y = np.zeros(dst.ncells); regrid_(y, R, x)
    # conservation of the area-weighted integral (test/usecases/oceananigans.jl:37-159: rtol 1e-10)
    assert abs((y * R.dst_areas).sum() / (x * R.src_areas).sum() - 1) < 1e-12
    # and the regridded field follows the analytic one (sweat_field.jl: rtol 1e-2, 5e-2 for longitude)
    lon_d, lat_d = grids.cell_centers_lonlat(dst)
    ref = f(np.mod(lon_d, 360.0) if field == "longitude" else lon_d, lat_d)
    ok = np.abs(lat_d) < 80
    if field == "longitude":
        ok &= (np.mod(lon_d, 360.0) > 5) & (np.mod(lon_d, 360.0) < 355)      # cells crossing the seam average 0 and 360
    err = np.abs(y[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-3)
    assert np.quantile(err, 0.99) < 5e-2
    # round trip back to the source grid stays close (oceananigans.jl:15-35: rtol 1e-5 on the mean)
    xb = np.zeros(src.ncells); regrid_(xb, transpose(R), y)
    assert abs((xb * R.src_areas).sum() / (x * R.src_areas).sum() - 1) < 1e-12


COINCIDENT = [
    ("healpix8 nested<-ring", lambda: (grids.healpix_spec(8, "nested"), grids.healpix_spec(8, "ring"))),
    ("healpix32 nested<-ring", lambda: (grids.healpix_spec(32, "nested"), grids.healpix_spec(32, "ring"))),
    ("healpix128 nested<-ring", lambda: (grids.healpix_spec(128, "nested"), grids.healpix_spec(128, "ring"))),
    ("healpix16<-healpix64", lambda: (grids.healpix_spec(16, "ring"), grids.healpix_spec(64, "ring"))),
    ("lonlat90x45 onto itself", lambda: (grids.lonlat_spec(90, 45), grids.lonlat_spec(90, 45))),
    ("lonlat45x30<-lonlat360x180", lambda: (grids.lonlat_spec(45, 30), grids.lonlat_spec(360, 180))),
    ("C8<-C32", lambda: (grids.cubed_sphere_spec(8), grids.cubed_sphere_spec(32))),
]


@pytest.mark.parametrize("name,make", COINCIDENT, ids=[c[0] for c in COINCIDENT])
def test_coincident_edges_identical_and_nested_grids(gpu, name, make):
    """Identical cells in another order and exactly nested refinements: every clip line carries
    subject vertices whose signed distances are rounding noise.  The matrix must still reproduce
    the cell areas on both sides (a crossing at an end of an edge is that vertex -- geom.cuh)."""
    dst, src = make()
    R = Regridder(dst, src)
    A = R.intersections.tocsr()
    assert np.allclose(np.asarray(A.sum(1)).ravel(), R.dst_areas, rtol=1e-9, atol=0)
    assert np.allclose(np.asarray(A.sum(0)).ravel(), R.src_areas, rtol=1e-9, atol=0)

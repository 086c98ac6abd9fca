"""Host-side synthetic grid generators (numpy, Float64).

Every generator returns a :class:`Grid`: a flat "vertex soup" of open rings in
**field-linear order** -- the same order ``Trees.getcell(tree)`` iterates in the
reference -- which is the input contract of ``crg_build`` (include/crg_b200.h).

Cell conventions follow the reference (all citations relative to /root/reference):

* ``CellBasedGrid``: (Nx+1)x(Ny+1) vertex matrix P; cell (i,j) ring =
  P[i,j], P[i+1,j], P[i+1,j+1], P[i,j+1]; linear index i + (j-1)*Nx
  (src/trees/grids.jl:71-84, src/trees/interfaces.jl:236-243).  Edges are
  great-circle arcs between unit vectors on the sphere.
* lon-lat (Oceananigans ``LatitudeLongitudeGrid``): vertices at Face/Face nodes
  (ext/ConservativeRegriddingOceananigansExt.jl:23-60,242-264).
* HEALPix: 4 pixel corners (N, W, S, E) joined by great-circle arcs, nested or
  ring field order (ext/ConservativeRegriddingHealpixExt.jl:76-90,138-167).
* RingGrids full grids: pole-pinned latitude edges at mid-points of the ring
  latitudes, longitude edges at lond[1]-dlon/2+(i-1)dlon, ring-major north->south
  field order (ext/ConservativeRegriddingRingGridsExt.jl:22-50).

``UnitSphereFromGeographic`` in the reference uses degree-exact trigonometry
(``sincosd``), so poles are exactly (0,0,+-1) and lon=360 == lon=0 bit for bit;
:func:`sincosd` reproduces that property.
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field
from typing import Optional

import numpy as np

PLANAR = 0
SPHERICAL = 1


@dataclass
class Grid:
    """Flat polygon soup in field-linear order.

    verts : (ncells, nv, dim) float64 for fixed-size rings, or (total_verts, dim)
            with ``offsets`` (int32, ncells+1) for ragged rings.  Rings are open
            (the closing vertex of the reference's 5-point rings is dropped).
    manifold : PLANAR (dim=2) or SPHERICAL (dim=3, unit vectors).
    radius : sphere radius carried by ``best_manifold`` (OceananigansExt.jl:284-289).
    """

    verts: np.ndarray
    manifold: int
    offsets: Optional[np.ndarray] = None
    radius: float = 1.0
    name: str = ""
    meta: dict = _dc_field(default_factory=dict)

    @property
    def ncells(self) -> int:
        if self.offsets is not None:
            return int(self.offsets.shape[0] - 1)
        return int(self.verts.shape[0])

    @property
    def dim(self) -> int:
        return 3 if self.manifold == SPHERICAL else 2

    @property
    def nv(self) -> int:
        return 0 if self.offsets is not None else int(self.verts.shape[1])

    def slice(self, lo: int, hi: int) -> "Grid":
        """Contiguous block of cells [lo, hi) in field order (a destination shard)."""
        if self.offsets is None:
            return Grid(self.verts[lo:hi], self.manifold, None, self.radius,
                        f"{self.name}[{lo}:{hi}]", dict(self.meta))
        off = self.offsets[lo:hi + 1]
        return Grid(self.verts[off[0]:off[-1]], self.manifold, (off - off[0]).astype(np.int32),
                    self.radius, f"{self.name}[{lo}:{hi}]", dict(self.meta))

    def cell(self, i: int) -> np.ndarray:
        if self.offsets is None:
            return self.verts[i]
        return self.verts[self.offsets[i]:self.offsets[i + 1]]


@dataclass
class GridSpec:
    """A structured grid described by a few numbers; its cell vertices are generated ON THE DEVICE
    inside ``crg_build_grids`` (csrc/gridgen.cuh), so no vertex soup is built or uploaded by the
    host.  ``kind``: "lonlat" | "healpix" | "full_ring" | "cubed_sphere" | "reduced_ring" (include/crg_b200.h).
    ``materialize()`` gives the equivalent host :class:`Grid` (same conventions, same order)."""

    kind: str
    n1: int
    n2: int = 0
    p: tuple = (0.0, 0.0, 0.0, 0.0)
    flags: int = 0
    lat_deg: Optional[np.ndarray] = None
    radius: float = 1.0
    name: str = ""
    manifold: int = 1
    cell_lo: int = 0          # a slice [cell_lo, cell_hi) of the field-linear order (0, 0 = the whole grid)
    cell_hi: int = 0

    @property
    def ncells(self) -> int:
        if self.cell_lo or self.cell_hi:
            return self.cell_hi - self.cell_lo
        return self.ncells_full

    def slice(self, lo: int, hi: int) -> "GridSpec":
        """Cells [lo, hi) of this (possibly already sliced) grid -- a destination block or a source halo of a
        sharded build; still generated on the device."""
        import dataclasses
        n = self.ncells
        if not (0 <= lo <= hi <= n):
            raise IndexError(f"slice [{lo}, {hi}) outside [0, {n})")
        base = self.cell_lo
        if lo == 0 and hi == n and not (self.cell_lo or self.cell_hi):
            return self
        return dataclasses.replace(self, cell_lo=base + lo, cell_hi=base + hi, name=f"{self.name}[{lo}:{hi}]")

    @property
    def ncells_full(self) -> int:
        if self.kind == "healpix":
            return 12 * self.n1 * self.n1
        if self.kind == "cubed_sphere":
            return 6 * self.n1 * self.n1
        if self.kind == "reduced_ring":
            a, b, nh = int(self.p[1]), int(self.p[2]), self.n2 // 2
            return 2 * (a * nh + b * (nh * (nh + 1) // 2))
        return self.n1 * self.n2

    def materialize(self) -> "Grid":
        if self.cell_lo or self.cell_hi:
            import dataclasses
            return dataclasses.replace(self, cell_lo=0, cell_hi=0).materialize().slice(self.cell_lo, self.cell_hi)
        if self.kind == "lonlat":
            return lonlat_grid(self.n1, self.n2, *self.p, radius=self.radius)
        if self.kind == "healpix":
            return healpix_grid(self.n1, "nested" if self.flags & 1 else "ring", radius=self.radius)
        if self.kind == "full_ring":
            return full_ring_grid(self.lat_deg, self.n1, self.p[0], self.radius, self.name or "fullring")
        if self.kind == "cubed_sphere":
            return cubed_sphere_grid(self.n1, radius=self.radius)
        if self.kind == "reduced_ring":
            return octahedral_gaussian_grid(self.n2 // 2, radius=self.radius)
        raise ValueError(self.kind)


def ring_table(g):
    """For DESCRIBED grids whose field-linear order is ring-major (cells of one latitude band are contiguous): per
    ring the first cell index (of the whole grid) and the z = sin(lat) range of its cells' VERTICES --
    ``(start[nr + 1], zlo[nr], zhi[nr])`` in ring order -- else None.  A destination block's z-range then selects the contiguous source cell range that can
    reach it (the halo of a sharded build, dist.py)."""
    if not isinstance(g, GridSpec):
        return None
    kind, n1, n2, p, flags, latd = g.kind, g.n1, g.n2, g.p, g.flags, g.lat_deg
    sind = lambda d: np.sin(np.radians(np.asarray(d, dtype=np.float64)))  # noqa: E731
    if kind == "lonlat":
        e = sind(p[2] + (p[3] - p[2]) * np.arange(n2 + 1) / n2)           # south -> north
        return np.arange(n2 + 1, dtype=np.int64) * n1, e[:-1], e[1:]
    if kind in ("full_ring", "reduced_ring"):
        e = sind(_pole_pinned_lat_edges(np.asarray(latd, dtype=np.float64)))    # north -> south
        if kind == "full_ring":
            start = np.arange(n2 + 1, dtype=np.int64) * n1
        else:
            a, b, nh = int(p[1]), int(p[2]), n2 // 2
            j = np.concatenate([np.arange(1, nh + 1), np.arange(nh, 0, -1)])
            start = np.concatenate([[0], np.cumsum(a + b * j)]).astype(np.int64)
        return start, e[1:], e[:-1]
    if kind == "healpix" and not (flags & 1):
        ns = n1
        i = np.arange(0, 4 * ns + 1, dtype=np.float64)                   # ring index 0 (north pole) .. 4 nside (south pole)
        zc = np.where(i <= ns, 1.0 - i * i / (3.0 * ns * ns),
                      np.where(i <= 3 * ns, 4.0 / 3.0 - 2.0 * i / (3.0 * ns), -1.0 + (4 * ns - i) ** 2 / (3.0 * ns * ns)))
        r = np.arange(1, 4 * ns)                                          # rings 1 .. 4 nside - 1, north -> south
        npix = np.where(r < ns, 4 * r, np.where(r <= 3 * ns, 4 * ns, 4 * (4 * ns - r)))
        start = np.concatenate([[0], np.cumsum(npix)]).astype(np.int64)
        return start, zc[r + 1], zc[r - 1]                                # a pixel's N / S corners sit on the neighbouring rings
    return None


def lonlat_spec(nlon, nlat, lon0=0.0, lon1=360.0, lat0=-90.0, lat1=90.0, radius=1.0) -> GridSpec:
    return GridSpec("lonlat", nlon, nlat, (float(lon0), float(lon1), float(lat0), float(lat1)), 0, None, radius,
                    f"lonlat{nlon}x{nlat}")


def healpix_spec(nside, order="ring", radius=1.0) -> GridSpec:
    assert nside >= 1 and (nside & (nside - 1)) == 0
    return GridSpec("healpix", nside, 0, (0.0,) * 4, 1 if order == "nested" else 0, None, radius, f"healpix{nside}{order}")


def full_gaussian_spec(nlat_half, radius=1.0) -> GridSpec:
    return GridSpec("full_ring", 4 * nlat_half, 2 * nlat_half, (0.0, 0.0, 0.0, 0.0), 0,
                    np.ascontiguousarray(gaussian_latitudes(2 * nlat_half)), radius, f"F{nlat_half}")


def full_clenshaw_spec(nlat_half, radius=1.0) -> GridSpec:
    nlat = 2 * nlat_half - 1
    latd = 90.0 - 90.0 * (np.arange(1, nlat + 1) / nlat_half)
    return GridSpec("full_ring", 4 * nlat_half, nlat, (0.0, 0.0, 0.0, 0.0), 0, np.ascontiguousarray(latd), radius,
                    f"FullClenshaw{nlat_half}")


def octahedral_gaussian_spec(nlat_half, radius=1.0) -> GridSpec:
    """O<nlat_half> (SpeedyWeather's default grid, BASELINE config 4): 2*nlat_half Gaussian rings, 16 + 4j points in
    the ring of rank j from either pole (= :func:`octahedral_gaussian_grid`, generated on the device)."""
    return GridSpec("reduced_ring", 0, 2 * nlat_half, (0.0, 16.0, 4.0, 0.0), 0,
                    np.ascontiguousarray(gaussian_latitudes(2 * nlat_half)), radius, f"O{nlat_half}")


def cubed_sphere_spec(n, radius=1.0) -> GridSpec:
    return GridSpec("cubed_sphere", n, 0, (0.0,) * 4, 0, None, radius, f"C{n}")


# ----------------------------------------------------------------------------
# degree-exact trigonometry
# ----------------------------------------------------------------------------

def sincosd(deg):
    """sin and cos of an angle in degrees, exact at multiples of 90 (like Julia's
    ``sincosd``): the argument is reduced in degrees before conversion to radians."""
    x = np.asarray(deg, dtype=np.float64)
    r = np.remainder(x, 360.0)                   # exact for |x| < 2^53
    q = np.floor((r + 45.0) / 90.0)              # nearest quadrant 0..4
    a = np.deg2rad(r - 90.0 * q)                 # |a| <= pi/4, exact 0 at multiples of 90
    s, c = np.sin(a), np.cos(a)
    qi = q.astype(np.int64) % 4
    sin = np.select([qi == 0, qi == 1, qi == 2, qi == 3], [s, c, -s, -c])
    cos = np.select([qi == 0, qi == 1, qi == 2, qi == 3], [c, -s, -c, s])
    return sin + 0.0, cos + 0.0                  # +0.0 normalises -0.0


def unit_sphere_from_geographic(lon_deg, lat_deg) -> np.ndarray:
    """(lon, lat) in degrees -> unit xyz; GeometryOps ``UnitSphereFromGeographic``
    as used at src/trees/grids.jl:195 and OceananigansExt.jl:250."""
    slon, clon = sincosd(lon_deg)
    slat, clat = sincosd(lat_deg)
    return np.stack([clat * clon, clat * slon, slat + 0.0 * clon], axis=-1)


def geographic_from_unit_sphere(xyz):
    xyz = np.asarray(xyz)
    lon = np.degrees(np.arctan2(xyz[..., 1], xyz[..., 0]))
    lat = np.degrees(np.arctan2(xyz[..., 2], np.hypot(xyz[..., 0], xyz[..., 1])))
    return lon, lat


# ----------------------------------------------------------------------------
# structured ("CellBasedGrid") helpers
# ----------------------------------------------------------------------------

def cells_from_vertex_matrix(P: np.ndarray) -> np.ndarray:
    """P: (Nx+1, Ny+1, dim) vertex matrix -> (Nx*Ny, 4, dim) cells, i fastest.

    Ring order and linear index follow src/trees/grids.jl:71-84 and
    src/trees/interfaces.jl:236-243."""
    nx, ny = P.shape[0] - 1, P.shape[1] - 1
    c = np.empty((ny, nx, 4, P.shape[2]), dtype=np.float64)
    Pt = np.transpose(P, (1, 0, 2))            # [j, i]
    c[:, :, 0] = Pt[:-1, :-1]
    c[:, :, 1] = Pt[:-1, 1:]
    c[:, :, 2] = Pt[1:, 1:]
    c[:, :, 3] = Pt[1:, :-1]
    return np.ascontiguousarray(c.reshape(nx * ny, 4, P.shape[2]))


def lonlat_grid(nlon: int, nlat: int, lon0: float = 0.0, lon1: float = 360.0,
                lat0: float = -90.0, lat1: float = 90.0, radius: float = 1.0) -> Grid:
    """Regular lon-lat grid on the sphere, cells bounded by great-circle arcs
    between Face/Face nodes (OceananigansExt.jl:47-60,242-264).  Field order:
    longitude fastest, south -> north (``vec(interior(field))``)."""
    lon = lon0 + (lon1 - lon0) * (np.arange(nlon + 1) / nlon)
    lat = lat0 + (lat1 - lat0) * (np.arange(nlat + 1) / nlat)
    P = unit_sphere_from_geographic(lon[:, None], lat[None, :])
    g = Grid(cells_from_vertex_matrix(P), SPHERICAL, None, radius, f"lonlat{nlon}x{nlat}")
    g.meta.update(kind="lonlat", nlon=nlon, nlat=nlat, lon0=lon0, lon1=lon1, lat0=lat0, lat1=lat1,
                  shape=(nlon, nlat))
    return g


def planar_regular_grid(x: np.ndarray, y: np.ndarray) -> Grid:
    """``RegularGrid(x, y)`` on the plane (src/trees/grids.jl:96-115)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    P = np.stack(np.broadcast_arrays(x[:, None], y[None, :]), axis=-1)
    g = Grid(cells_from_vertex_matrix(P), PLANAR, None, 1.0, f"planar{len(x)-1}x{len(y)-1}")
    g.meta.update(kind="planar_regular", shape=(len(x) - 1, len(y) - 1))
    return g


def planar_unit_square_grid(nx: int, ny: int) -> Grid:
    """The ``make_grid(nx, ny)`` helper of test/regridding.jl:46-55 (cells of the unit
    square, i fastest)."""
    return planar_regular_grid(np.arange(nx + 1) / nx, np.arange(ny + 1) / ny)


def polygons_grid(polys, manifold: int = PLANAR, radius: float = 1.0, name: str = "polys") -> Grid:
    """Arbitrary iterable of (open or closed) rings -> ragged Grid in iteration order
    (``FlatNoTree`` path, src/trees/interfaces.jl:109-118)."""
    rings = []
    for p in polys:
        r = np.asarray(p, dtype=np.float64)
        if len(r) > 1 and np.array_equal(r[0], r[-1]):
            r = r[:-1]
        rings.append(r)
    nvs = {len(r) for r in rings}
    if len(nvs) == 1:
        return Grid(np.ascontiguousarray(np.stack(rings)), manifold, None, radius, name)
    off = np.zeros(len(rings) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(r) for r in rings])
    return Grid(np.ascontiguousarray(np.concatenate(rings)), manifold, off, radius, name)


# ----------------------------------------------------------------------------
# HEALPix
# ----------------------------------------------------------------------------

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _compress_bits(v):
    """De-interleave: keep the even bits of a 64-bit integer."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def _spread_bits(v):
    v = v & 0x00000000FFFFFFFF
    v = (v | (v << 16)) & 0x0000FFFF0000FFFF
    v = (v | (v << 8)) & 0x00FF00FF00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v << 2)) & 0x3333333333333333
    v = (v | (v << 1)) & 0x5555555555555555
    return v


def healpix_nest2xyf(nside: int, pix):
    pix = np.asarray(pix, dtype=np.int64)
    npface = nside * nside
    face = pix // npface
    p = pix % npface
    return _compress_bits(p), _compress_bits(p >> 1), face


def healpix_xyf2nest(nside: int, ix, iy, face):
    return np.asarray(face, np.int64) * nside * nside + _spread_bits(np.asarray(ix, np.int64)) \
        + (_spread_bits(np.asarray(iy, np.int64)) << 1)


def healpix_xyf2ring(nside: int, ix, iy, face):
    """0-based ring-order pixel index of face coordinates (standard HEALPix)."""
    ix = np.asarray(ix, np.int64)
    iy = np.asarray(iy, np.int64)
    face = np.asarray(face, np.int64)
    nl4 = 4 * nside
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    jr = _JRLL[face] * nside - ix - iy - 1
    north = jr < nside
    south = jr > 3 * nside
    nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
    n_before = np.where(north, 2 * nr * (nr - 1),
                        np.where(south, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * nl4))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    return n_before + jp - 1


def healpix_nest2ring(nside: int, pix):
    ix, iy, f = healpix_nest2xyf(nside, pix)
    return healpix_xyf2ring(nside, ix, iy, f)


def healpix_ring2nest(nside: int, pix):
    """Inverse permutation of :func:`healpix_nest2ring` (``Healpix.ring2nest`` used by
    ``getcell`` in ring order, HealpixExt.jl:162)."""
    npix = 12 * nside * nside
    nest = np.arange(npix, dtype=np.int64)
    ring_of_nest = healpix_nest2ring(nside, nest)
    inv = np.empty(npix, dtype=np.int64)
    inv[ring_of_nest] = nest
    return inv[np.asarray(pix, np.int64)]


def _healpix_loc(x, y, face):
    """Point (x, y) in [0,1]^2 face coordinates -> unit vector (HEALPix ``xyf2loc``)."""
    jr = _JRLL[face] - x - y
    north = jr < 1.0
    south = jr > 3.0
    nr = np.where(north, jr, np.where(south, 4.0 - jr, 1.0))
    tmp = nr * nr / 3.0
    z = np.where(north, 1.0 - tmp, np.where(south, tmp - 1.0, (2.0 - jr) * 2.0 / 3.0))
    # sin(theta) without cancellation in the polar caps: 1 - z^2 = tmp (2 - tmp)
    sth = np.where(north | south, np.sqrt(np.maximum(tmp * (2.0 - tmp), 0.0)),
                   np.sqrt(np.maximum((1.0 - z) * (1.0 + z), 0.0)))
    t = _JPLL[face] * nr + x - y
    t = np.where(t < 0.0, t + 8.0, t)
    t = np.where(t >= 8.0, t - 8.0, t)
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = np.where(nr < 1e-15, 0.0, (0.25 * np.pi * t) / np.where(nr < 1e-15, 1.0, nr))
    return np.stack([sth * np.cos(phi), sth * np.sin(phi), z], axis=-1)


def healpix_corners_nested(nside: int, pix) -> np.ndarray:
    """(len(pix), 4, 3) corners N, W, S, E (CCW seen from outside) of nested pixels,
    = ``Healpix.boundariesRing(res, nest2ring(p), 1)`` (HealpixExt.jl:76-90)."""
    ix, iy, face = healpix_nest2xyf(nside, pix)
    x0 = ix / nside
    x1 = (ix + 1) / nside
    y0 = iy / nside
    y1 = (iy + 1) / nside
    out = np.empty(ix.shape + (4, 3), dtype=np.float64)
    out[..., 0, :] = _healpix_loc(x1, y1, face)   # N
    out[..., 1, :] = _healpix_loc(x0, y1, face)   # W
    out[..., 2, :] = _healpix_loc(x0, y0, face)   # S
    out[..., 3, :] = _healpix_loc(x1, y0, face)   # E
    return out


def healpix_grid(nside: int, order: str = "ring", radius: float = 1.0) -> Grid:
    """HEALPix cells in ``nested`` or ``ring`` field order (HealpixExt.jl:138-167)."""
    assert nside >= 1 and (nside & (nside - 1)) == 0, "nside must be a power of two"
    npix = 12 * nside * nside
    if order == "nested":
        nest = np.arange(npix, dtype=np.int64)
    elif order == "ring":
        nest = healpix_ring2nest(nside, np.arange(npix, dtype=np.int64))
    else:
        raise ValueError("order must be 'nested' or 'ring'")
    g = Grid(np.ascontiguousarray(healpix_corners_nested(nside, nest)), SPHERICAL, None, radius,
             f"healpix{nside}{order}")
    g.meta.update(kind="healpix", nside=nside, order=order)
    return g


# ----------------------------------------------------------------------------
# cubed sphere (equiangular gnomonic), Gaussian and octahedral grids
# ----------------------------------------------------------------------------

def cubed_sphere_grid(n: int, radius: float = 1.0) -> Grid:
    """Equiangular gnomonic cubed sphere C<n>: 6 panels concatenated panel-major, i
    fastest inside a panel.  Stand-in for the Oceananigans conformal cubed sphere of
    BASELINE config 3 (the reference builds one regridder per panel,
    examples/oceananigans_cubed_sphere.jl:14-23); each panel is a ``CellBasedGrid``."""
    a = np.tan(-np.pi / 4 + (np.pi / 2) * (np.arange(n + 1) / n))
    a[0], a[-1], = -1.0, 1.0
    if n % 2 == 0:
        a[n // 2] = 0.0
    X, Y = np.meshgrid(a, a, indexing="ij")
    one = np.ones_like(X)
    # (right-handed) panel frames: local (xi, eta, outward)
    panels = [
        (one, X, Y),      # +x : xi -> +y, eta -> +z
        (-X, one, Y),     # +y : xi -> -x, eta -> +z
        (-one, -X, Y),    # -x : xi -> -y, eta -> +z
        (X, -one, Y),     # -y : xi -> +x, eta -> +z
        (-Y, X, one),     # +z : xi -> +y, eta -> -x
        (Y, X, -one),     # -z : xi -> +y, eta -> +x
    ]
    cells = []
    for (px, py, pz) in panels:
        P = np.stack([px, py, pz], axis=-1)
        P = P / np.linalg.norm(P, axis=-1, keepdims=True)
        cells.append(cells_from_vertex_matrix(P))
    g = Grid(np.ascontiguousarray(np.concatenate(cells)), SPHERICAL, None, radius, f"C{n}")
    g.meta.update(kind="cubed_sphere", n=n, panels=6, shape=(n, n))
    return g


def gaussian_latitudes(nlat: int) -> np.ndarray:
    """Gaussian latitudes in degrees, north -> south (``RingGrids.get_latd``)."""
    x, _ = np.polynomial.legendre.leggauss(nlat)
    return np.degrees(np.arcsin(x[::-1]))


def _pole_pinned_lat_edges(latd: np.ndarray) -> np.ndarray:
    """RingGridsExt.jl:28-34: +90, mid-points of consecutive ring latitudes, -90."""
    e = np.empty(len(latd) + 1)
    e[0], e[-1] = 90.0, -90.0
    e[1:-1] = 0.5 * (latd[:-1] + latd[1:])
    return e


def full_ring_grid(latd: np.ndarray, nlon: int, lon_first: float = 0.0, radius: float = 1.0,
                   name: str = "fullring") -> Grid:
    """RingGrids ``AbstractFullGrid`` cells (RingGridsExt.jl:22-50): field order is
    ring-major north -> south, longitude fastest; the first cell of every ring
    straddles ``lon_first``."""
    nlat = len(latd)
    lat_edges = _pole_pinned_lat_edges(np.asarray(latd, dtype=np.float64))
    dlon = 360.0 / nlon
    lon_edges = lon_first - dlon / 2 + np.arange(nlon + 1) * dlon
    # vertex matrix stored south -> north in j, like the reference
    P = unit_sphere_from_geographic(lon_edges[:, None], lat_edges[::-1][None, :])
    cells = cells_from_vertex_matrix(P).reshape(nlat, nlon, 4, 3)   # [j south->north, i]
    cells = cells[::-1]                                            # ring north->south
    g = Grid(np.ascontiguousarray(cells.reshape(nlat * nlon, 4, 3)), SPHERICAL, None, radius, name)
    g.meta.update(kind="full_ring", nlon=nlon, nlat=nlat, shape=(nlon, nlat))
    return g


def full_gaussian_grid(nlat_half: int, radius: float = 1.0) -> Grid:
    """FullGaussianGrid F<nlat_half>: 2*nlat_half rings of 4*nlat_half points."""
    return full_ring_grid(gaussian_latitudes(2 * nlat_half), 4 * nlat_half, 0.0, radius,
                          f"F{nlat_half}")


def full_clenshaw_grid(nlat_half: int, radius: float = 1.0) -> Grid:
    """FullClenshawGrid: 2*nlat_half-1 equi-spaced rings incl. the equator, 4*nlat_half
    points per ring (test/usecases/fullclenshaw.jl)."""
    nlat = 2 * nlat_half - 1
    latd = 90.0 - 90.0 * (np.arange(1, nlat + 1) / nlat_half)
    return full_ring_grid(latd, 4 * nlat_half, 0.0, radius, f"FullClenshaw{nlat_half}")


def octahedral_gaussian_grid(nlat_half: int, radius: float = 1.0) -> Grid:
    """Octahedral Gaussian O<nlat_half>: ring j (from either pole) has 16+4j points,
    first point at lon 0.  The reference has NO implementation for reduced RingGrids
    (RingGridsExt.jl:18-20); cells here generalise the full-grid rule: latitude band
    between pole-pinned mid-latitudes x longitude interval centred on the point."""
    nlat = 2 * nlat_half
    latd = gaussian_latitudes(nlat)
    lat_edges = _pole_pinned_lat_edges(latd)
    cells = []
    for r in range(nlat):
        j = r + 1 if r < nlat_half else nlat - r
        n = 16 + 4 * j
        dlon = 360.0 / n
        lon_w = -dlon / 2 + np.arange(n) * dlon
        lon_e = lon_w + dlon
        top, bot = lat_edges[r], lat_edges[r + 1]
        c = np.empty((n, 4, 3))
        c[:, 0] = unit_sphere_from_geographic(lon_w, np.full(n, bot))
        c[:, 1] = unit_sphere_from_geographic(lon_e, np.full(n, bot))
        c[:, 2] = unit_sphere_from_geographic(lon_e, np.full(n, top))
        c[:, 3] = unit_sphere_from_geographic(lon_w, np.full(n, top))
        cells.append(c)
    g = Grid(np.ascontiguousarray(np.concatenate(cells)), SPHERICAL, None, radius, f"O{nlat_half}")
    g.meta.update(kind="octahedral", nlat_half=nlat_half)
    return g


# ----------------------------------------------------------------------------
# tripolar fold (Oceananigans RightCenterFolded)
# ----------------------------------------------------------------------------

def fold_row_slots(nx: int):
    """0-based local indices along the fold row of a ``RightCenterFolded`` grid: ``(real, partner)`` with
    ``partner[k] = nx - 1 - real[k]`` -- the reference's partition: real locals 1..Nq and Nh+1..Nh+Nq (1-based),
    the partner of r is Nx+1-r (ext/ConservativeRegriddingOceananigansExt.jl:119-187,216-240)."""
    nh, nq = nx // 2, nx // 4
    real = np.concatenate([np.arange(nq), nh + np.arange(nq)]).astype(np.int64)
    return real, nx - 1 - real


def tripolar_fold_grid(nx: int, ny: int, lat_south: float = -80.0, lat_fold: float = 84.0, radius: float = 1.0) -> Grid:
    """Synthetic stand-in of an Oceananigans ``TripolarGrid`` with a ``RightCenterFolded`` north row, cells exactly as
    ``Trees.treeify`` makes them (OceananigansExt.jl:76-187): an (nx+1) x (ny+1) vertex matrix, cell (i, j) at field
    index i + j nx; rows 0 .. ny-2 are lon-lat bands from ``lat_south`` to ``lat_fold``; the LAST row covers the cap
    north of ``lat_fold`` by strips across it, vertex (i, ny) = vertex (nx - i, ny - 1), so that fold-row cell i is the
    same physical quadrilateral as cell nx - 1 - i -- the fold.  Of every such pair one slot carries the polygon and
    its partner is a ghost: a degenerate ring of four equal points (zero area, no overlaps), like the reference's
    ``PaddedTreeWrapper``.  ``regrid!`` then copies each primary's value into its partner
    (:func:`mirror_fold_partners`)."""
    assert nx % 4 == 0 and ny >= 2
    lon = 360.0 * np.arange(nx + 1) / nx
    lat = lat_south + (lat_fold - lat_south) * np.arange(ny) / (ny - 1)        # vertex rows 0 .. ny-1
    P = np.empty((nx + 1, ny + 1, 3))
    P[:, :ny] = unit_sphere_from_geographic(lon[:, None], lat[None, :])
    P[:, ny] = P[::-1, ny - 1]                                                 # the fold
    cells = cells_from_vertex_matrix(P)                                        # [ny * nx, 4, 3], i fastest
    real, partner = fold_row_slots(nx)
    ghost = np.ones(nx, dtype=bool); ghost[real] = False
    base = (ny - 1) * nx
    cells[base + np.nonzero(ghost)[0]] = P[0, ny - 1]                          # ghost_polygon: p, p, p, p
    g = Grid(np.ascontiguousarray(cells), SPHERICAL, None, radius, f"tripolar{nx}x{ny}")
    g.meta.update(kind="tripolar_fold", shape=(nx, ny), fold=(nx, ny))
    return g


def mirror_fold_partners(field, nx: int, ny: int, axis: int = 0):
    """``mirror_fold_partners!`` (OceananigansExt.jl:216-240) on a host array: along ``axis`` (the cells) the value of
    every primary of the fold row is copied into its partner slot.  In place; returns ``field``."""
    real, partner = fold_row_slots(nx)
    base = (ny - 1) * nx
    v = np.moveaxis(field, axis, 0)
    v[base + partner] = v[base + real]
    return field


# ----------------------------------------------------------------------------
# cell centres (for sampling analytic fields)
# ----------------------------------------------------------------------------

def cell_centers_lonlat(grid: Grid):
    """(lon, lat) in degrees of the normalised vertex mean of each spherical cell."""
    assert grid.manifold == SPHERICAL
    if grid.offsets is None:
        c = grid.verts.mean(axis=1)
    else:
        counts = np.diff(grid.offsets)
        c = np.add.reduceat(grid.verts, grid.offsets[:-1], axis=0) / counts[:, None]
    c = c / np.linalg.norm(c, axis=-1, keepdims=True)
    return geographic_from_unit_sphere(c)

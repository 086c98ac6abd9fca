"""Python mirror of the reference's public API for the two hot paths, on top of the C ABI.

=====================================  ====================================================
reference (Julia)                      here
=====================================  ====================================================
``Regridder(dst, src; normalize, …)``  :func:`Regridder` / :class:`RegridderB200`
                                       (src/regridder/regridder.jl:105-163)
``regrid!(dst, R, src; dims, …)``      :func:`regrid_` (src/regridder/regrid.jl:63-118,205-318)
``regrid(R, src)``                     :func:`regrid`  (regrid.jl:322-330)
``transpose(R)``                       :func:`transpose` / ``R.T`` (regridder.jl:49-50) -- shares
                                       every array with ``R`` (``is``), flips a flag
``LinearAlgebra.normalize!(R)``        :func:`normalize_` (regridder.jl:54-62)
``R.intersections``                    :class:`B200Matrix` (device CSR(A) + CSR(A^T) handle)
``R.dst_areas / R.src_areas``          numpy vectors (geometric cell areas, regridder.jl:165-178)
``R.dst_temp / R.src_temp``            numpy work vectors for non-contiguous fields (:155-156)
=====================================  ====================================================

Julia is not installed in this image, so this module is the executable host side; the
``ccall`` binding a Julia maintainer would add is in INTEGRATION.md / julia/CRGB200.jl.
Fields may be numpy arrays (host; copied through the device inside the call) or CUDA
``torch`` tensors (device; zero-copy).  Indices are 0-based here, ``dims`` is the 0-based
axis (the reference's ``dims`` is 1-based; ``dims=0`` here == ``dims=1`` there).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib
from .grids import (Grid, GridSpec, PLANAR, SPHERICAL, cells_from_vertex_matrix, planar_regular_grid,
                    polygons_grid)


class DimensionMismatch(ValueError):
    """Julia's ``DimensionMismatch`` (regrid.jl:285-298)."""


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


# -----------------------------------------------------------------------------------------
# Trees.treeify equivalent: anything -> flat cell list in field-linear order
# -----------------------------------------------------------------------------------------

def as_grid(obj, manifold: Optional[int] = None, radius: float = 1.0) -> Grid:
    """``Trees.treeify(manifold, x)`` + ``collect(getcell(tree))`` (src/trees/interfaces.jl:94-129).

    Accepts a :class:`Grid`; an ``(x, y)`` tuple of 1-D vectors (``RegularGrid``, planar); an
    ``(nx+1, ny+1, dim)`` array of corner points (``CellBasedGrid``); an ``(nx, ny)`` object array /
    nested list of polygons; or a flat iterable of polygons (``FlatNoTree``)."""
    if isinstance(obj, (Grid, GridSpec)):
        return obj
    if isinstance(obj, tuple) and len(obj) == 2 and np.ndim(obj[0]) == 1 and np.ndim(obj[1]) == 1 \
            and np.asarray(obj[0]).dtype.kind in "fiu":
        return planar_regular_grid(obj[0], obj[1])
    if isinstance(obj, np.ndarray) and obj.dtype.kind == "f" and obj.ndim == 3 and obj.shape[2] in (2, 3):
        mf = SPHERICAL if obj.shape[2] == 3 else PLANAR
        return Grid(cells_from_vertex_matrix(obj.astype(np.float64)), mf, None, radius, "cellbased")
    if isinstance(obj, np.ndarray) and obj.dtype == object and obj.ndim == 2:
        # matrix of polygons: linear index i + j*nx (column-major, interfaces.jl:236-243)
        polys = [obj[i, j] for j in range(obj.shape[1]) for i in range(obj.shape[0])]
        return polygons_grid(polys, PLANAR if manifold is None else manifold, radius)
    polys = list(obj)
    if not polys:
        raise TypeError("cannot treeify an empty iterable")
    dim = np.asarray(polys[0]).shape[-1]
    mf = manifold if manifold is not None else (SPHERICAL if dim == 3 else PLANAR)
    return polygons_grid(polys, mf, radius)


_KINDS = {"lonlat": _lib.GRID_LONLAT, "healpix": _lib.GRID_HEALPIX, "full_ring": _lib.GRID_FULL_RING,
          "cubed_sphere": _lib.GRID_CUBED_SPHERE, "reduced_ring": _lib.GRID_REDUCED_RING}


def _grid_struct(g, keep: list) -> _lib.GridDesc:
    """``crg_grid`` of an explicit :class:`Grid` or a described :class:`GridSpec`."""
    d = _lib.GridDesc()
    if isinstance(g, GridSpec):
        d.kind = _KINDS[g.kind]
        d.flags = g.flags
        d.n1, d.n2 = g.n1, g.n2
        for i in range(4):
            d.p[i] = g.p[i]
        d.cell_lo, d.cell_hi = g.cell_lo, g.cell_hi
        if g.lat_deg is not None:
            lat = np.ascontiguousarray(g.lat_deg, dtype=np.float64)
            keep.append(lat)
            d.lat_deg = lat.ctypes.data
    else:
        d.kind = _lib.GRID_CELLS
        d.cells = _cells_struct(g, keep)
    return d


def grid_cells(spec: GridSpec, device: Optional[int] = None, out=None):
    """Vertices [ncells, 4, 3] of a described grid as generated on the device (``crg_grid_cells``).
    ``out`` may be a CUDA torch tensor (filled in place, zero copy) or None (numpy array)."""
    keep = []
    d = _grid_struct(spec, keep)
    if out is None:
        out = np.empty((spec.ncells, 4, 3), dtype=np.float64)
    _lib.check(_lib.lib().crg_grid_cells(C.byref(d), -1 if device is None else int(device), C.c_void_p(_ptr(out))))
    return out


def _cells_struct(g: Grid, keep: list) -> _lib.Cells:
    c = _lib.Cells()
    v = g.verts
    if _is_torch(v):
        import torch
        if v.dtype != torch.float64 or not v.is_contiguous():
            v = v.to(torch.float64).contiguous()
        keep.append(v)
        c.verts = v.data_ptr()
    else:
        v = np.ascontiguousarray(v, dtype=np.float64)
        keep.append(v)
        c.verts = v.ctypes.data
    if g.offsets is not None:
        o = g.offsets
        if _is_torch(o):
            import torch
            if o.dtype != torch.int32 or not o.is_contiguous():
                o = o.to(torch.int32).contiguous()
            keep.append(o)
            c.offsets = o.data_ptr()
        else:
            o = np.ascontiguousarray(o, dtype=np.int32)
            keep.append(o)
            c.offsets = o.ctypes.data
        c.nv = 0
    else:
        c.offsets = None
        c.nv = g.nv
    c.ncells = g.ncells
    return c


# -----------------------------------------------------------------------------------------
# the matrix handle
# -----------------------------------------------------------------------------------------

CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the legacy default stream, named explicitly (0 / None = library stream)


def torch_stream_ptr(device=None) -> int:
    """``cudaStream_t`` of torch's current stream on ``device`` as the library wants it: torch reports the
    legacy default stream as 0, which the C ABI reserves for "the handle's own stream"."""
    import torch
    s = torch.cuda.current_stream(device).cuda_stream
    return s if s else CUDA_STREAM_LEGACY


class _Handle:
    """Owns one ``crg_regridder*``; freed with the last reference (Julia finalizer analogue)."""

    def __init__(self, ptr, stream=None):
        self.ptr = ptr
        self.stream = stream or None      # the cudaStream_t the handle currently runs on (None = its own)

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().crg_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class B200Matrix:
    """``R.intersections``: n_dst x n_src sparse matrix living on the device.  ``transpose`` is a
    flag flip over the same handle (regridder.jl:49-50)."""

    def __init__(self, handle: _Handle, n_dst: int, n_src: int, transposed: bool = False):
        self._h = handle
        self._n_dst = n_dst
        self._n_src = n_src
        self.transposed = transposed

    @property
    def shape(self):
        return (self._n_src, self._n_dst) if self.transposed else (self._n_dst, self._n_src)

    @property
    def nnz(self) -> int:
        nnz = C.c_int64()
        _lib.check(_lib.lib().crg_dims(self._h.ptr, None, None, C.byref(nnz)))
        return int(nnz.value)

    @property
    def T(self):
        return B200Matrix(self._h, self._n_dst, self._n_src, not self.transposed)

    def transpose(self):
        return self.T

    def tocsc(self):
        """``SparseMatrixCSC`` of this (possibly transposed) matrix as a scipy matrix."""
        import scipy.sparse as sp
        nnz = self.nnz
        colptr = np.empty(self._n_src + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        nzval = np.empty(nnz, dtype=np.float64)
        _lib.check(_lib.lib().crg_export_csc(self._h.ptr, 0, colptr.ctypes.data, rowval.ctypes.data,
                                             nzval.ctypes.data))
        A = sp.csc_matrix((nzval, rowval, colptr), shape=(self._n_dst, self._n_src))
        return A.T.tocsc() if self.transposed else A

    def tocsr(self):
        import scipy.sparse as sp
        nnz = self.nnz
        rowptr = np.empty(self._n_dst + 1, dtype=np.int64)
        colval = np.empty(nnz, dtype=np.int64)
        nzval = np.empty(nnz, dtype=np.float64)
        _lib.check(_lib.lib().crg_export_csr(self._h.ptr, 0, rowptr.ctypes.data, colval.ctypes.data,
                                             nzval.ctypes.data))
        A = sp.csr_matrix((nzval, colval, rowptr), shape=(self._n_dst, self._n_src))
        return A.T.tocsr() if self.transposed else A

    def findnz(self):
        """``SparseArrays.findnz`` (column-major order), 0-based."""
        A = self.tocsc().tocoo()
        return A.row, A.col, A.data

    def toarray(self):
        return self.tocsc().toarray()

    def maximum(self) -> float:
        """``maximum(A)`` (device max-reduce; 0 for an empty matrix)."""
        m = C.c_double(0.0)
        _lib.check(_lib.lib().crg_maximum(self._h.ptr, C.byref(m)))
        return float(m.value)

    def stats(self) -> dict:
        s = _lib.BuildStats()
        _lib.check(_lib.lib().crg_stats(self._h.ptr, C.byref(s)))
        return s.asdict()

    def candidates(self):
        n = self.stats()["n_candidates"]
        s = np.empty(n, dtype=np.int64)
        d = np.empty(n, dtype=np.int64)
        _lib.check(_lib.lib().crg_candidates(self._h.ptr, s.ctypes.data, d.ctypes.data))
        return s, d

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        """Run on a caller stream: None = the library's own stream, 0 = the legacy default stream
        (``cudaStreamLegacy``), anything else a ``cudaStream_t``.  No host synchronisation."""
        if cuda_stream_ptr is not None and cuda_stream_ptr == 0:
            cuda_stream_ptr = CUDA_STREAM_LEGACY
        if cuda_stream_ptr == self._h.stream:
            return
        _lib.check(_lib.lib().crg_set_stream(self._h.ptr, C.c_void_p(cuda_stream_ptr)))
        self._h.stream = cuda_stream_ptr

    def follow_torch_stream(self, tensor):
        """CUDA-tensor applies run on torch's current stream of the tensor's device, so that they are
        ordered after the kernels that produced the inputs and before whatever consumes the outputs."""
        self.set_stream(torch_stream_ptr(tensor.device))

    def synchronize(self):
        _lib.check(_lib.lib().crg_synchronize(self._h.ptr))

    def apply_bytes(self, K: int = 1, divide: bool = True) -> int:
        b = C.c_int64()
        _lib.check(_lib.lib().crg_apply_bytes(self._h.ptr, int(self.transposed), int(divide), K, C.byref(b)))
        return int(b.value)

    # y = M x (./ areas) -- the fused perform_regridding! + finalize_regridding!
    def apply(self, dst_ptr: int, src_ptr: int, K: int, ld_dst: int, ld_src: int, level_fastest: bool,
              divide: bool, asynchronous: bool = False):
        f = _lib.lib().crg_apply_async if asynchronous else _lib.lib().crg_apply
        _lib.check(f(self._h.ptr, int(self.transposed), int(divide), C.c_void_p(dst_ptr), C.c_void_p(src_ptr),
                     K, ld_dst, ld_src, int(level_fastest)))


# -----------------------------------------------------------------------------------------
# Regridder
# -----------------------------------------------------------------------------------------

class _Lazy:
    """A host vector materialised on first access and then shared (``is``) by R and transpose(R)."""

    def __init__(self, make):
        self._make = make
        self._val = None

    def get(self):
        if self._val is None:
            self._val = self._make()
        return self._val


class RegridderB200:
    """``Regridder{W,A,V}`` (regridder.jl:25-36).  The area vectors are fetched from the device on
    first access and the work vectors allocated on first use; transpose(R) shares the same objects."""

    def __init__(self, intersections: B200Matrix, dst_areas, src_areas, dst_temp, src_temp, dst_fold=None, src_fold=None):
        self.intersections = intersections
        # (nx, ny) of a tripolar destination / source grid with a RightCenterFolded north row: regrid! mirrors the fold
        # partners of its result (finalize_regridding! of the Oceananigans extension, OceananigansExt.jl:206-240)
        self.dst_fold, self.src_fold = dst_fold, src_fold
        wrap = lambda v: v if isinstance(v, _Lazy) else _Lazy(lambda v=v: v)  # noqa: E731
        self._dst_areas, self._src_areas = wrap(dst_areas), wrap(src_areas)
        self._dst_temp, self._src_temp = wrap(dst_temp), wrap(src_temp)

    dst_areas = property(lambda self: self._dst_areas.get())
    src_areas = property(lambda self: self._src_areas.get())
    dst_temp = property(lambda self: self._dst_temp.get())
    src_temp = property(lambda self: self._src_temp.get())

    @property
    def shape(self):
        return self.intersections.shape

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def T(self):
        return transpose(self)

    def __repr__(self):
        n2, n1 = self.shape
        return f"{n2}x{n1} RegridderB200(nnz={self.intersections.nnz})"


def transpose(R: RegridderB200) -> RegridderB200:
    """``LinearAlgebra.transpose(::Regridder)``: no copy, areas and temps swapped (regridder.jl:49-50)."""
    return RegridderB200(R.intersections.T, R._src_areas, R._dst_areas, R._src_temp, R._dst_temp, R.src_fold, R.dst_fold)


def _refresh_areas(R: RegridderB200):
    M = R.intersections
    da = R.src_areas if M.transposed else R.dst_areas
    sa = R.dst_areas if M.transposed else R.src_areas
    for a in (da, sa):
        a.flags.writeable = True
    try:
        _lib.check(_lib.lib().crg_areas(M._h.ptr, da.ctypes.data, sa.ctypes.data))
    finally:
        for a in (da, sa):
            a.flags.writeable = False


def set_areas(R: RegridderB200, dst_areas=None, src_areas=None) -> RegridderB200:
    """Replace ``R.dst_areas`` / ``R.src_areas`` (the reference lets a user edit them in place, and
    ``regrid!`` divides by them, regrid.jl:104-118): pushed to the device copy the fused division uses
    (``crg_set_areas``) and mirrored in the shared host vectors.  For transpose(R) the roles are swapped."""
    M = R.intersections
    n_out, n_in = M.shape
    vals = []
    for v, n, name in ((dst_areas, n_out, "dst_areas"), (src_areas, n_in, "src_areas")):
        if v is not None:
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != (n,):
                raise DimensionMismatch(f"{name} must have {n} entries, got {v.shape}")
        vals.append(v)
    d, s_ = (vals[1], vals[0]) if M.transposed else (vals[0], vals[1])
    _lib.check(_lib.lib().crg_set_areas(M._h.ptr, d.ctypes.data if d is not None else None,
                                        s_.ctypes.data if s_ is not None else None))
    _refresh_areas(R)
    return R


def areas_to(R: RegridderB200, dst_areas_out=None, src_areas_out=None):
    """Copy ``R.dst_areas`` / ``R.src_areas`` into caller buffers (numpy arrays or CUDA tensors --
    device to device, no host round trip).  For transpose(R) the roles are already swapped."""
    M = R.intersections
    d, s_ = (src_areas_out, dst_areas_out) if M.transposed else (dst_areas_out, src_areas_out)
    _lib.check(_lib.lib().crg_areas(M._h.ptr, C.c_void_p(_ptr(d)) if d is not None else None,
                                    C.c_void_p(_ptr(s_)) if s_ is not None else None))


def normalize_(R: RegridderB200) -> RegridderB200:
    """``LinearAlgebra.normalize!(R)``: divide A and both area vectors by maximum(A)."""
    _lib.check(_lib.lib().crg_normalize(R.intersections._h.ptr))
    _refresh_areas(R)
    return R


def scale_(R: RegridderB200, divisor: float) -> RegridderB200:
    """Divide A and both area vectors by ``divisor`` -- ``normalize!`` with a given maximum (a
    destination-sharded regridder scales every row block by the maximum over all blocks)."""
    _lib.check(_lib.lib().crg_scale(R.intersections._h.ptr, float(divisor)))
    _refresh_areas(R)
    return R


def _make_options(manifold, normalize, radius, device, area_threshold, build_transpose, keep_candidates, stream=None):
    o = _lib.Options()
    _lib.check(_lib.lib().crg_options_init(C.byref(o)))
    o.manifold = manifold
    o.normalize = int(bool(normalize))
    o.radius = float(radius)
    o.device = -1 if device is None else int(device)
    o.area_threshold = float(area_threshold)
    o.build_transpose = int(bool(build_transpose))
    o.keep_candidates = int(bool(keep_candidates))
    o.stream = stream or None
    return o


def _host_empty(n: int) -> np.ndarray:
    """Float64 host vector for device->host results; page-locked when torch's caching host allocator is
    available (a pageable 25 MB read-back costs ~3 ms, a pinned one < 1 ms)."""
    try:
        import torch
        if torch.cuda.is_available() and n > 1 << 16:
            return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(n)


def _fold_of(g):
    return tuple(g.meta["fold"]) if isinstance(g, Grid) and g.meta.get("fold") else None


def mirror_fold_partners_(field, fold, *, dims: int = 0):
    """``mirror_fold_partners!`` (OceananigansExt.jl:216-240) on a numpy array or a CUDA tensor whose axis ``dims``
    runs over the nx * ny cells of a folded grid (``fold = (nx, ny)``): in place, a tiny kernel for device fields."""
    nx, ny = fold
    lay = _layout(field, dims, nx * ny)
    if lay is None or (not _is_torch(field) and field.dtype != np.float64) or \
            (_is_torch(field) and str(field.dtype) != "torch.float64"):
        if _is_torch(field):
            import torch
            from .grids import fold_row_slots
            real, partner = fold_row_slots(nx)
            base = (ny - 1) * nx
            v = field.movedim(dims, 0)
            v[torch.as_tensor(base + partner, device=field.device)] = v[torch.as_tensor(base + real, device=field.device)]
            return field
        from .grids import mirror_fold_partners
        return mirror_fold_partners(field, nx, ny, axis=dims)
    K, ld, lf = lay
    dev, stream = -1, None
    if _is_torch(field) and field.is_cuda:
        dev, stream = field.device.index, torch_stream_ptr(field.device)
    _lib.check(_lib.lib().crg_mirror_fold_partners(C.c_void_p(_ptr(field)), nx, ny, K, ld, int(lf), dev,
                                                  C.c_void_p(stream) if stream else None))
    return field


def _wrap(ptr, n_dst, n_src, stream=None) -> RegridderB200:
    h = _Handle(ptr, stream)
    M = B200Matrix(h, n_dst, n_src)

    def fetch(which):
        def make():
            a = _host_empty(n_dst if which == 0 else n_src)
            _lib.check(_lib.lib().crg_areas(h.ptr, a.ctypes.data if which == 0 else None,
                                            a.ctypes.data if which == 1 else None))
            # the fused division of regrid! uses the DEVICE copy: in-place edits of this host vector would be
            # silently ignored, so it is read-only -- change the areas with set_areas(R, ...)
            a.flags.writeable = False
            return a
        return make
    return RegridderB200(M, _Lazy(fetch(0)), _Lazy(fetch(1)), _Lazy(lambda: np.zeros(n_dst)),
                         _Lazy(lambda: np.zeros(n_src)))


def Regridder(dst, src, **kw) -> RegridderB200:
    """``Regridder(dst, src; normalize=false, intersection_operator, threaded, ...)`` (regridder.jl:105-163); keywords
    and behaviour: see :func:`_regridder`.  A grid with a tripolar fold row (``tripolar_fold_grid``) is remembered, so
    that ``regrid!`` mirrors the fold partners of its result like the Oceananigans extension does."""
    R = _regridder(dst, src, **kw)
    R.dst_fold, R.src_fold = _fold_of(dst), _fold_of(src)
    return R


def _regridder(dst, src, *, manifold: Optional[int] = None, normalize: bool = False,
              intersection_operator: Optional[Callable] = None, threaded=True, radius: Optional[float] = None,
              device: Optional[int] = None, area_threshold: float = 0.0, build_transpose: bool = True,
              keep_candidates: bool = False, stream: Optional[int] = None, **_ignored) -> RegridderB200:
    """``Regridder(dst, src; normalize=false, intersection_operator, threaded, …)``
    (regridder.jl:105-163).  ``threaded`` is accepted and ignored (the device is the parallelism);
    a custom ``intersection_operator(src_polygon, dst_polygon) -> area`` is evaluated on the host for
    every candidate pair of the device broad phase and assembled on the device
    (intersection_areas.jl:20-27)."""
    gd = as_grid(dst, manifold)
    gs = as_grid(src, manifold)
    if gd.manifold != gs.manifold:
        raise ValueError(f"Destination and source manifolds must be the same. Got {gd.manifold} and {gs.manifold}.")
    mf = gd.manifold
    if radius is None:
        radius = gd.radius
    L = _lib.lib()
    out = C.c_void_p()
    keep = []
    if stream is None:
        # device-resident vertices: build on torch's current stream, after the kernels that wrote them
        for g in (gd, gs):
            if isinstance(g, Grid) and _is_torch(g.verts) and g.verts.is_cuda:
                stream = torch_stream_ptr(g.verts.device)
                break
    elif stream == 0:
        stream = CUDA_STREAM_LEGACY
    if isinstance(gd, GridSpec) or isinstance(gs, GridSpec):
        if intersection_operator is not None:
            gd = gd.materialize() if isinstance(gd, GridSpec) else gd
            gs = gs.materialize() if isinstance(gs, GridSpec) else gs
        else:
            o = _make_options(mf, normalize, radius, device, area_threshold, build_transpose, keep_candidates, stream)
            dd, ds = _grid_struct(gd, keep), _grid_struct(gs, keep)
            _lib.check(L.crg_build_grids(C.byref(o), C.byref(dd), C.byref(ds), C.byref(out)))
            return _wrap(out.value, gd.ncells, gs.ncells, stream)
    if mf == PLANAR and intersection_operator is None and not (_is_torch(gd.verts) or _is_torch(gs.verts)):
        # non-convex (or > CRG_MAX_VERTS-vertex) planar cells: the reference's planar operator is Foster-Hormann
        # (regridder.jl:87-94); here they are split into convex parts whose pair areas add up (decompose.py)
        from .decompose import convex_parts
        pd_, ps_ = convex_parts(gd), convex_parts(gs)
        if pd_ is not None or ps_ is not None:
            return _regridder_from_parts(gd, gs, pd_, ps_, normalize, radius, device, build_transpose, stream)
    cd = _cells_struct(gd, keep)
    cs = _cells_struct(gs, keep)
    if intersection_operator is None:
        o = _make_options(mf, normalize, radius, device, area_threshold, build_transpose, keep_candidates, stream)
        _lib.check(L.crg_build(C.byref(o), C.byref(cd), C.byref(cs), C.byref(out)))
        return _wrap(out.value, gd.ncells, gs.ncells, stream)
    # plugin path: device broad phase -> host operator per pair -> device assembly
    o = _make_options(mf, False, radius, device, 0.0, False, True)
    _lib.check(L.crg_build(C.byref(o), C.byref(cd), C.byref(cs), C.byref(out)))
    tmp = _wrap(out.value, gd.ncells, gs.ncells)
    ps, pd = tmp.intersections.candidates()
    order = np.lexsort((ps, pd))
    ps, pd = ps[order], pd[order]
    rows, cols, vals = [], [], []
    for s, d in zip(ps.tolist(), pd.tolist()):
        a = float(intersection_operator(np.asarray(gs.cell(s)), np.asarray(gd.cell(d))))
        if a > 0:
            rows.append(d); cols.append(s); vals.append(a)
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    o2 = _make_options(mf, normalize, radius, device, 0.0, build_transpose, False)
    out2 = C.c_void_p()
    _lib.check(L.crg_build_from_coo(C.byref(o2), gd.ncells, gs.ncells, len(vals), rows.ctypes.data, cols.ctypes.data,
                                    vals.ctypes.data, tmp.dst_areas.ctypes.data, tmp.src_areas.ctypes.data,
                                    C.byref(out2)))
    return _wrap(out2.value, gd.ncells, gs.ncells)


def _regridder_from_parts(gd, gs, parts_d, parts_s, normalize, radius, device, build_transpose, stream):
    """Regridder of planar grids with non-convex cells from the regridder of their convex parts."""
    pg_d, own_d = parts_d if parts_d is not None else (gd, np.arange(gd.ncells, dtype=np.int64))
    pg_s, own_s = parts_s if parts_s is not None else (gs, np.arange(gs.ncells, dtype=np.int64))
    P = Regridder(pg_d, pg_s, manifold=PLANAR, radius=radius, device=device, build_transpose=False, stream=stream)
    A = P.intersections.tocsr().tocoo()
    da = np.bincount(own_d, weights=P.dst_areas, minlength=gd.ncells)
    sa = np.bincount(own_s, weights=P.src_areas, minlength=gs.ncells)
    rows = np.ascontiguousarray(own_d[A.row], dtype=np.int64)
    cols = np.ascontiguousarray(own_s[A.col], dtype=np.int64)
    vals = np.ascontiguousarray(A.data, dtype=np.float64)
    o = _make_options(PLANAR, normalize, radius, device, 0.0, build_transpose, False, stream)
    out = C.c_void_p()
    _lib.check(_lib.lib().crg_build_from_coo(C.byref(o), gd.ncells, gs.ncells, len(vals), rows.ctypes.data,
                                             cols.ctypes.data, vals.ctypes.data, da.ctypes.data, sa.ctypes.data,
                                             C.byref(out)))
    return _wrap(out.value, gd.ncells, gs.ncells, stream)


def regridder_from_coo(n_dst, n_src, dst_idx, src_idx, areas, dst_areas, src_areas, *, normalize=False,
                       device=None, manifold=SPHERICAL) -> RegridderB200:
    """``SparseArrays.sparse(i2s, i1s, areas, n_dst, n_src)`` wrapped into a Regridder
    (intersection_areas.jl:115-121): duplicates are summed."""
    r = np.ascontiguousarray(dst_idx, dtype=np.int64)
    c = np.ascontiguousarray(src_idx, dtype=np.int64)
    v = np.ascontiguousarray(areas, dtype=np.float64)
    da = np.ascontiguousarray(dst_areas, dtype=np.float64)
    sa = np.ascontiguousarray(src_areas, dtype=np.float64)
    o = _make_options(manifold, normalize, 1.0, device, 0.0, True, False)
    out = C.c_void_p()
    _lib.check(_lib.lib().crg_build_from_coo(C.byref(o), n_dst, n_src, len(v), r.ctypes.data, c.ctypes.data,
                                             v.ctypes.data, da.ctypes.data, sa.ctypes.data, C.byref(out)))
    return _wrap(out.value, n_dst, n_src)


# -----------------------------------------------------------------------------------------
# regrid!
# -----------------------------------------------------------------------------------------

def _layout(arr, ax: int, n: int):
    """Describe an N-D field as K vectors of n cells for crg_apply without copying if possible.
    Returns (K, ld, level_fastest) or None when the memory is not expressible."""
    if _is_torch(arr):
        shape, strides = tuple(arr.shape), tuple(arr.stride())
    else:
        shape, strides = arr.shape, tuple(s // arr.itemsize for s in arr.strides)
    K = 1
    for i, s in enumerate(shape):
        if i != ax:
            K *= s
    if K == 1:
        return (1, n, False) if (strides[ax] == 1 or n <= 1) else None
    # the non-spatial axes, in C order, must collapse into one index with a single stride
    other = [(shape[i], strides[i]) for i in range(len(shape)) if i != ax and shape[i] > 1]
    if not other:
        return (1, n, False) if strides[ax] == 1 else None
    base = other[-1][1]
    run = base
    for sz, st in reversed(other):
        if st != run:
            return None
        run *= sz
    if strides[ax] == 1 and base >= n:          # each level contiguous, levels `base` apart
        return (K, base, False)
    if base == 1 and strides[ax] >= K:          # levels contiguous per cell
        return (K, strides[ax], True)
    return None


def _ptr(a) -> int:
    return a.data_ptr() if _is_torch(a) else a.ctypes.data


def regrid_(dst_field, R: RegridderB200, src_field, *, dims: int = 0, normalize: bool = True,
            asynchronous: bool = False):
    """``regrid!`` (see :func:`_regrid`) + ``mirror_fold_partners!`` when the destination grid has a tripolar fold
    row (``finalize_regridding!`` of the Oceananigans extension, OceananigansExt.jl:206-214)."""
    out = _regrid(dst_field, R, src_field, dims=dims, normalize=normalize, asynchronous=asynchronous)
    if R.dst_fold is not None:
        mirror_fold_partners_(out, R.dst_fold, dims=dims if out.ndim > 1 else 0)
    return out


def _regrid(dst_field, R: RegridderB200, src_field, *, dims: int = 0, normalize: bool = True,
            asynchronous: bool = False):
    """``regrid!(dst_field, regridder, src_field; dims, normalize)`` (regrid.jl:63-118,205-318).

    1-D dense fields go straight to the device kernel; 1-D strided views are staged through
    ``R.src_temp`` / ``R.dst_temp`` exactly like the reference; N-D arrays are regridded along axis
    ``dims`` for all other indices in ONE batched SpMM launch (the reference loops K SpMVs)."""
    for f in (dst_field, src_field):
        if not (isinstance(f, np.ndarray) or _is_torch(f)):
            raise TypeError(f"no method matching extract_arraylike(::{type(f).__name__}); "
                            "fields must be numpy arrays or CUDA torch tensors")
    M = R.intersections
    n_out, n_in = M.shape
    nd_s, nd_d = src_field.ndim, dst_field.ndim
    if _is_torch(src_field) and _is_torch(dst_field) and dst_field.is_cuda:
        M.follow_torch_stream(dst_field)
    if nd_s == 1 and nd_d == 1:
        if src_field.shape[0] != n_in or dst_field.shape[0] != n_out:
            raise DimensionMismatch(f"regridder is {n_out}x{n_in}, fields have {dst_field.shape[0]} and {src_field.shape[0]} cells")
        return _regrid_1d(dst_field, R, src_field, normalize, asynchronous)
    # ---- NDSliceLoop semantics ----
    if not (isinstance(dims, (int, np.integer)) and 0 <= dims < nd_s):
        raise ValueError(f"dims={dims} is out of range for a {nd_s}-dimensional array")
    if not (0 <= dims < nd_d):
        raise ValueError(f"dims={dims} is out of range for a {nd_d}-dimensional array")
    if nd_s != nd_d:
        raise DimensionMismatch(f"source and destination ranks must match; got source rank {nd_s} and destination rank {nd_d}")
    s_other = tuple(s for i, s in enumerate(src_field.shape) if i != dims)
    d_other = tuple(s for i, s in enumerate(dst_field.shape) if i != dims)
    if s_other != d_other:
        raise DimensionMismatch(f"source and destination non-spatial axes must match; got source axes {s_other} and destination axes {d_other}")
    if src_field.shape[dims] != n_in or dst_field.shape[dims] != n_out:
        raise DimensionMismatch(f"regridder is {n_out}x{n_in}, spatial axes have {dst_field.shape[dims]} and {src_field.shape[dims]} cells")
    ls = _layout(src_field, dims, n_in)
    ld = _layout(dst_field, dims, n_out)
    src_use, dst_use, copy_back = src_field, dst_field, False
    if _is_torch(src_field) != _is_torch(dst_field):
        raise TypeError("source and destination fields must both be numpy arrays or both CUDA tensors")
    if ls is None or ld is None or ls[2] != ld[2] or src_field.dtype != _f64_dtype(src_field) \
            or dst_field.dtype != _f64_dtype(dst_field):
        # general strided / non-Float64 case: stage through dense (K, n) Float64 copies
        src_use = _dense_levels(src_field, dims)
        dst_use = _empty_like_levels(dst_field, dims)
        ls = (src_use.shape[0], n_in, False)
        ld = (dst_use.shape[0], n_out, False)
        copy_back = True
    K = ls[0]
    if K > 0 and n_out > 0:
        M.apply(_ptr(dst_use), _ptr(src_use), K, ld[1], ls[1], ls[2], normalize, asynchronous and not copy_back)
    if copy_back:
        _scatter_levels(dst_field, dst_use, dims)
    return dst_field


def _f64_dtype(a):
    if _is_torch(a):
        import torch
        return torch.float64
    return np.dtype(np.float64)


def _dense_levels(a, ax):
    if _is_torch(a):
        import torch
        return a.movedim(ax, -1).reshape(-1, a.shape[ax]).to(torch.float64).contiguous()
    return np.ascontiguousarray(np.moveaxis(a, ax, -1).reshape(-1, a.shape[ax]), dtype=np.float64)


def _empty_like_levels(a, ax):
    K = 1
    for i, s in enumerate(a.shape):
        if i != ax:
            K *= s
    if _is_torch(a):
        import torch
        return torch.empty((K, a.shape[ax]), dtype=torch.float64, device=a.device)
    return np.empty((K, a.shape[ax]), dtype=np.float64)


def _scatter_levels(dst, dense, ax):
    if _is_torch(dst):
        view = dst.movedim(ax, -1)
        view.copy_(dense.reshape(view.shape).to(dst.dtype))
    else:
        view = np.moveaxis(dst, ax, -1)
        view[...] = dense.reshape(view.shape)


def _regrid_1d(dst, R, src, normalize, asynchronous):
    M = R.intersections
    n_out, n_in = M.shape
    if _is_torch(src) or _is_torch(dst):
        if not (_is_torch(src) and _is_torch(dst)):
            raise TypeError("source and destination fields must both be numpy arrays or both CUDA tensors")
        import torch
        s = src if (src.is_contiguous() and src.dtype == torch.float64) else src.to(torch.float64).contiguous()
        d = dst if (dst.is_contiguous() and dst.dtype == torch.float64) else torch.empty(n_out, dtype=torch.float64, device=dst.device)
        M.apply(d.data_ptr(), s.data_ptr(), 1, n_out, n_in, False, normalize, asynchronous and d is dst)
        if d is not dst:
            dst.copy_(d.to(dst.dtype))
        return dst
    # extract_source_arraylike / extract_dest_arraylike (regrid.jl:75-79,137-138,273-274)
    dense_s = src.strides[0] == src.itemsize and src.dtype == np.float64
    dense_d = dst.strides[0] == dst.itemsize and dst.dtype == np.float64
    s_like = src if dense_s else R.src_temp
    d_like = dst if dense_d else R.dst_temp
    if not dense_s:
        s_like[:] = src                      # initialize_regridding! (regrid.jl:85-88)
    M.apply(d_like.ctypes.data, s_like.ctypes.data, 1, n_out, n_in, False, normalize)
    if not dense_d:
        dst[:] = d_like                      # finalize_regridding! (regrid.jl:104-111); divide already fused
    return dst


def regrid(R: RegridderB200, src_field, **kw):
    """``regrid(regridder, src_field)``: allocates the destination (regrid.jl:322-330)."""
    n_out, _ = R.shape
    if _is_torch(src_field):
        import torch
        dst = torch.zeros(n_out, dtype=src_field.dtype, device=src_field.device)
    else:
        dst = np.zeros(n_out, dtype=np.result_type(src_field.dtype, np.float64))
    return regrid_(dst, R, src_field, **kw)


def areas(grid, manifold: Optional[int] = None, out=None, device: Optional[int] = None, stream: Optional[int] = None):
    """``areas(manifold, x, tree)`` = ``[GO.area(manifold, cell) for cell in getcell(tree)]`` for one grid
    (regridder.jl:165-178), computed on the device (``crg_grid_areas``).  ``out``: numpy array or CUDA tensor to
    fill (default: a new numpy vector)."""
    g = as_grid(grid, manifold)
    keep = []
    d = _grid_struct(g, keep)
    if out is None:
        out = np.empty(g.ncells, dtype=np.float64)
    o = _make_options(g.manifold, False, g.radius, device, 0.0, False, False, stream)
    _lib.check(_lib.lib().crg_grid_areas(C.byref(o), C.byref(d), C.c_void_p(_ptr(out))))
    return out


def clip_pairs(dst, src, src_idx, dst_idx, *, manifold: Optional[int] = None, radius: Optional[float] = None,
               device: Optional[int] = None, area_threshold: float = 0.0) -> np.ndarray:
    """``compute_intersection_areas`` (intersection_areas.jl:4-32) with the default operator for an explicit
    list of 0-based (src, dst) pairs, run by the kernels of the build (``crg_clip_pairs``): the area of every
    pair (x radius^2), 0 where the pair does not survive ``area > 0``."""
    gd, gs = as_grid(dst, manifold), as_grid(src, manifold)
    if isinstance(gd, GridSpec):
        gd = gd.materialize()
    if isinstance(gs, GridSpec):
        gs = gs.materialize()
    if gd.manifold != gs.manifold:
        raise ValueError(f"Destination and source manifolds must be the same. Got {gd.manifold} and {gs.manifold}.")
    keep = []
    cd, cs = _cells_struct(gd, keep), _cells_struct(gs, keep)
    si = np.ascontiguousarray(src_idx, dtype=np.int64)
    di = np.ascontiguousarray(dst_idx, dtype=np.int64)
    if si.shape != di.shape or si.ndim != 1:
        raise DimensionMismatch("src_idx and dst_idx must be 1-D arrays of the same length")
    out = np.zeros(si.shape[0], dtype=np.float64)
    o = _make_options(gd.manifold, False, gd.radius if radius is None else radius, device, area_threshold, False, False)
    _lib.check(_lib.lib().crg_clip_pairs(C.byref(o), C.byref(cd), C.byref(cs), si.shape[0], si.ctypes.data,
                                         di.ctypes.data, out.ctypes.data))
    return out

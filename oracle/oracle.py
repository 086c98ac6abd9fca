"""Python face of the CPU ORACLE (oracle/crg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs -- never by the product
package.  See the header of crg_oracle.c for the reference citations and the parity
status ("parity unpinned" for per-entry spherical values; pinned on the planar KATs and
the spherical invariants of the reference's own tests).

The grid arguments are ``conservativeregridding.jl_b200.grids.Grid`` objects (plain
numpy containers).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "crg_oracle.c")
_LIB = os.path.join(_HERE, "libcrg_oracle.so")

_lib = None

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """gcc recipe for the oracle (also driven by oracle/Makefile)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=c11",
               "-o", _LIB, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_sph_polygon_area.restype = C.c_double
        L.orc_sph_polygon_area.argtypes = [c_f64p, C.c_int]
        L.orc_planar_polygon_area.restype = C.c_double
        L.orc_planar_polygon_area.argtypes = [c_f64p, C.c_int]
        L.orc_intersection_area.restype = C.c_double
        L.orc_intersection_area.argtypes = [C.c_int, c_f64p, C.c_int, c_f64p, C.c_int]
        L.orc_sph_clip.restype = C.c_int
        L.orc_sph_clip.argtypes = [c_f64p, C.c_int, c_f64p, C.c_int, c_f64p]
        L.orc_planar_clip.restype = C.c_int
        L.orc_planar_clip.argtypes = [c_f64p, C.c_int, c_f64p, C.c_int, c_f64p]
        L.orc_cell_areas.restype = None
        L.orc_cell_areas.argtypes = [C.c_int, c_f64p, c_i32p, C.c_int64, C.c_int, c_f64p]
        L.orc_compute_intersection_areas.restype = C.c_int64
        L.orc_compute_intersection_areas.argtypes = [
            C.c_int, c_f64p, c_i32p, C.c_int64, C.c_int, c_f64p, c_i32p, C.c_int64, C.c_int,
            c_i64p, c_i64p, C.c_int64, C.c_int, c_i64p, c_i64p, c_f64p]
        L.orc_coo_to_csc.restype = C.c_int64
        L.orc_coo_to_csc.argtypes = [C.c_int64, C.c_int64, C.c_int64, c_i64p, c_i64p, c_f64p,
                                     c_i64p, c_i64p, c_f64p]
        L.orc_normalize.restype = None
        L.orc_normalize.argtypes = [C.c_int64, c_f64p, C.c_int64, c_f64p, C.c_int64, c_f64p]
        for f in (L.orc_csc_mul, L.orc_csc_tmul):
            f.restype = None
            f.argtypes = [C.c_int64, C.c_int64, c_i64p, c_i64p, c_f64p, c_f64p, c_f64p, c_f64p, C.c_int]
        L.orc_cell_caps.restype = None
        L.orc_cell_caps.argtypes = [c_f64p, c_i32p, C.c_int64, C.c_int, c_f64p]
        L.orc_dual_query.restype = C.c_int64
        L.orc_dual_query.argtypes = [C.c_int, C.c_int, C.c_int] + \
            [C.c_int64, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_f64p, c_f64p] * 2 + \
            [C.POINTER(c_i64p), C.POINTER(c_i64p)]
        L.orc_structured_node_caps.restype = None
        L.orc_structured_node_caps.argtypes = [c_f64p, C.c_int64, C.c_int64, c_i64p, C.c_int64, c_f64p,
                                               C.c_int]
        L.orc_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = None
        _lib = L
    return _lib


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p)


def _grid_args(g):
    v, vp = _f64(g.verts)
    if g.offsets is not None:
        o = np.ascontiguousarray(g.offsets, dtype=np.int32)
        return (v, o), (vp, o.ctypes.data_as(c_i32p), C.c_int64(g.ncells), C.c_int(0))
    return (v,), (vp, None, C.c_int64(g.ncells), C.c_int(g.nv))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def use_all_cores() -> int:
    """Let OpenMP use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_threads(C.c_int(n))
    return max_threads()


# ----------------------------------------------------------------------------
# per-pair / per-cell arithmetic
# ----------------------------------------------------------------------------

def intersection_area(manifold: int, p1, p2) -> float:
    """``DefaultIntersectionOperator(manifold)(p1, p2)`` on the unit sphere / plane."""
    a, ap = _f64(p1)
    b, bp = _f64(p2)
    return float(lib().orc_intersection_area(manifold, ap, a.shape[0], bp, b.shape[0]))


def polygon_area(manifold: int, p) -> float:
    a, ap = _f64(p)
    f = lib().orc_sph_polygon_area if manifold else lib().orc_planar_polygon_area
    return float(f(ap, a.shape[0]))


def clip(manifold: int, subj, clipper) -> np.ndarray:
    a, ap = _f64(subj)
    b, bp = _f64(clipper)
    dim = 3 if manifold else 2
    out = np.zeros((64, dim))
    f = lib().orc_sph_clip if manifold else lib().orc_planar_clip
    m = f(ap, a.shape[0], bp, b.shape[0], out.ctypes.data_as(c_f64p))
    return out[:m].copy()


def cell_areas(grid) -> np.ndarray:
    """``areas(manifold, tree)`` * 1/R^2 -> scaled by R^2 here (regridder.jl:165-178)."""
    keep, args = _grid_args(grid)
    out = np.empty(grid.ncells)
    lib().orc_cell_areas(grid.manifold, args[0], args[1], args[2], args[3], out.ctypes.data_as(c_f64p))
    if grid.manifold:
        out *= grid.radius ** 2
    return out


def cell_caps(grid) -> np.ndarray:
    keep, args = _grid_args(grid)
    out = np.empty((grid.ncells, 4))
    lib().orc_cell_caps(args[0], args[1], args[2], args[3], out.ctypes.data_as(c_f64p))
    return out


# ----------------------------------------------------------------------------
# a deliberately simple, independent broad phase (superset of overlapping pairs)
# ----------------------------------------------------------------------------

def _bounding_circles(grid):
    """(centre, radius) of a ball (chord metric) that contains every cell."""
    if grid.manifold:
        caps = cell_caps(grid)
        return caps[:, :3], 2.0 * np.sin(np.minimum(caps[:, 3], np.pi) / 2.0) * (1 + 1e-9) + 1e-12
    if grid.offsets is None:
        lo = grid.verts.min(axis=1)
        hi = grid.verts.max(axis=1)
    else:
        lo = np.minimum.reduceat(grid.verts, grid.offsets[:-1], axis=0)
        hi = np.maximum.reduceat(grid.verts, grid.offsets[:-1], axis=0)
    c = 0.5 * (lo + hi)
    r = 0.5 * np.linalg.norm(hi - lo, axis=1)
    return c, r * (1 + 1e-9) + 1e-12 * (1.0 + np.abs(c).max())


def candidate_pairs_safe(dst, src):
    """All (src, dst) pairs whose bounding balls intersect (0-based int64 arrays).
    Independent of the device broad phase and of the dual DFS below."""
    from scipy.spatial import cKDTree
    cd, rd = _bounding_circles(dst)
    cs, rs = _bounding_circles(src)
    if dst.ncells * src.ncells <= 4_000_000:
        dd = np.linalg.norm(cd[:, None, :] - cs[None, :, :], axis=-1)
        ok = dd <= rd[:, None] + rs[None, :]
        d_idx, s_idx = np.nonzero(ok)
        return s_idx.astype(np.int64), d_idx.astype(np.int64)
    tree = cKDTree(cs)
    # bucket the destination cells by radius so that a few huge cells do not inflate every query
    out_s, out_d = [], []
    rs_max = float(rs.max())
    lists = tree.query_ball_point(cd, rd + rs_max, workers=-1)
    lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    d_idx = np.repeat(np.arange(dst.ncells, dtype=np.int64), lens)
    s_idx = np.fromiter((j for l in lists for j in l), dtype=np.int64, count=int(lens.sum()))
    dd = np.linalg.norm(cd[d_idx] - cs[s_idx], axis=-1)
    ok = dd <= rd[d_idx] + rs[s_idx]
    return s_idx[ok], d_idx[ok]


# ----------------------------------------------------------------------------
# Regridder build + apply
# ----------------------------------------------------------------------------

@dataclass
class OracleRegridder:
    """Mirror of ``Regridder{W,A,V}`` (regridder.jl:25-36) with a 0-based CSC matrix."""
    n_dst: int
    n_src: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray
    dst_areas: np.ndarray
    src_areas: np.ndarray
    n_candidates: int = 0

    @property
    def nnz(self):
        return int(self.nzval.shape[0])

    def tocsc(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval, self.colptr), shape=(self.n_dst, self.n_src))

    def regrid(self, src_field, transpose=False, normalize=True):
        """``regrid!`` on a dense vector: mul! + divide (regrid.jl:63-118)."""
        x, xp = _f64(src_field)
        n_out = self.n_src if transpose else self.n_dst
        y = np.empty(n_out)
        areas = self.src_areas if transpose else self.dst_areas
        a, ap = _f64(areas)
        cp, cpp = _i64(self.colptr)
        rv, rvp = _i64(self.rowval)
        nz, nzp = _f64(self.nzval)
        f = lib().orc_csc_tmul if transpose else lib().orc_csc_mul
        f(self.n_dst, self.n_src, cpp, rvp, nzp, xp, y.ctypes.data_as(c_f64p), ap, int(normalize))
        return y


def compute_intersection_areas(dst, src, pair_src, pair_dst, nthreads=1):
    """intersection_areas.jl:4-32 on the unit sphere / plane; returns (src, dst, area)."""
    kd, ad = _grid_args(dst)
    ks, as_ = _grid_args(src)
    ps, psp = _i64(pair_src)
    pd, pdp = _i64(pair_dst)
    n = ps.shape[0]
    o1 = np.empty(max(n, 1), dtype=np.int64)
    o2 = np.empty(max(n, 1), dtype=np.int64)
    oa = np.empty(max(n, 1))
    m = lib().orc_compute_intersection_areas(
        dst.manifold, ad[0], ad[1], ad[2], ad[3], as_[0], as_[1], as_[2], as_[3], psp, pdp, n, nthreads,
        o1.ctypes.data_as(c_i64p), o2.ctypes.data_as(c_i64p), oa.ctypes.data_as(c_f64p))
    return o1[:m], o2[:m], oa[:m]


def coo_to_csc(n_rows, n_cols, rows, cols, vals):
    """``SparseArrays.sparse(I, J, V, m, n)`` (duplicates summed, sorted rows)."""
    r, rp = _i64(rows)
    c, cp = _i64(cols)
    v, vp = _f64(vals)
    nnz = r.shape[0]
    colptr = np.zeros(n_cols + 1, dtype=np.int64)
    rowval = np.empty(max(nnz, 1), dtype=np.int64)
    nzval = np.empty(max(nnz, 1))
    m = lib().orc_coo_to_csc(n_rows, n_cols, nnz, rp, cp, vp, colptr.ctypes.data_as(c_i64p),
                             rowval.ctypes.data_as(c_i64p), nzval.ctypes.data_as(c_f64p))
    return colptr, rowval[:m].copy(), nzval[:m].copy()


def _ghost_cells(grid):
    """Mask of the cells whose vertices all coincide (fixed-size rings), or None when there are none."""
    if grid.offsets is not None or grid.ncells == 0:
        return None
    g = (grid.verts == grid.verts[:, :1]).all(axis=(1, 2))
    return g if g.any() else None


def build_regridder(dst, src, normalize=False, candidates=None, nthreads=1) -> OracleRegridder:
    """``Regridder(manifold, dst, src; normalize)`` (regridder.jl:125-163)."""
    assert dst.manifold == src.manifold
    if candidates is None:
        candidates = candidate_pairs_safe(dst, src)
    ps, pd = candidates
    # ghost cells (a ring of equal points: the padding polygon of a tripolar fold row) are not in the reference's
    # spatial tree and hence never candidates (ext/ConservativeRegriddingOceananigansExt.jl:66-75,141-160)
    gd, gs = _ghost_cells(dst), _ghost_cells(src)
    if gd is not None or gs is not None:
        keep = np.ones(len(ps), dtype=bool)
        if gd is not None:
            keep &= ~gd[pd]
        if gs is not None:
            keep &= ~gs[ps]
        ps, pd = ps[keep], pd[keep]
    i1, i2, a = compute_intersection_areas(dst, src, ps, pd, nthreads)
    r2 = dst.radius ** 2 if dst.manifold else 1.0
    a = a * r2
    colptr, rowval, nzval = coo_to_csc(dst.ncells, src.ncells, i2, i1, a)
    da = cell_areas(dst)
    sa = cell_areas(src)
    if dst.manifold:
        sa = sa / src.radius ** 2 * r2      # one manifold (one radius) per regridder
    R = OracleRegridder(dst.ncells, src.ncells, colptr, rowval, nzval, da, sa, len(ps))
    if normalize and R.nnz:
        lib().orc_normalize(R.nnz, R.nzval.ctypes.data_as(c_f64p), R.n_dst,
                            R.dst_areas.ctypes.data_as(c_f64p), R.n_src,
                            R.src_areas.ctypes.data_as(c_f64p))
    return R


# ----------------------------------------------------------------------------
# the reference's broad phase, restated: implicit quadtrees + bounding caps + dual DFS
# ----------------------------------------------------------------------------

FULL_SPHERE_CAP = np.array([0.0, 0.0, 1.0, np.nextafter(np.pi, 4.0)])


@dataclass
class OracleTree:
    child_lo: np.ndarray
    child_hi: np.ndarray
    leaf_lo: np.ndarray
    leaf_hi: np.ndarray
    leaf_cells: np.ndarray
    node_ext: np.ndarray
    cell_ext: np.ndarray
    manifold: int

    def cargs(self):
        keep = [np.ascontiguousarray(a, dtype=np.int64) for a in
                (self.child_lo, self.child_hi, self.leaf_lo, self.leaf_hi, self.leaf_cells)]
        ne = np.ascontiguousarray(self.node_ext, dtype=np.float64)
        ce = np.ascontiguousarray(self.cell_ext, dtype=np.float64)
        args = [C.c_int64(len(self.child_lo))] + [k.ctypes.data_as(c_i64p) for k in keep] + \
               [ne.ctypes.data_as(c_f64p), ce.ctypes.data_as(c_f64p)]
        return keep + [ne, ce], args


def _cell_extents(grid):
    if grid.manifold:
        return cell_caps(grid)
    if grid.offsets is None:
        lo = grid.verts.min(axis=1)
        hi = grid.verts.max(axis=1)
    else:
        lo = np.minimum.reduceat(grid.verts, grid.offsets[:-1], axis=0)
        hi = np.maximum.reduceat(grid.verts, grid.offsets[:-1], axis=0)
    return np.stack([lo[:, 0], hi[:, 0], lo[:, 1], hi[:, 1]], axis=1)


def flat_tree(grid) -> OracleTree:
    """``FlatNoTree`` (src/trees/interfaces.jl:280-296): one leaf holding every cell."""
    ce = _cell_extents(grid)
    if grid.manifold:
        root = FULL_SPHERE_CAP.copy()
    else:
        root = np.array([ce[:, 0].min(), ce[:, 1].max(), ce[:, 2].min(), ce[:, 3].max()])
    z = np.zeros(1, dtype=np.int64)
    return OracleTree(z, z.copy(), z.copy(), np.array([grid.ncells], dtype=np.int64),
                      np.arange(grid.ncells, dtype=np.int64), root[None, :], ce, grid.manifold)


def structured_tree(grid, nx: int, ny: int, field_index=None, full_sphere=True,
                    cell_offset: int = 0) -> OracleTree:
    """``TopDownQuadtreeCursor`` over a CellBasedGrid (quadtree_cursors.jl:214-316): halve the
    index ranges until both are <= 2 long; node extent = ``cell_range_extent``
    (grids.jl:245-287).  ``field_index[j, i]`` maps cartesian cells to field-linear indices
    (``Reorderer2D``); default i + j*nx.  Cells of ``grid`` [cell_offset, cell_offset+nx*ny)
    must be stored i-fastest in cartesian order *or* be addressed through field_index."""
    if field_index is None:
        field_index = cell_offset + np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    # --- node ranges, breadth first ---
    ranges = [(0, nx, 0, ny)]
    child_lo, child_hi = [], []
    k = 0
    while k < len(ranges):
        i0, i1, j0, j1 = ranges[k]
        li, lj = i1 - i0, j1 - j0
        if li <= 2 and lj <= 2:
            child_lo.append(0); child_hi.append(0)
        else:
            child_lo.append(len(ranges))
            if li == 1:
                s = lj // 2
                ranges += [(i0, i1, j0, j0 + s), (i0, i1, j0 + s, j1)]
            elif lj == 1:
                s = li // 2
                ranges += [(i0, i0 + s, j0, j1), (i0 + s, i1, j0, j1)]
            else:
                si, sj = li // 2, lj // 2
                ranges += [(i0, i0 + si, j0, j0 + sj), (i0, i0 + si, j0 + sj, j1),
                           (i0 + si, i1, j0, j0 + sj), (i0 + si, i1, j0 + sj, j1)]
            child_hi.append(len(ranges))
        k += 1
    R = np.array(ranges, dtype=np.int64)
    child_lo = np.array(child_lo, dtype=np.int64)
    child_hi = np.array(child_hi, dtype=np.int64)
    nn = len(R)
    is_leaf = child_hi == child_lo
    counts = np.where(is_leaf, (R[:, 1] - R[:, 0]) * (R[:, 3] - R[:, 2]), 0)
    leaf_hi = np.cumsum(counts)
    leaf_lo = leaf_hi - counts
    leaf_cells = np.empty(int(leaf_hi[-1]), dtype=np.int64)
    for n in np.nonzero(is_leaf)[0]:
        i0, i1, j0, j1 = R[n]
        # child_indices_extents iterates i fastest (quadtree_cursors.jl:233-237)
        leaf_cells[leaf_lo[n]:leaf_hi[n]] = field_index[j0:j1, i0:i1].reshape(-1)
    all_ext = _cell_extents(grid)
    cell_ext = all_ext[leaf_cells]
    if grid.manifold:
        # vertex matrix from the cartesian cell array
        cart = grid.verts[field_index.reshape(-1)].reshape(ny, nx, 4, 3)
        P = np.empty((nx + 1, ny + 1, 3))
        P[:-1, :-1] = np.transpose(cart[:, :, 0], (1, 0, 2))
        P[1:, :-1] = np.transpose(cart[:, :, 1], (1, 0, 2))
        P[1:, 1:] = np.transpose(cart[:, :, 2], (1, 0, 2))
        P[:-1, 1:] = np.transpose(cart[:, :, 3], (1, 0, 2))
        node_ext = np.empty((nn, 4))
        Pc, Pp = _f64(P)
        Rc, Rp = _i64(R)
        lib().orc_structured_node_caps(Pp, nx, ny, Rp, nn, node_ext.ctypes.data_as(c_f64p), max_threads())
        if full_sphere:
            node_ext[0] = FULL_SPHERE_CAP      # KnownFullSphereExtentWrapper (wrappers.jl:49-65)
    else:
        node_ext = np.empty((nn, 4))
        # planar: extents nest, so reduce bottom-up over the (BFS-ordered) nodes
        for n in range(nn - 1, -1, -1):
            if is_leaf[n]:
                e = cell_ext[leaf_lo[n]:leaf_hi[n]]
            else:
                e = node_ext[child_lo[n]:child_hi[n]]
            node_ext[n] = (e[:, 0].min(), e[:, 1].max(), e[:, 2].min(), e[:, 3].max())
    return OracleTree(child_lo, child_hi, leaf_lo, leaf_hi, leaf_cells, node_ext, cell_ext, grid.manifold)


def healpix_tree(grid) -> OracleTree:
    """``HealpixRootNode`` / ``HealpixTreeNode`` (HealpixExt.jl:32-148): 12 base faces, 4
    children per node in nested order, caps from the 4 pixel corners at each level."""
    from crg_b200.grids import healpix_corners_nested, healpix_nest2ring, Grid, SPHERICAL
    nside = grid.meta["nside"]
    order = grid.meta["order"]
    L = int(np.log2(nside))
    level_off = [1]
    for l in range(L + 1):
        level_off.append(level_off[-1] + 12 * 4 ** l)
    nn = level_off[-1]
    child_lo = np.zeros(nn, dtype=np.int64)
    child_hi = np.zeros(nn, dtype=np.int64)
    leaf_lo = np.zeros(nn, dtype=np.int64)
    leaf_hi = np.zeros(nn, dtype=np.int64)
    node_ext = np.empty((nn, 4))
    node_ext[0] = FULL_SPHERE_CAP
    child_lo[0], child_hi[0] = 1, 13
    for l in range(L + 1):
        npx = 12 * 4 ** l
        pix = np.arange(npx, dtype=np.int64)
        corners = healpix_corners_nested(2 ** l, pix)
        node_ext[level_off[l]:level_off[l + 1]] = cell_caps(Grid(corners, SPHERICAL))
        if l < L:
            child_lo[level_off[l]:level_off[l + 1]] = level_off[l + 1] + 4 * pix
            child_hi[level_off[l]:level_off[l + 1]] = level_off[l + 1] + 4 * pix + 4
        else:
            leaf_lo[level_off[l]:level_off[l + 1]] = pix
            leaf_hi[level_off[l]:level_off[l + 1]] = pix + 1
    nest = np.arange(12 * nside * nside, dtype=np.int64)
    leaf_cells = nest if order == "nested" else healpix_nest2ring(nside, nest)
    cell_ext = node_ext[level_off[L]:level_off[L + 1]].copy()
    return OracleTree(child_lo, child_hi, leaf_lo, leaf_hi, leaf_cells, node_ext, cell_ext, 1)


def multi_tree(trees) -> OracleTree:
    """``MultiTreeWrapper`` / ``CubedSphereToplevelTree`` (wrappers.jl:140-207): a root whose
    children are the per-panel trees; leaf cell indices must already be global."""
    nn = 1 + sum(len(t.child_lo) for t in trees)
    child_lo = np.zeros(nn, dtype=np.int64); child_hi = np.zeros(nn, dtype=np.int64)
    leaf_lo = np.zeros(nn, dtype=np.int64); leaf_hi = np.zeros(nn, dtype=np.int64)
    node_ext = np.empty((nn, 4)); node_ext[0] = FULL_SPHERE_CAP
    # BFS layout requires children contiguous: put the panel roots first, then the rest of each
    # panel, remapping node ids.
    k = len(trees)
    child_lo[0], child_hi[0] = 1, 1 + k
    remaps = []
    nxt = 1 + k
    for p, t in enumerate(trees):
        n = len(t.child_lo)
        m = np.empty(n, dtype=np.int64)
        m[0] = 1 + p
        m[1:] = nxt + np.arange(n - 1)
        nxt += n - 1
        remaps.append(m)
    leaf_cells, cell_ext = [], []
    lo_off = 0
    for p, t in enumerate(trees):
        m = remaps[p]
        internal = t.child_hi > t.child_lo
        child_lo[m] = np.where(internal, m[np.minimum(t.child_lo, len(m) - 1)], 0)
        child_hi[m] = np.where(internal, child_lo[m] + (t.child_hi - t.child_lo), 0)
        leaf_lo[m] = t.leaf_lo + lo_off
        leaf_hi[m] = t.leaf_hi + lo_off
        node_ext[m] = t.node_ext
        leaf_cells.append(t.leaf_cells); cell_ext.append(t.cell_ext)
        lo_off += len(t.leaf_cells)
    return OracleTree(child_lo, child_hi, leaf_lo, leaf_hi, np.concatenate(leaf_cells),
                      node_ext, np.concatenate(cell_ext), trees[0].manifold)


def treeify(grid) -> OracleTree:
    """``Trees.treeify(manifold, grid)`` for the synthetic grid kinds of grids.py."""
    kind = grid.meta.get("kind")
    if kind in ("lonlat", "planar_regular"):
        nx, ny = grid.meta["shape"]
        full = kind == "lonlat"
        return structured_tree(grid, nx, ny, full_sphere=full)
    if kind == "full_ring":
        nx, ny = grid.meta["shape"]
        # field order: ring-major north->south; cartesian j runs south->north
        fi = (np.arange(nx * ny, dtype=np.int64).reshape(ny, nx))[::-1]
        return structured_tree(grid, nx, ny, field_index=np.ascontiguousarray(fi))
    if kind == "healpix":
        return healpix_tree(grid)
    if kind == "cubed_sphere":
        n = grid.meta["n"]
        return multi_tree([structured_tree(grid, n, n, full_sphere=False, cell_offset=p * n * n)
                           for p in range(6)])
    return flat_tree(grid)


def dual_query(src_tree: OracleTree, dst_tree: OracleTree, nthreads=1, spawn_depth=4):
    """``get_all_candidate_pairs`` (intersection_areas.jl:48-65): (src, dst) index pairs."""
    k1, a1 = src_tree.cargs()
    k2, a2 = dst_tree.cargs()
    ps = c_i64p()
    pd = c_i64p()
    n = lib().orc_dual_query(src_tree.manifold, nthreads, spawn_depth, *a1, *a2, C.byref(ps), C.byref(pd))
    s = np.ctypeslib.as_array(ps, shape=(max(n, 1),))[:n].copy()
    d = np.ctypeslib.as_array(pd, shape=(max(n, 1),))[:n].copy()
    lib().orc_free(ps)
    lib().orc_free(pd)
    return s, d


def build_regridder_reference_path(dst, src, normalize=False, nthreads=None) -> OracleRegridder:
    """The full restated reference build: treeify -> dual DFS -> per-pair areas (threaded like
    the reference) -> serial sparse() -> serial areas.  This is what bench.py times as the
    CPU baseline ("port")."""
    if nthreads is None:
        nthreads = max_threads()
    cands = dual_query(treeify(src), treeify(dst), nthreads)
    return build_regridder(dst, src, normalize=normalize, candidates=cands, nthreads=nthreads)

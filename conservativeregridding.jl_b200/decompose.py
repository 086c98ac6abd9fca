"""General planar polygons through the convex device clip.

The reference's planar operator is ``FosterHormannClipping`` (src/regridder/regridder.jl:87-94), which takes
any simple polygon; the device kernels are convex-convex Sutherland-Hodgman (and refuse non-convex rings
with ``CRG_ERR_UNSUPPORTED``).  For simple polygons the two meet through a partition: split every non-convex
(or more-than-``MAX_VERTS``-vertex) ring into triangles ``T_k`` (ear clipping), clip all part pairs on the
device and add up, ``area(P n Q) = sum_kl area(P_k n Q_l)`` -- the parts are interior-disjoint -- which is what
``crg_build_from_coo`` does with duplicate (dst, src) entries (``SparseArrays.sparse`` semantics,
intersection_areas.jl:115-121).  Cell areas are the sums of the parts' areas (= the shoelace area,
regridder.jl:165-178).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .grids import Grid, PLANAR

MAX_VERTS = 8          # CRG_MAX_VERTS of include/crg_b200.h


def _cross2(o, a, b):
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def ring_is_convex(r: np.ndarray) -> bool:
    """Every vertex on the inner side of every edge (same criterion and tolerance as the device's
    ring_is_convex, csrc/broadphase.cuh)."""
    n = len(r)
    e = np.roll(r, -1, axis=0) - r
    # h[e, i] = cross(edge_e, r_i - r_e)
    d = r[None, :, :] - r[:, None, :]
    h = e[:, None, 0] * d[:, :, 1] - e[:, None, 1] * d[:, :, 0]
    area2 = float(np.sum(r[:, 0] * np.roll(r[:, 1], -1) - np.roll(r[:, 0], -1) * r[:, 1]))
    sg = -1.0 if area2 < 0 else 1.0
    diam = float(np.sqrt(((r[None] - r[:, None]) ** 2).sum(-1).max())) if n else 0.0
    mag = np.abs(r).sum(axis=1) + np.abs(np.roll(r, -1, axis=0)).sum(axis=1)
    tol = np.linalg.norm(e, axis=1)[:, None] * (1e-9 * diam + 1e-14 * mag[:, None])
    return bool((sg * h >= -tol).all())


def triangulate(ring: np.ndarray) -> list:
    """Ear clipping of a simple polygon (any orientation, collinear vertices allowed) -> list of (3, 2)
    counter-clockwise triangles covering it exactly once."""
    r = np.asarray(ring, dtype=np.float64)
    area2 = np.sum(r[:, 0] * np.roll(r[:, 1], -1) - np.roll(r[:, 0], -1) * r[:, 1])
    if area2 < 0:
        r = r[::-1]
    idx = list(range(len(r)))
    tris = []
    guard = 0
    while len(idx) > 3 and guard < 10 * len(r) ** 2:
        guard += 1
        n = len(idx)
        found = False
        for k in range(n):
            ia, ib, ic = idx[k - 1], idx[k], idx[(k + 1) % n]
            a, b, c = r[ia], r[ib], r[ic]
            cr = _cross2(a, b, c)
            if cr < 0:
                continue                                  # reflex corner
            if cr == 0:                                   # collinear: drop the middle vertex, no triangle
                idx.pop(k); found = True
                break
            ok = True
            for j in idx:
                if j in (ia, ib, ic):
                    continue
                p = r[j]
                if _cross2(a, b, p) >= 0 and _cross2(b, c, p) >= 0 and _cross2(c, a, p) >= 0 \
                        and not (np.array_equal(p, a) or np.array_equal(p, b) or np.array_equal(p, c)):
                    ok = False
                    break
            if ok:
                tris.append(np.array([a, b, c]))
                idx.pop(k); found = True
                break
        if not found:
            raise ValueError("polygon is not simple (ear clipping found no ear)")
    if len(idx) == 3 and _cross2(r[idx[0]], r[idx[1]], r[idx[2]]) != 0:
        tris.append(r[idx])
    return tris


def convex_parts(g: Grid) -> Optional[Tuple[Grid, np.ndarray]]:
    """``None`` when every cell of the planar grid already fits the device clip (convex, at most MAX_VERTS
    vertices); otherwise ``(parts_grid, owner)``: the convex parts in cell order (convex cells stay whole) and
    the cell each part belongs to."""
    if g.manifold != PLANAR:
        return None
    v = np.asarray(g.verts)
    if g.offsets is None:
        nv = v.shape[1]
        if nv <= 3:
            return None
        # vectorised convexity of fixed-size rings
        e = np.roll(v, -1, axis=1) - v
        d = v[:, None, :, :] - v[:, :, None, :]                       # [cell, edge, vertex, 2]
        h = e[:, :, None, 0] * d[..., 1] - e[:, :, None, 1] * d[..., 0]
        area2 = np.sum(v[:, :, 0] * np.roll(v[:, :, 1], -1, axis=1) - np.roll(v[:, :, 0], -1, axis=1) * v[:, :, 1], axis=1)
        sg = np.where(area2 < 0, -1.0, 1.0)[:, None, None]
        diam = np.sqrt(((v[:, None] - v[:, :, None]) ** 2).sum(-1).max(axis=(1, 2)))
        mag = np.abs(v).sum(axis=2) + np.abs(np.roll(v, -1, axis=1)).sum(axis=2)
        tol = np.linalg.norm(e, axis=2)[:, :, None] * (1e-9 * diam[:, None, None] + 1e-14 * mag[:, :, None])
        bad = ~((sg * h >= -tol).all(axis=(1, 2)))
        if nv <= MAX_VERTS and not bad.any():
            return None
        rings = [v[i] for i in range(v.shape[0])]
        needs = bad | (nv > MAX_VERTS)
    else:
        off = np.asarray(g.offsets)
        rings = [v[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        needs = np.array([len(r) > MAX_VERTS or (len(r) > 3 and not ring_is_convex(r)) for r in rings], dtype=bool)
        if not needs.any():
            return None
    parts, owner = [], []
    for i, r in enumerate(rings):
        if needs[i]:
            for t in triangulate(r):
                parts.append(t); owner.append(i)
        else:
            parts.append(np.asarray(r, dtype=np.float64)); owner.append(i)
    sizes = {len(p) for p in parts}
    if len(sizes) == 1:
        pg = Grid(np.ascontiguousarray(np.stack(parts)), PLANAR, None, g.radius, g.name + "/parts")
    else:
        o = np.zeros(len(parts) + 1, dtype=np.int32)
        o[1:] = np.cumsum([len(p) for p in parts])
        pg = Grid(np.ascontiguousarray(np.concatenate(parts)), PLANAR, o, g.radius, g.name + "/parts")
    return pg, np.asarray(owner, dtype=np.int64)

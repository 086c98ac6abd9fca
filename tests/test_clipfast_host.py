"""The fast clip path (csrc/clipfast.cuh: wedge sum over the subject's edges, no intersection polygon) compiled for
the HOST from the same source as the CUDA kernel, checked pair by pair against the oracle's Sutherland-Hodgman clip
(oracle/crg_oracle.c) on candidate lists of several grid pairs: aligned / nested / identical grids, polar cells,
coarse cells, both size orders.  Runs without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from crg_b200 import grids
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "clipfast_host.cpp")
LIB = os.path.join(HERE, "native", "libclipfast_host.so")


@pytest.fixture(scope="module")
def host_lib():
    hdr = os.path.join(HERE, "..", "conservativeregridding.jl_b200", "csrc", "clipfast.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    L.cf_pairs_host.restype = None
    return L


def fast_areas(L, dst, src, ps, pd, swap=False, stored_vertex_as_P=False):
    """(area, kind) per candidate pair; subject = source cell, clip = destination cell (or swapped)."""
    sub, clp, si, ci = (dst, src, pd, ps) if swap else (src, dst, ps, pd)
    sa, ca = oracle.cell_areas(sub), oracle.cell_areas(clp)          # signed: clockwise cells are negative
    sf, cf = (sa < 0).astype(np.uint8), (ca < 0).astype(np.uint8)
    sv, cv = np.ascontiguousarray(sub.verts, dtype=np.float64), np.ascontiguousarray(clp.verts, dtype=np.float64)
    si, ci = np.ascontiguousarray(si, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int64)
    n = len(si)
    area, kind = np.zeros(n), np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.cf_pairs_host(p(sv), p(sf), p(cv), p(cf), p(si), p(ci), C.c_longlong(n), p(area), p(kind), C.c_int(int(stored_vertex_as_P)))
    return area, kind


PAIRS = {
    "1deg<-healpix64": (lambda: grids.lonlat_grid(360, 180), lambda: grids.healpix_grid(64, "ring"), False),
    "4deg<-2deg nested": (lambda: grids.lonlat_grid(90, 45), lambda: grids.lonlat_grid(180, 90), False),
    "healpix16 nested<-ring (identical cells)": (lambda: grids.healpix_grid(16, "nested"), lambda: grids.healpix_grid(16, "ring"), False),
    "healpix8<-healpix32 nested": (lambda: grids.healpix_grid(8, "nested"), lambda: grids.healpix_grid(32, "ring"), False),
    "10deg<-healpix4 (coarse)": (lambda: grids.lonlat_grid(36, 18), lambda: grids.healpix_grid(4, "ring"), False),
    "F24<-C12": (lambda: grids.full_gaussian_grid(24), lambda: grids.cubed_sphere_grid(12), False),
    "healpix32<-2deg, roles swapped": (lambda: grids.healpix_grid(32, "ring"), lambda: grids.lonlat_grid(180, 90), True),
}


@pytest.mark.parametrize("name", list(PAIRS))
def test_fast_path_against_oracle_clip(host_lib, name):
    dst, src, swap = PAIRS[name][0](), PAIRS[name][1](), PAIRS[name][2]
    ps, pd = oracle.candidate_pairs_safe(dst, src)
    area, kind = fast_areas(host_lib, dst, src, ps, pd, swap)
    o1, o2, oa = oracle.compute_intersection_areas(dst, src, ps, pd, nthreads=oracle.use_all_cores())
    key = ps.astype(np.int64) * dst.ncells + pd
    order = np.argsort(key)
    ref = np.zeros(len(ps))
    ref[order[np.searchsorted(key[order], o1 * dst.ncells + o2)]] = oa
    sub_area = np.abs(oracle.cell_areas(dst if swap else src))[pd if swap else ps]
    fast, inside, empty, general = kind == 1, kind == 0, kind == -1, kind == 2
    assert fast.sum() > 0.15 * len(ps), "the fast path must carry a real share of the pairs"
    # CF_EMPTY: the oracle finds nothing either (at most round-off slivers); CF_INSIDE: the whole subject cell
    tau = 1e-9 * min(np.abs(oracle.cell_areas(dst)).min(), np.abs(oracle.cell_areas(src)).min())
    assert (ref[empty] <= tau).all()
    assert np.allclose(ref[inside], sub_area[inside], rtol=1e-12, atol=0)
    assert np.allclose(area[inside], sub_area[inside], rtol=1e-12, atol=0)
    # CF_FAST: the comparator of tests/helpers.py::compare_matrices
    floor = 1e-12 * ref.max()
    err = np.abs(area[fast] - ref[fast])
    assert (err <= 1e-10 * ref[fast] + floor).all(), (err.max(), floor)
    # pattern: where one side keeps an entry and the other has none (or a non-positive area), it is a sub-tau sliver
    a, r = area[fast], ref[fast]
    assert (r[a <= 0] <= tau).all() and (a[r == 0] <= tau).all()
    assert general.sum() + fast.sum() + inside.sum() + empty.sum() == len(ps)


def test_untouched_pairs_give_exact_zero(host_lib):
    """A subject cell cut by two adjacent clip edges but lying outside the corner: every edge interval is empty and the
    sum is exactly 0.0 (idle edges multiply the accumulator by 1 + 0i), so no junk entries survive `area > 0`."""
    dst, src = grids.lonlat_grid(36, 18), grids.healpix_grid(8, "ring")
    ps, pd = oracle.candidate_pairs_safe(dst, src)
    area, kind = fast_areas(host_lib, dst, src, ps, pd)
    o1, o2, oa = oracle.compute_intersection_areas(dst, src, ps, pd)
    key = ps.astype(np.int64) * dst.ncells + pd
    hit = np.isin(key, o1 * dst.ncells + o2)
    far = (kind == 1) & ~hit
    assert far.sum() > 0
    assert (np.abs(area[far]) < 1e-18).all()          # (pairs that touch in a point or along an edge: round-off slivers)
    assert (area[far] == 0.0).mean() > 0.9

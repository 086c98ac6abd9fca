"""The CUDA clip itself against an INDEPENDENT exact / 50-digit answer, at the scale of the bench configs
("Next" 1b): tests/golden/highprec_pairs.npz holds 12 600 (source cell, destination cell) pairs
sampled from the candidate lists of BASELINE configs 5, 2 and 1 -- random, polar, slivers, non-overlapping,
HEALPix corners on lon-lat lines, nested / edge-coincident cells -- with the area of their intersection from
oracle/highprec.py::independent_intersection_area (exact integer predicates, vertex enumeration, Girard's
excess; no Sutherland-Hodgman, no Float64).  The pairs go through crg_clip_pairs, i.e. the kernels of the build."""
import os

import numpy as np
import pytest

from crg_b200 import grids
from crg_b200.regridder import clip_pairs
from helpers import GOLDEN, highprec_check, write_report

pytestmark = pytest.mark.gpu


def test_device_clip_against_independent_50_digit_areas(gpu):
    z = np.load(os.path.join(GOLDEN, "highprec_pairs.npz"))
    n = len(z["area"])
    dst = grids.Grid(np.ascontiguousarray(z["dst_verts"]), grids.SPHERICAL)
    src = grids.Grid(np.ascontiguousarray(z["src_verts"]), grids.SPHERICAL)
    k = np.arange(n, dtype=np.int64)
    got = clip_pairs(dst, src, k, k)
    rep = highprec_check(got, z)
    write_report("highprec_pairs device", rep)
    assert rep["n_beyond_tolerance"] == 0, rep
    assert rep["n_kept_above_tau_where_exact_is_zero"] == 0 and rep["n_dropped_where_exact_above_tau"] == 0, rep


def test_device_clip_pairs_on_full_grids(gpu):
    """Same pairs addressed inside the full config-2 grids (the gather path of the build)."""
    z = np.load(os.path.join(GOLDEN, "highprec_pairs.npz"))
    m = z["cfg"] == list(z["configs"]).index("cfg2")
    dst, src = grids.lonlat_grid(720, 360), grids.healpix_grid(256, "ring")
    assert np.array_equal(dst.verts[z["dst_idx"][m]], z["dst_verts"][m])       # the fixture's cells ARE these grids' cells
    assert np.array_equal(src.verts[z["src_idx"][m]], z["src_verts"][m])
    got = clip_pairs(dst, src, z["src_idx"][m], z["dst_idx"][m])
    sub = {k: (z[k][m] if k in ("area", "cat", "cfg", "min_cell_area") else z[k]) for k in z.files}
    rep = highprec_check(got, sub)
    assert rep["n_beyond_tolerance"] == 0 and rep["n_kept_above_tau_where_exact_is_zero"] == 0, rep

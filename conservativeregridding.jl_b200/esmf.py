"""ESMF offline-weights export / import.

Mirrors ``save_esmf_weights(path, regridder; ...)`` of the reference's NCDatasets extension
(/root/reference/ext/ConservativeRegriddingNCDatasetsExt.jl:15-59): variables ``S = A[row, col] /
dst_area[row]``, 1-based ``row`` (destination) / ``col`` (source) in ``findnz`` (column-major) order,
``frac_a``, ``frac_b``, ``area_a``, ``area_b``; attributes ``normalization = "destarea"`` etc.
Written with ``scipy.io.netcdf_file`` (NetCDF-3, 64-bit offsets: ``row``/``col`` are int32, which
is what ESMF itself writes).  ``load_esmf_weights`` rebuilds a device regridder from such a file --
the persistent on-disk form of a Regridder.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from .regridder import RegridderB200, regridder_from_coo


def save_esmf_weights(path: str, r: RegridderB200, *, src_grid_name: str = "source",
                      dst_grid_name: str = "destination", src_shape: Optional[Sequence[int]] = None,
                      dst_shape: Optional[Sequence[int]] = None, created_at: Optional[str] = None) -> str:
    from scipy.io import netcdf_file
    A = r.intersections.tocsc()                       # device -> host SparseMatrixCSC
    src_areas = np.asarray(r.src_areas, dtype=np.float64)
    dst_areas = np.asarray(r.dst_areas, dtype=np.float64)
    coo = A.tocoo()                                   # findnz order: column-major
    row, col, vals = coo.row, coo.col, coo.data
    S = vals / dst_areas[row]
    frac_a = np.asarray(A.sum(axis=0)).ravel() / src_areas
    frac_b = np.asarray(A.sum(axis=1)).ravel() / dst_areas
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with netcdf_file(path, "w", version=2) as ds:
        ds.createDimension("n_s", len(vals))
        ds.createDimension("n_a", len(src_areas))
        ds.createDimension("n_b", len(dst_areas))

        def put(name, data, dim, dtype, long_name):
            v = ds.createVariable(name, dtype, (dim,))
            if len(data):
                v[:] = data
            v.long_name = long_name
        put("S", S, "n_s", "d", "weight value")
        put("row", (row + 1).astype(np.int32), "n_s", "i", "destination cell index (1-based)")
        put("col", (col + 1).astype(np.int32), "n_s", "i", "source cell index (1-based)")
        put("frac_a", frac_a, "n_a", "d", "source cell fraction covered")
        put("frac_b", frac_b, "n_b", "d", "destination cell fraction covered")
        put("area_a", src_areas, "n_a", "d", "source cell areas")
        put("area_b", dst_areas, "n_b", "d", "destination cell areas")
        ds.title = "ConservativeRegridding.jl weights (ESMF format)"
        ds.created_by = "crg_b200.esmf.save_esmf_weights"
        ds.source_grid = str(src_grid_name)
        ds.destination_grid = str(dst_grid_name)
        ds.normalization = "destarea"
        ds.map_method = "Conservative remapping"
        if created_at is not None:
            ds.created_at = str(created_at)
        if src_shape is not None:
            ds.source_grid_shape = np.asarray(src_shape, dtype=np.int32)
        if dst_shape is not None:
            ds.destination_grid_shape = np.asarray(dst_shape, dtype=np.int32)
    return path


def load_esmf_weights(path: str, *, device: Optional[int] = None) -> RegridderB200:
    """Rebuild a regridder from an ESMF weight file: A[row, col] = S * area_b[row], assembled on
    the device (``crg_build_from_coo``)."""
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as ds:
        S = np.array(ds.variables["S"][:], dtype=np.float64)
        row = np.array(ds.variables["row"][:], dtype=np.int64) - 1
        col = np.array(ds.variables["col"][:], dtype=np.int64) - 1
        area_a = np.array(ds.variables["area_a"][:], dtype=np.float64)
        area_b = np.array(ds.variables["area_b"][:], dtype=np.float64)
    return regridder_from_coo(len(area_b), len(area_a), row, col, S * area_b[row], area_b, area_a, device=device)

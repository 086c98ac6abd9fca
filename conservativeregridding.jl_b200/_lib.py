"""ctypes binding of libcrgb200.so (the C ABI of include/crg_b200.h) and its nvcc recipe.

The library is built IN-TREE (``conservativeregridding.jl_b200/csrc/libcrgb200.so``) for
sm_100a only.  There is no fallback: if the library is missing or no CUDA device is usable,
every compute entry point raises :class:`CrgError`.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
# (CRG_LIB / CRG_NVCC_EXTRA: experiment hooks -- a variant of the library built with extra -D flags, see scripts/variants.py)
LIB_PATH = os.environ.get("CRG_LIB") or os.path.join(CSRC, "libcrgb200.so")
SOURCES = ["crg_b200.cu"]
HEADERS = ["common.cuh", "scan.cuh", "sort.cuh", "geom.cuh", "gridgen.cuh", "broadphase.cuh", "kernels.cuh", "sell.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]

CRG_OK, CRG_ERR_INVALID, CRG_ERR_CUDA, CRG_ERR_NOMEM, CRG_ERR_UNSUPPORTED, CRG_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
CRG_PLANAR, CRG_SPHERICAL = 0, 1

# every symbol include/crg_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "crg_options_init", "crg_build", "crg_build_from_coo", "crg_free", "crg_dims", "crg_stats", "crg_areas",
    "crg_export_csc", "crg_export_csr", "crg_candidates", "crg_normalize", "crg_maximum", "crg_scale", "crg_apply", "crg_apply_async",
    "crg_set_stream", "crg_synchronize", "crg_apply_bytes", "crg_last_error", "crg_device_count", "crg_version",
    "crg_fp64_peak", "crg_launch_count", "crg_build_grids", "crg_grid_ncells", "crg_grid_cells",
    "crg_clip_pairs", "crg_set_areas", "crg_grid_areas", "crg_mirror_fold_partners",
]


class CrgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcrgb200 error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("manifold", C.c_int32), ("normalize", C.c_int32), ("radius", C.c_double),
                ("area_threshold", C.c_double), ("device", C.c_int32), ("build_transpose", C.c_int32),
                ("keep_candidates", C.c_int32), ("reserved", C.c_int32), ("stream", C.c_void_p)]


class Cells(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("offsets", C.c_void_p), ("ncells", C.c_int64), ("nv", C.c_int32),
                ("reserved", C.c_int32)]


class GridDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("cells", Cells), ("n1", C.c_int64), ("n2", C.c_int64),
                ("p", C.c_double * 4), ("lat_deg", C.c_void_p), ("cell_lo", C.c_int64), ("cell_hi", C.c_int64)]


GRID_CELLS, GRID_LONLAT, GRID_HEALPIX, GRID_FULL_RING, GRID_CUBED_SPHERE, GRID_REDUCED_RING = 0, 1, 2, 3, 4, 5


class BuildStats(C.Structure):
    _fields_ = [("n_dst", C.c_int64), ("n_src", C.c_int64), ("n_candidates", C.c_int64), ("nnz", C.c_int64),
                ("n_bins", C.c_int64), ("n_bin_entries", C.c_int64), ("n_big_dst", C.c_int64),
                ("n_big_src", C.c_int64), ("ms_total", C.c_double), ("ms_h2d", C.c_double),
                ("ms_device", C.c_double), ("ms_bounds", C.c_double), ("ms_bin", C.c_double),
                ("ms_query", C.c_double), ("ms_clip", C.c_double), ("ms_sort_csr", C.c_double),
                ("ms_sort_csc", C.c_double), ("ms_areas", C.c_double), ("ms_finish", C.c_double),
                ("bin_size", C.c_double), ("sort_passes_csr", C.c_int32), ("sort_passes_csc", C.c_int32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "crg_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("CRG_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", INCLUDE, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    return LIB_PATH


_lib = None


def lib():
    """Load libcrgb200.so (never builds implicitly on a box without nvcc; fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        else:
            raise CrgError(CRG_ERR_NO_DEVICE, f"{LIB_PATH} is missing and nvcc is not available; "
                                              "run __graft_entry__.build() first (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    L.crg_last_error.restype = C.c_char_p
    L.crg_version.restype = C.c_char_p
    L.crg_options_init.argtypes = [P(Options)]
    L.crg_build.argtypes = [P(Options), P(Cells), P(Cells), P(vp)]
    L.crg_build_grids.argtypes = [P(Options), P(GridDesc), P(GridDesc), P(vp)]
    L.crg_grid_ncells.argtypes = [P(GridDesc), P(i64)]
    L.crg_grid_cells.argtypes = [P(GridDesc), i32, vp]
    L.crg_grid_areas.argtypes = [P(Options), P(GridDesc), vp]
    L.crg_mirror_fold_partners.argtypes = [vp, i64, i64, i64, i64, i32, i32, vp]
    L.crg_build_from_coo.argtypes = [P(Options), i64, i64, i64, vp, vp, vp, vp, vp, P(vp)]
    L.crg_free.argtypes = [vp]
    L.crg_dims.argtypes = [vp, P(i64), P(i64), P(i64)]
    L.crg_stats.argtypes = [vp, P(BuildStats)]
    L.crg_areas.argtypes = [vp, vp, vp]
    L.crg_set_areas.argtypes = [vp, vp, vp]
    L.crg_clip_pairs.argtypes = [P(Options), P(Cells), P(Cells), i64, vp, vp, vp]
    L.crg_export_csc.argtypes = [vp, i32, vp, vp, vp]
    L.crg_export_csr.argtypes = [vp, i32, vp, vp, vp]
    L.crg_candidates.argtypes = [vp, vp, vp]
    L.crg_normalize.argtypes = [vp]
    L.crg_maximum.argtypes = [vp, P(f64)]
    L.crg_scale.argtypes = [vp, f64]
    L.crg_apply.argtypes = [vp, i32, i32, vp, vp, i64, i64, i64, i32]
    L.crg_apply_async.argtypes = [vp, i32, i32, vp, vp, i64, i64, i64, i32]
    L.crg_set_stream.argtypes = [vp, vp]
    L.crg_synchronize.argtypes = [vp]
    L.crg_apply_bytes.argtypes = [vp, i32, i32, i64, P(i64)]
    L.crg_device_count.argtypes = [P(i32)]
    L.crg_fp64_peak.argtypes = [i32, P(f64)]
    L.crg_launch_count.argtypes = [P(C.c_uint64)]
    for name in EXPORTS:
        if name not in ("crg_last_error", "crg_version"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != CRG_OK:
        raise CrgError(rc, lib().crg_last_error().decode(errors="replace"))


def launch_count() -> int:
    n = C.c_uint64(0)
    check(lib().crg_launch_count(C.byref(n)))
    return int(n.value)


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().crg_device_count(C.byref(n)))
    return int(n.value)

"""CPU tests: the C-ABI library loads, exports every symbol include/crg_b200.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device.  No compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from crg_b200 import _lib, grids

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "crg_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in declared_functions():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.crg_version()


def test_struct_layouts_match_header():
    # sizes follow from the field lists in the header (natural alignment)
    assert C.sizeof(_lib.Options) == 48
    assert C.sizeof(_lib.Cells) == 32
    o = _lib.Options()
    _lib.check(_lib.lib().crg_options_init(C.byref(o)))
    assert (o.manifold, o.normalize, o.radius, o.area_threshold, o.device, o.build_transpose) == (1, 0, 1.0, 0.0, -1, 1)


def test_argument_validation_needs_no_device():
    L = _lib.lib()
    out = C.c_void_p()
    o = _lib.Options()
    L.crg_options_init(C.byref(o))
    assert L.crg_build(None, None, None, C.byref(out)) == _lib.CRG_ERR_INVALID
    c = _lib.Cells()
    v = np.zeros((2, 2, 3))
    c.verts, c.offsets, c.ncells, c.nv = v.ctypes.data, None, 2, 2      # nv = 2 is not a polygon
    assert L.crg_build(C.byref(o), C.byref(c), C.byref(c), C.byref(out)) == _lib.CRG_ERR_UNSUPPORTED
    assert b"nv=2" in L.crg_last_error()
    o.manifold = 7
    assert L.crg_build(C.byref(o), C.byref(c), C.byref(c), C.byref(out)) == _lib.CRG_ERR_INVALID
    assert L.crg_dims(None, None, None, None) == _lib.CRG_ERR_INVALID
    assert L.crg_apply(None, 0, 1, None, None, 1, 0, 0, 0) == _lib.CRG_ERR_INVALID


def test_no_cpu_fallback():
    """Without a device the product path must refuse to compute (never route to the oracle)."""
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from crg_b200.regridder import Regridder
    with pytest.raises(_lib.CrgError) as e:
        Regridder(grids.lonlat_grid(4, 2), grids.lonlat_grid(4, 2))
    assert e.value.code == _lib.CRG_ERR_NO_DEVICE
    tf = C.c_double()
    assert _lib.lib().crg_fp64_peak(-1, C.byref(tf)) == _lib.CRG_ERR_NO_DEVICE


def test_product_package_never_imports_oracle():
    pkg = os.path.join(os.path.dirname(HEADER), "..", "conservativeregridding.jl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(root, f)


# --- the Julia binding (untestable here: no Julia in the image) is at least held to the header ----------------------
JL = os.path.join(os.path.dirname(HEADER), "..", "conservativeregridding.jl_b200", "julia", "CRGB200.jl")
JL_SIZES = {"Int32": 4, "Int64": 8, "Float64": 8, "Ptr{Cvoid}": 8, "Ptr{Float64}": 8, "Ptr{Int32}": 8,
            "NTuple{4, Float64}": 32, "CrgCells": 32}
JL_ALIGN = {"NTuple{4, Float64}": 8, "CrgCells": 8}


def julia_structs():
    src = open(JL).read()
    out = {}
    for name, body in re.findall(r"^struct (Crg\w+)[^\n]*\n(.*?)^end", src, flags=re.S | re.M):
        out[name] = [(f, t.strip()) for f, t in re.findall(r"^\s+(\w+)::([^#\n]+)", body, flags=re.M)]
    return out


def layout(fields):
    off, offs = 0, []
    for _, t in fields:
        size, align = JL_SIZES[t], JL_ALIGN.get(t, JL_SIZES[t])
        off = (off + align - 1) // align * align
        offs.append((off, size))
        off += size
    return offs, (off + 7) // 8 * 8


def test_julia_struct_layouts_match_the_ctypes_mirror_and_header():
    js = julia_structs()
    for jl_name, ct in (("CrgOptions", _lib.Options), ("CrgCells", _lib.Cells), ("CrgGrid", _lib.GridDesc)):
        fields = js[jl_name]
        assert [f for f, _ in fields] == [f for f, _ in ct._fields_], jl_name          # same names, same order
        offs, total = layout(fields)
        assert total == C.sizeof(ct), (jl_name, total, C.sizeof(ct))
        for (f, _), (o, sz) in zip(fields, offs):
            d = getattr(ct, f)
            assert (d.offset, d.size) == (o, sz), (jl_name, f, (d.offset, d.size), (o, sz))
    # and the header's struct bodies list the same fields in the same order
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for c_name, ct in (("crg_options", _lib.Options), ("crg_cells", _lib.Cells), ("crg_grid", _lib.GridDesc),
                       ("crg_build_stats", _lib.BuildStats)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (c_name, c_name), hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*?\]", "", part.strip().split()[-1].lstrip("*")))
        assert names == [f for f, _ in ct._fields_], (c_name, names)


def test_julia_ccalls_name_declared_symbols_with_the_right_arity():
    src = open(JL).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    arity = {}
    for name, args in re.findall(r"\b(crg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        arity[name] = 0 if args.strip() in ("", "void") else len(args.split(","))
    calls = re.findall(r"ccall\(\(:(crg_\w+), lib\), \w+,\s*\(([^()]*)\)", src, flags=re.S)
    assert len(calls) >= 10
    for name, sig in calls:
        assert name in arity, name
        n = 0 if not sig.strip() else len([a for a in re.split(r",(?![^{]*\})", sig) if a.strip()])
        assert n == arity[name], (name, n, arity[name], sig)

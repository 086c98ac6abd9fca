# CRGB200.jl -- the reference-side binding of libcrgb200.so (include/crg_b200.h).
#
# UNTESTED IN THIS REPOSITORY: Julia is not installed in the build image.  Written to the letter of
# the reference's own tests (test/usecases/simple.jl, test/regridding.jl); a maintainer drops this
# file into ConservativeRegridding.jl as a package extension (weak deps: Libdl only for the core; CUDA,
# Healpix, RingGrids, Oceananigans for the optional sections at the end).  tests/test_abi.py parses the
# struct definitions and the ccall signatures below and checks them against the header / the ctypes
# mirror, so the layouts cannot drift silently.
#
# It plugs in at the two boundaries named in SURVEY.md section 8(b):
#   * build : replaces intersection_areas(...) + areas(...) (src/regridder/regridder.jl:125-163)
#   * apply : Regridder{W} is parametric in the matrix type W and perform_regridding! only needs
#             LinearAlgebra.mul!(y, R.intersections, x) (src/regridder/regrid.jl:95-98)
#
# The Julia API stays the same.  After `CRGB200.enable!()` the reference's own entry point
#     Regridder(dst, src; normalize)          (src/regridder/regridder.jl:105-123 -> :125-163)
# builds on the GPU (the methods below are more specific than the reference's `::Manifold` method and
# fall back to it with `invoke` when disabled, when a custom `intersection_operator` is given, or for
# ClimaCore spectral-element spaces, whose weights are not area weights); `Regridder(B200(), dst, src)`
# asks for it explicitly.  regrid!, transpose, R.dst_areas / R.src_areas, normalize! work unchanged.
module CRGB200

import ConservativeRegridding
import ConservativeRegridding: Regridder, Trees
import GeometryOps as GO, GeometryOpsCore as GOCore, GeoInterface as GI
import LinearAlgebra, SparseArrays

const lib = get(ENV, "CRGB200_LIB", "libcrgb200.so")

struct CrgOptions            # mirrors crg_options (48 bytes)
    manifold::Int32
    normalize::Int32
    radius::Float64
    area_threshold::Float64
    device::Int32
    build_transpose::Int32
    keep_candidates::Int32
    reserved::Int32
    stream::Ptr{Cvoid}
end

struct CrgCells              # mirrors crg_cells (32 bytes)
    verts::Ptr{Float64}
    offsets::Ptr{Int32}
    ncells::Int64
    nv::Int32
    reserved::Int32
end

struct CrgGrid               # mirrors crg_grid (120 bytes)
    kind::Int32
    flags::Int32
    cells::CrgCells
    n1::Int64
    n2::Int64
    p::NTuple{4, Float64}
    lat_deg::Ptr{Float64}
    cell_lo::Int64
    cell_hi::Int64
end

const GRID_CELLS, GRID_LONLAT, GRID_HEALPIX, GRID_FULL_RING, GRID_CUBED_SPHERE, GRID_REDUCED_RING = Int32.(0:5)
const NOCELLS = CrgCells(C_NULL, C_NULL, 0, 0, 0)

check(rc) = rc == 0 || error("libcrgb200: " * unsafe_string(ccall((:crg_last_error, lib), Cstring, ())))

"`R.intersections`: the device-resident matrix; `transpose` flips a flag over the same handle."
mutable struct B200Matrix <: AbstractMatrix{Float64}
    h::Ptr{Cvoid}
    n_dst::Int
    n_src::Int
    transposed::Bool
    owner::Union{Nothing, B200Matrix}   # transposed views keep the owning matrix (and its handle) alive
    nnz::Int                            # cached at construction (crg_dims), refreshed never: the pattern is immutable
    function B200Matrix(h, n_dst, n_src, transposed, owner, nnz)
        A = new(h, n_dst, n_src, transposed, owner, nnz)
        owner === nothing && finalizer(a -> ccall((:crg_free, lib), Cint, (Ptr{Cvoid},), a.h), A)
        return A
    end
end
function B200Matrix(h::Ptr{Cvoid})
    n_dst = Ref{Int64}(0); n_src = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(ccall((:crg_dims, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h, n_dst, n_src, nnz))
    return B200Matrix(h, n_dst[], n_src[], false, nothing, nnz[])
end
Base.size(A::B200Matrix) = A.transposed ? (A.n_src, A.n_dst) : (A.n_dst, A.n_src)
LinearAlgebra.transpose(A::B200Matrix) =
    B200Matrix(A.h, A.n_dst, A.n_src, !A.transposed, A.owner === nothing ? A : A.owner, A.nnz)
SparseArrays.nnz(A::B200Matrix) = A.nnz

function SparseArrays.sparse(A::B200Matrix)
    colptr = Vector{Int64}(undef, A.n_src + 1); rowval = Vector{Int64}(undef, A.nnz); nzval = Vector{Float64}(undef, A.nnz)
    check(ccall((:crg_export_csc, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                A.h, 1, colptr, rowval, nzval))
    S = SparseArrays.SparseMatrixCSC(A.n_dst, A.n_src, colptr, rowval, nzval)
    return A.transposed ? SparseArrays.sparse(transpose(S)) : S
end
SparseArrays.findnz(A::B200Matrix) = SparseArrays.findnz(SparseArrays.sparse(A))
Base.getindex(A::B200Matrix, i::Int, j::Int) = SparseArrays.sparse(A)[i, j]   # exports the matrix: tests only
Base.sum(A::B200Matrix; dims) = sum(SparseArrays.sparse(A); dims)
function Base.maximum(A::B200Matrix)
    m = Ref{Float64}(0.0)
    check(ccall((:crg_maximum, lib), Cint, (Ptr{Cvoid}, Ref{Float64}), A.h, m))
    return m[]
end

# Host or device memory alike: crg_apply detects device pointers (cudaPointerGetAttributes).
_ptr(x::Array{Float64}) = pointer(x)
_ptr(x::SubArray{Float64}) = pointer(x)
_apply!(A::B200Matrix, divide, y, x, K, ldy, ldx, level_fastest) =
    check(ccall((:crg_apply, lib), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int32),
                A.h, A.transposed, divide, _ptr(y), _ptr(x), K, ldy, ldx, level_fastest))

# y = A x without the area division (plain mul! semantics, used by generic code paths)
function LinearAlgebra.mul!(y::StridedVecOrMat{Float64}, A::B200Matrix, x::StridedVecOrMat{Float64})
    GC.@preserve y x _apply!(A, 0, y, x, size(x, 2), stride(y, 2), stride(x, 2), 0)
    return y
end

const B200Regridder = Regridder{B200Matrix}

# Fused path: one launch does mul! and `./= dst_areas` (regrid.jl:95-118).  The docs allow dispatching
# on perform_regridding!/finalize_regridding! "if absolutely necessary" (regrid.jl:93-94).
# R.dst_areas / R.src_areas are host vectors; the division uses the device copy, so edits go through
# `set_areas!` (or `normalize!`) below.
function ConservativeRegridding.perform_regridding!(dst::DenseVector{Float64}, r::B200Regridder,
                                                    src::DenseVector{Float64}; normalize = true, kwargs...)
    GC.@preserve dst src _apply!(r.intersections, normalize, dst, src, 1, length(dst), length(src), 0)
    return dst
end
ConservativeRegridding.finalize_regridding!(dst::DenseVector{Float64}, r::B200Regridder, dst_like::AbstractVector;
                                            kwargs...) = dst   # division already fused
# N-D StridedArray, any rank: ONE batched launch instead of the NDSliceLoop of K SpMVs (regrid.jl:225-318).
# dims = 1 (each level contiguous, the reference default) or dims = ndims (levels fastest) map straight onto the
# kernel's two layouts; anything else goes through the reference's slice loop.
function ConservativeRegridding.regrid!(dst::StridedArray{Float64, N}, r::B200Regridder, src::StridedArray{Float64, N};
                                        dims::Int = 1, normalize = true, kwargs...) where {N}
    N == 1 && return invoke(ConservativeRegridding.regrid!, Tuple{Any, Regridder, Any}, dst, r, src; normalize, kwargs...)
    1 <= dims <= N || throw(ArgumentError("dims=$dims is out of range for a $N-dimensional array"))
    other_d = ntuple(i -> i < dims ? size(dst, i) : size(dst, i + 1), N - 1)
    other_s = ntuple(i -> i < dims ? size(src, i) : size(src, i + 1), N - 1)
    other_d == other_s || throw(DimensionMismatch("source and destination non-spatial axes must match; got source axes $other_s and destination axes $other_d"))
    A = r.intersections
    n_out, n_in = size(A)
    (size(dst, dims) == n_out && size(src, dims) == n_in) || throw(DimensionMismatch("regridder is $(n_out)x$(n_in)"))
    K = prod(other_s)
    dense(a) = Base.iscontiguous(a)
    if dims == 1 && dense(dst) && dense(src)
        GC.@preserve dst src _apply!(A, normalize, dst, src, K, n_out, n_in, 0)
    elseif dims == N && dense(dst) && dense(src)
        GC.@preserve dst src _apply!(A, normalize, dst, src, K, K, K, 1)
    else
        return invoke(ConservativeRegridding.regrid!, Tuple{Any, Regridder, Any}, dst, r, src; dims, normalize, kwargs...)
    end
    return dst
end

"`LinearAlgebra.normalize!(R)` (regridder.jl:54-62): A, dst_areas, src_areas ./= maximum(A), on the device and in the host vectors."
function LinearAlgebra.normalize!(r::B200Regridder)
    check(ccall((:crg_normalize, lib), Cint, (Ptr{Cvoid},), r.intersections.h))
    refresh_areas!(r)
    return r
end
function refresh_areas!(r::B200Regridder)
    A = r.intersections
    d, s = A.transposed ? (r.src_areas, r.dst_areas) : (r.dst_areas, r.src_areas)
    check(ccall((:crg_areas, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), A.h, d, s))
    return r
end
"Replace the area vectors `regrid!` divides by (masking, custom normalisation): host vectors and device copy together."
function set_areas!(r::B200Regridder; dst_areas = nothing, src_areas = nothing)
    A = r.intersections
    dst_areas === nothing || copyto!(r.dst_areas, dst_areas)
    src_areas === nothing || copyto!(r.src_areas, src_areas)
    d, s = A.transposed ? (r.src_areas, r.dst_areas) : (r.dst_areas, r.src_areas)
    check(ccall((:crg_set_areas, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), A.h, d, s))
    return r
end

"Flatten `collect(Trees.getcell(tree))` into the (xyz | xy) vertex soup + offsets (open rings), preallocated."
function flatten_cells(manifold, tree)
    dim = manifold isa GO.Spherical ? 3 : 2
    cells = collect(Trees.getcell(tree))
    counts = map(c -> (n = GI.npoint(GI.getexterior(c)); n - 1), cells)      # reference rings are closed: drop the last point
    offs = Vector{Int32}(undef, length(cells) + 1); offs[1] = 0
    cumsum!(view(offs, 2:length(offs)), counts)
    verts = Vector{Float64}(undef, dim * offs[end])
    Threads.@threads for k in eachindex(cells)
        ring = GI.getexterior(cells[k]); o = dim * offs[k]
        for (j, p) in enumerate(GI.getpoint(ring))
            j > counts[k] && break
            if dim == 3
                verts[o + 3j - 2], verts[o + 3j - 1], verts[o + 3j] = p[1], p[2], p[3]
            else
                verts[o + 2j - 1], verts[o + 2j] = GI.x(p), GI.y(p)
            end
        end
    end
    return verts, offs
end

# ---- grid objects that a few numbers describe: no vertex soup, cells are generated on the device ------------------
"`describe(x)`: a `CrgGrid` descriptor (+ objects to keep alive) for grids the device can generate, else `nothing`."
describe(x) = nothing
"HealpixMap -> descriptor (ext/ConservativeRegriddingHealpixExt.jl:18,138-167): nside + ordering."
healpix_grid(nside::Integer; nested::Bool = false) =
    CrgGrid(GRID_HEALPIX, nested ? 1 : 0, NOCELLS, nside, 0, (0.0, 0.0, 0.0, 0.0), C_NULL, 0, 0)
"Oceananigans LatitudeLongitudeGrid -> descriptor (ext/ConservativeRegriddingOceananigansExt.jl:23-60,242-264)."
lonlat_grid(nlon, nlat; longitude = (0.0, 360.0), latitude = (-90.0, 90.0)) =
    CrgGrid(GRID_LONLAT, 0, NOCELLS, nlon, nlat, (longitude[1], longitude[2], latitude[1], latitude[2]), C_NULL, 0, 0)
"RingGrids AbstractFullGrid -> descriptor (ext/ConservativeRegriddingRingGridsExt.jl:22-50); `latd` north -> south."
full_ring_grid(nlon, latd::Vector{Float64}; lon_first = 0.0) =
    CrgGrid(GRID_FULL_RING, 0, NOCELLS, nlon, length(latd), (lon_first, 0.0, 0.0, 0.0), pointer(latd), 0, 0)
"Octahedral Gaussian grid O<n> -> descriptor (no cells in the reference, RingGridsExt.jl:18-20): ring of rank j has 16 + 4j points."
octahedral_grid(latd::Vector{Float64}; lon_first = 0.0) =
    CrgGrid(GRID_REDUCED_RING, 0, NOCELLS, 0, length(latd), (lon_first, 16.0, 4.0, 0.0), pointer(latd), 0, 0)

struct B200 end          # backend tag: Regridder(B200(), dst, src; ...)
const ENABLED = Ref(false)
"Route `Regridder(dst, src; ...)` through the GPU engine (`enable!(false)` restores the reference's CPU path)."
enable!(on::Bool = true) = (ENABLED[] = on)

function _wrap(h::Ptr{Cvoid})
    A = B200Matrix(h)
    dst_areas = Vector{Float64}(undef, A.n_dst); src_areas = Vector{Float64}(undef, A.n_src)
    check(ccall((:crg_areas, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, dst_areas, src_areas))
    return Regridder(A, dst_areas, src_areas, zeros(A.n_dst), zeros(A.n_src))     # regridder.jl:155-156
end

"""
    Regridder(B200(), [manifold,] dst, src; normalize = false, device = -1)

The reference constructor (regridder.jl:105-163) on the GPU: same struct, `intersections::B200Matrix`,
`dst_areas` / `src_areas` host `Vector{Float64}` (geometric areas, regridder.jl:165-178), `transpose(R)`
shares every array (`===`, test/usecases/simple.jl:58-64).  Grids a few numbers describe (HealpixMap,
LatitudeLongitudeGrid, RingGrids full / octahedral grids: `describe`) are generated on the device; every
other grid goes through `Trees.treeify` + `Trees.getcell` like the reference.
"""
function Regridder(::B200, manifold::GOCore.Manifold, dst, src; normalize = false, device = -1, kwargs...)
    sph = manifold isa GO.Spherical
    opts = CrgOptions(sph ? 1 : 0, normalize, sph ? manifold.radius : 1.0, 0.0, device, 1, 0, 0, C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dd, ds = sph ? (describe(dst), describe(src)) : (nothing, nothing)
    keep = Any[]
    function as_grid(x, d)
        d === nothing || return d isa Tuple ? (push!(keep, d[2]); d[1]) : d
        verts, offs = flatten_cells(manifold, Trees.treeify(manifold, x))
        push!(keep, verts, offs)
        return CrgGrid(GRID_CELLS, 0, CrgCells(pointer(verts), pointer(offs), length(offs) - 1, 0, 0), 0, 0,
                       (0.0, 0.0, 0.0, 0.0), C_NULL, 0, 0)
    end
    gd, gs = as_grid(dst, dd), as_grid(src, ds)
    GC.@preserve keep begin
        if gd.kind == GRID_CELLS && gs.kind == GRID_CELLS
            check(ccall((:crg_build, lib), Cint, (Ref{CrgOptions}, Ref{CrgCells}, Ref{CrgCells}, Ptr{Ptr{Cvoid}}),
                        opts, gd.cells, gs.cells, h))
        else
            check(ccall((:crg_build_grids, lib), Cint, (Ref{CrgOptions}, Ref{CrgGrid}, Ref{CrgGrid}, Ptr{Ptr{Cvoid}}),
                        opts, gd, gs, h))
        end
    end
    return _wrap(h[])
end
function Regridder(b::B200, dst, src; kwargs...)
    md, ms = GOCore.best_manifold(dst), GOCore.best_manifold(src)          # regridder.jl:105-123
    m = (md isa GO.Spherical || ms isa GO.Spherical) ? (md isa GO.Spherical ? md : ms) : md
    return Regridder(b, m, dst, src; kwargs...)
end

# `Regridder(manifold, dst, src; ...)` itself, when enabled: more specific than the reference's `::Manifold` method.
for M in (:(GO.Spherical), :(GO.Planar))
    @eval function Regridder(manifold::$M, dst, src; normalize = false, intersection_operator = nothing, kwargs...)
        if ENABLED[] && intersection_operator === nothing
            return Regridder(B200(), manifold, dst, src; normalize, kwargs...)
        end
        kw = intersection_operator === nothing ? (; normalize, kwargs...) : (; normalize, intersection_operator, kwargs...)
        return invoke(Regridder, Tuple{GOCore.Manifold, Any, Any}, manifold, dst, src; kw...)
    end
end

# ---- optional sections: loaded when the corresponding package is (package extensions in a real checkout) -----------
# CUDA.jl: device-resident fields go to the kernels as they are (no host round trip); Oceananigans GPU fields and
# Healpix / RingGrids fields whose parent array is a CuArray land here through the reference's own
# extract_*_arraylike plumbing (ext/ConservativeRegriddingHealpixExt.jl:186-202, ...OceananigansExt.jl:199-214).
function __init_cuda__(CUDA)
    @eval begin
        _ptr(x::$CUDA.CuArray{Float64}) = reinterpret(Ptr{Float64}, pointer(x))
        function ConservativeRegridding.perform_regridding!(dst::$CUDA.CuVector{Float64}, r::B200Regridder,
                                                            src::$CUDA.CuVector{Float64}; normalize = true, kwargs...)
            $CUDA.synchronize()      # the library runs on its own stream: order it after the producers of `src`
            GC.@preserve dst src _apply!(r.intersections, normalize, dst, src, 1, length(dst), length(src), 0)
            return dst
        end
        ConservativeRegridding.finalize_regridding!(dst::$CUDA.CuVector{Float64}, r::B200Regridder, dst_like::AbstractVector; kwargs...) = dst
        # regridder temporaries on the device for package fields (regridder.jl:152 "TODO: make this GPU-compatible?")
        on_device(r::B200Regridder) = Regridder(r.intersections, r.dst_areas, r.src_areas, $CUDA.zeros(Float64, length(r.dst_temp)), $CUDA.zeros(Float64, length(r.src_temp)))
    end
end
function __init_healpix__(Healpix)
    @eval describe(m::$Healpix.HealpixMap{T, O}) where {T, O} = healpix_grid(m.resolution.nside; nested = O <: $Healpix.NestedOrder)
end
function __init_ringgrids__(RingGrids)
    @eval begin
        function describe(g::$RingGrids.AbstractFullGrid)
            latd = collect(Float64, $RingGrids.get_latd(g)); lond = $RingGrids.get_lond(g)
            return (full_ring_grid(length(lond), latd; lon_first = Float64(lond[1])), latd)
        end
        function describe(g::$RingGrids.OctahedralGaussianGrid)
            latd = collect(Float64, $RingGrids.get_latd(g))
            return (octahedral_grid(latd), latd)
        end
    end
end
function __init_oceananigans__(Oceananigans)
    @eval function describe(g::$Oceananigans.LatitudeLongitudeGrid)
        # regularly spaced faces only (ξnode/ηnode at Face, Face: OceananigansExt.jl:47-60); stretched grids go through getcell
        λ = $Oceananigans.Grids.λnodes(g, $Oceananigans.Face()); φ = $Oceananigans.Grids.φnodes(g, $Oceananigans.Face())
        (allequal(round.(diff(λ); digits = 10)) && allequal(round.(diff(φ); digits = 10))) || return nothing
        Nx, Ny = size(g, 1), size(g, 2)
        return lonlat_grid(Nx, Ny; longitude = (Float64(λ[1]), Float64(λ[1]) + Nx * Float64(λ[2] - λ[1])),
                           latitude = (Float64(φ[1]), Float64(φ[end])))
    end
    # tripolar fold rows: the ghost cells the extension's PaddedTreeWrapper adds (OceananigansExt.jl:66-187) arrive here
    # as degenerate rings from getcell; the engine keeps them out of the candidates, and finalize_regridding! of the
    # extension mirrors the fold partners as before (crg_mirror_fold_partners does the same for device fields).
end
function __init__()
    hooks = Dict("CUDA" => __init_cuda__, "Healpix" => __init_healpix__, "RingGrids" => __init_ringgrids__,
                 "Oceananigans" => __init_oceananigans__)
    for (key, mod) in Base.loaded_modules
        haskey(hooks, key.name) && hooks[key.name](mod)
    end
end

end # module

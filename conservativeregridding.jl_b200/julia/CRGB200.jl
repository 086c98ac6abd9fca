# CRGB200.jl -- the reference-side binding of libcrgb200.so (include/crg_b200.h).
#
# UNTESTED IN THIS REPOSITORY: Julia is not installed in the build image.  Written to the letter of
# the reference's own tests (test/usecases/simple.jl, test/regridding.jl); a maintainer drops this
# file into ConservativeRegridding.jl as a package extension (weak dep on nothing but Libdl).
#
# It plugs in at the two boundaries named in SURVEY.md section 8(b):
#   * build : replaces intersection_areas(...) + areas(...) (src/regridder/regridder.jl:125-163)
#   * apply : Regridder{W} is parametric in the matrix type W and perform_regridding! only needs
#             LinearAlgebra.mul!(y, R.intersections, x) (src/regridder/regrid.jl:95-98)
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

check(rc) = rc == 0 || error("libcrgb200: " * unsafe_string(ccall((:crg_last_error, lib), Cstring, ())))

"`R.intersections`: the device-resident matrix; `transpose` flips a flag over the same handle."
mutable struct B200Matrix <: AbstractMatrix{Float64}
    h::Ptr{Cvoid}
    n_dst::Int
    n_src::Int
    transposed::Bool
    owner::Union{Nothing, B200Matrix}   # transposed views keep the owning matrix (and its handle) alive
    function B200Matrix(h, n_dst, n_src, transposed, owner)
        A = new(h, n_dst, n_src, transposed, owner)
        owner === nothing && finalizer(a -> ccall((:crg_free, lib), Cint, (Ptr{Cvoid},), a.h), A)
        return A
    end
end
Base.size(A::B200Matrix) = A.transposed ? (A.n_src, A.n_dst) : (A.n_dst, A.n_src)
LinearAlgebra.transpose(A::B200Matrix) =
    B200Matrix(A.h, A.n_dst, A.n_src, !A.transposed, A.owner === nothing ? A : A.owner)
Base.getindex(A::B200Matrix, i::Int, j::Int) = SparseArrays.sparse(A)[i, j]   # slow; tests only

function SparseArrays.sparse(A::B200Matrix)
    nnz = Ref{Int64}(0)
    check(ccall((:crg_dims, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), A.h, C_NULL, C_NULL, nnz))
    colptr = Vector{Int64}(undef, A.n_src + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{Float64}(undef, nnz[])
    check(ccall((:crg_export_csc, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                A.h, 1, colptr, rowval, nzval))
    S = SparseArrays.SparseMatrixCSC(A.n_dst, A.n_src, colptr, rowval, nzval)
    return A.transposed ? SparseArrays.sparse(transpose(S)) : S
end
SparseArrays.findnz(A::B200Matrix) = SparseArrays.findnz(SparseArrays.sparse(A))
SparseArrays.nnz(A::B200Matrix) = SparseArrays.nnz(SparseArrays.sparse(A))
function Base.maximum(A::B200Matrix)
    m = Ref{Float64}(0.0)
    check(ccall((:crg_maximum, lib), Cint, (Ptr{Cvoid}, Ref{Float64}), A.h, m))
    return m[]
end

# y = A x without the area division (plain mul! semantics, used by generic code paths)
function LinearAlgebra.mul!(y::StridedVecOrMat{Float64}, A::B200Matrix, x::StridedVecOrMat{Float64})
    K = size(x, 2)
    check(ccall((:crg_apply, lib), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int32),
                A.h, A.transposed, 0, y, x, K, stride(y, 2), stride(x, 2), 0))
    return y
end

const B200Regridder = Regridder{B200Matrix}

# Fused path: one launch does mul! and `./= dst_areas` (regrid.jl:95-118).  The docs allow dispatching
# on perform_regridding!/finalize_regridding! "if absolutely necessary" (regrid.jl:93-94).
function ConservativeRegridding.perform_regridding!(dst::DenseVector{Float64}, r::B200Regridder,
                                                    src::DenseVector{Float64}; normalize = true, kwargs...)
    A = r.intersections
    check(ccall((:crg_apply, lib), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int32),
                A.h, A.transposed, normalize, dst, src, 1, length(dst), length(src), 0))
    return dst
end
ConservativeRegridding.finalize_regridding!(dst::DenseVector{Float64}, r::B200Regridder, dst_like::AbstractVector;
                                            kwargs...) = dst   # division already fused
# N-D StridedArray with dims = 1 (each level contiguous): ONE batched launch instead of the
# NDSliceLoop of K SpMVs (regrid.jl:303-318).
function ConservativeRegridding.regrid!(dst::StridedMatrix{Float64}, r::B200Regridder, src::StridedMatrix{Float64};
                                        dims::Int = 1, normalize = true, kwargs...)
    dims in (1, 2) || throw(ArgumentError("dims=$dims is out of range for a 2-dimensional array"))
    other = dims == 1 ? 2 : 1
    size(dst, other) == size(src, other) ||
        throw(DimensionMismatch("source and destination non-spatial axes must match"))
    A = r.intersections
    lf = dims == 2       # (K, ncells) column-major = level-fastest
    check(ccall((:crg_apply, lib), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int32),
                A.h, A.transposed, normalize, dst, src, size(src, other), stride(dst, 2), stride(src, 2), lf))
    return dst
end

"Flatten `collect(Trees.getcell(tree))` into the (xyz | xy) vertex soup + offsets (open rings)."
function flatten_cells(manifold, tree)
    dim = manifold isa GO.Spherical ? 3 : 2
    verts = Float64[]; offs = Int32[0]
    for cell in Trees.getcell(tree)
        pts = collect(GI.getpoint(GI.getexterior(cell)))
        pts[1] == pts[end] && pop!(pts)                       # drop the closing vertex
        for p in pts
            dim == 3 ? append!(verts, (p[1], p[2], p[3])) : append!(verts, (GI.x(p), GI.y(p)))
        end
        push!(offs, offs[end] + length(pts))
    end
    return verts, offs
end

"""
    b200_regridder(manifold, dst, src; normalize = false, device = -1)

Drop-in for `Regridder(manifold, dst, src; normalize)` (regridder.jl:125-163): same struct, with
`intersections::B200Matrix`; `dst_areas`, `src_areas` are host `Vector{Float64}` (geometric areas);
`transpose(R)` shares every array (`===`, test/usecases/simple.jl:58-64).
"""
function b200_regridder(manifold::GOCore.Manifold, dst, src; normalize = false, device = -1)
    dst_tree = Trees.treeify(manifold, dst); src_tree = Trees.treeify(manifold, src)
    dv, doff = flatten_cells(manifold, dst_tree); sv, soff = flatten_cells(manifold, src_tree)
    sph = manifold isa GO.Spherical
    opts = CrgOptions(sph ? 1 : 0, normalize, sph ? manifold.radius : 1.0, 0.0, device, 1, 0, 0, C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve dv doff sv soff begin
        cd = CrgCells(pointer(dv), pointer(doff), length(doff) - 1, 0, 0)
        cs = CrgCells(pointer(sv), pointer(soff), length(soff) - 1, 0, 0)
        check(ccall((:crg_build, lib), Cint, (Ref{CrgOptions}, Ref{CrgCells}, Ref{CrgCells}, Ptr{Ptr{Cvoid}}),
                    opts, cd, cs, h))
    end
    n_dst, n_src = length(doff) - 1, length(soff) - 1
    dst_areas = Vector{Float64}(undef, n_dst); src_areas = Vector{Float64}(undef, n_src)
    check(ccall((:crg_areas, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], dst_areas, src_areas))
    A = B200Matrix(h[], n_dst, n_src, false, nothing)
    return Regridder(A, dst_areas, src_areas, zeros(n_dst), zeros(n_src))
end

# ---- described grids: no vertex soup, cells are generated on the device (crg_build_grids) --------------
struct CrgGrid               # mirrors crg_grid
    kind::Int32
    flags::Int32
    cells::CrgCells
    n1::Int64
    n2::Int64
    p::NTuple{4, Float64}
    lat_deg::Ptr{Float64}
end
const NOCELLS = CrgCells(C_NULL, C_NULL, 0, 0, 0)
"HealpixMap -> descriptor (ext/ConservativeRegriddingHealpixExt.jl:18): nside + ordering."
healpix_grid(nside::Integer; nested::Bool = false) = CrgGrid(2, nested ? 1 : 0, NOCELLS, nside, 0, (0.0, 0.0, 0.0, 0.0), C_NULL)
"Oceananigans LatitudeLongitudeGrid -> descriptor (ext/ConservativeRegriddingOceananigansExt.jl:242-264)."
lonlat_grid(nlon, nlat; longitude = (0.0, 360.0), latitude = (-90.0, 90.0)) =
    CrgGrid(1, 0, NOCELLS, nlon, nlat, (longitude[1], longitude[2], latitude[1], latitude[2]), C_NULL)

function b200_regridder(dst::CrgGrid, src::CrgGrid; radius = 1.0, normalize = false, device = -1)
    opts = CrgOptions(1, normalize, radius, 0.0, device, 1, 0, 0, C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:crg_build_grids, lib), Cint, (Ref{CrgOptions}, Ref{CrgGrid}, Ref{CrgGrid}, Ptr{Ptr{Cvoid}}), opts, dst, src, h))
    n_dst = Ref{Int64}(0); n_src = Ref{Int64}(0)
    check(ccall((:crg_dims, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h[], n_dst, n_src, C_NULL))
    dst_areas = Vector{Float64}(undef, n_dst[]); src_areas = Vector{Float64}(undef, n_src[])
    check(ccall((:crg_areas, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], dst_areas, src_areas))
    return Regridder(B200Matrix(h[], n_dst[], n_src[], false, nothing), dst_areas, src_areas, zeros(n_dst[]), zeros(n_src[]))
end

end # module

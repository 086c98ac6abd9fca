/*
 * crg_b200.h -- C ABI of libcrgb200.so, the B200-native (sm_100a) engine behind
 * ConservativeRegridding.jl's two hot paths: building a Regridder and applying it.
 *
 * The reference (JuliaGeo/ConservativeRegridding.jl v0.2.5, pure Julia) has no FFI; the
 * entry points below are what a Julia `ccall` binding for this path binds, one per
 * reference interface they replace (citations relative to the reference tree):
 *
 *   crg_build          <- intersection_areas(manifold, threaded, dst_tree, src_tree)
 *                         src/regridder/intersection_areas.jl:67-122  +  areas(manifold, x, tree)
 *                         src/regridder/regridder.jl:149-150,165-178   +  normalize!  :54-62,160
 *                         (input = collect(Trees.getcell(tree)) in field-linear order,
 *                          src/trees/interfaces.jl:231-243)
 *   crg_build_from_coo <- SparseArrays.sparse(i2s, i1s, areas, n_dst, n_src)
 *                         src/regridder/intersection_areas.jl:115-121 (custom
 *                         `intersection_operator` results, regridder.jl:128)
 *   crg_apply          <- perform_regridding! (LinearAlgebra.mul!) + finalize_regridding!
 *                         src/regridder/regrid.jl:95-118, and the NDSliceLoop slice loop
 *                         :303-318 (K right-hand sides in one launch); `transpose` = the
 *                         transpose(R) regridder of regridder.jl:49-50
 *   crg_normalize      <- LinearAlgebra.normalize!(::Regridder)  regridder.jl:54-62
 *   crg_areas          <- R.dst_areas / R.src_areas              regridder.jl:29-31
 *   crg_export_csc     <- R.intersections :: SparseMatrixCSC{Float64,Int64}  regridder.jl:9-10
 *                         (findnz / ESMF export, ext/ConservativeRegriddingNCDatasetsExt.jl:25-32)
 *   crg_dims           <- Base.size(::Regridder)                 regridder.jl:52
 *
 * Conventions: every function returns CRG_OK (0) or a negative error code and never throws
 * or aborts; the message of the last error on the calling thread is crg_last_error().
 * Pointers may be host or device pointers (detected with cudaPointerGetAttributes) unless
 * stated otherwise.  The caller owns every buffer it passes; the library owns the device
 * memory of a crg_regridder until crg_free.  Calls are synchronous (the handle's stream is
 * synchronised before returning) except crg_apply_async.  A handle is not re-entrant (the
 * reference Regridder is not either: regrid! mutates src_temp/dst_temp, regrid.jl:85-88).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * CRG_ERR_NO_DEVICE.
 */
#ifndef CRG_B200_H
#define CRG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRG_OK 0
#define CRG_ERR_INVALID (-1)     /* bad argument */
#define CRG_ERR_CUDA (-2)        /* CUDA runtime error, see crg_last_error() */
#define CRG_ERR_NOMEM (-3)       /* allocation failed / size limit exceeded */
#define CRG_ERR_UNSUPPORTED (-4) /* e.g. polygon with more than CRG_MAX_VERTS vertices */
#define CRG_ERR_NO_DEVICE (-5)   /* no usable CUDA device */

#define CRG_PLANAR 0    /* GeometryOps.Planar():    vertices are (x, y) pairs      */
#define CRG_SPHERICAL 1 /* GeometryOps.Spherical(): vertices are unit (x, y, z)    */

#define CRG_MAX_VERTS 8 /* max vertices of one input cell (open ring) */

typedef struct crg_regridder crg_regridder;

/* Keyword arguments of Regridder(dst, src; ...) (regridder.jl:125-131) + engine knobs. */
typedef struct crg_options {
    int32_t manifold;        /* CRG_PLANAR | CRG_SPHERICAL                                   */
    int32_t normalize;       /* kw `normalize` (default false): divide A and both area        */
                             /* vectors by maximum(A)                                        */
    double radius;           /* Spherical(; radius): areas are unit-sphere areas * radius^2  */
    double area_threshold;   /* keep pairs with area > threshold * radius^2; 0 = reference    */
                             /* semantics (`area > 0`, intersection_areas.jl:24)              */
    int32_t device;          /* CUDA device ordinal; -1 = current device                      */
    int32_t build_transpose; /* also assemble CSR(A^T) (= the reference's CSC) for            */
                             /* transpose(R) and crg_export_csc; default 1                    */
    int32_t keep_candidates; /* keep the broad-phase pair list for crg_candidates (tests)     */
    int32_t reserved;
    void *stream;            /* cudaStream_t to run on (e.g. the caller's current stream);     */
                             /* NULL = a stream owned by the handle                           */
} crg_options;

/* A grid as the flat list of its cells in field-linear order (= collect(getcell(tree))).
 * Open rings (no repeated closing vertex), any orientation, convex.                         */
typedef struct crg_cells {
    const double *verts;    /* fixed: [ncells][nv][dim]; ragged: [offsets[ncells]][dim]        */
    const int32_t *offsets; /* NULL => every cell has `nv` vertices; else ncells+1 offsets     */
    int64_t ncells;
    int32_t nv;             /* vertices per cell when offsets == NULL (3..CRG_MAX_VERTS)       */
    int32_t reserved;
} crg_cells;

/* A grid given either explicitly (cells) or by a few numbers, in which case the cell vertices are
 * generated on the device, straight into the build's scratch memory (no host vertex soup, no
 * upload).  Cell conventions and field-linear order are the reference's:
 *   CRG_GRID_LONLAT       Oceananigans LatitudeLongitudeGrid   ext/ConservativeRegriddingOceananigansExt.jl:23-60,242-264
 *                         n1 = nlon, n2 = nlat, p = {lon0, lon1, lat0, lat1} degrees; lon fastest, S -> N
 *   CRG_GRID_HEALPIX      HealpixMap                           ext/ConservativeRegriddingHealpixExt.jl:76-90,138-167
 *                         n1 = nside (power of two), flags = 0 ring order | 1 nested order
 *   CRG_GRID_FULL_RING    RingGrids AbstractFullGrid           ext/ConservativeRegriddingRingGridsExt.jl:22-50
 *                         n1 = nlon, n2 = nlat, p[0] = longitude of the first point, lat_deg = the nlat
 *                         ring latitudes in degrees, north -> south (host or device pointer)
 *   CRG_GRID_CUBED_SPHERE equiangular gnomonic cubed sphere, n1 = cells per panel edge; 6 panels,
 *                         panel-major (each panel is a CellBasedGrid, src/trees/grids.jl:57-85)
 *   CRG_GRID_REDUCED_RING RingGrids reduced grid (octahedral Gaussian O<n>: SpeedyWeather's default); the reference
 *                         has NO cells for it (ext/ConservativeRegriddingRingGridsExt.jl:18-20 errors): the full-grid
 *                         rule generalised -- band between pole-pinned mid-latitudes x longitude interval centred on
 *                         the point.  n2 = number of rings (even), lat_deg = ring latitudes north -> south, the ring of
 *                         rank j = 1, 2, .. from either pole has p[1] + p[2] * j points (octahedral: 16 + 4 j), p[0] =
 *                         longitude of the first point of every ring; ring-major north -> south, longitude fastest */
#define CRG_GRID_CELLS 0
#define CRG_GRID_LONLAT 1
#define CRG_GRID_HEALPIX 2
#define CRG_GRID_FULL_RING 3
#define CRG_GRID_CUBED_SPHERE 4
#define CRG_GRID_REDUCED_RING 5

typedef struct crg_grid {
    int32_t kind;
    int32_t flags;
    crg_cells cells;       /* kind == CRG_GRID_CELLS */
    int64_t n1, n2;
    double p[4];
    const double *lat_deg;
    int64_t cell_lo, cell_hi; /* described grids: only the cells [cell_lo, cell_hi) of the field-linear order (a        */
                              /* destination block or a source halo of a sharded build); 0, 0 = the whole grid       */
} crg_grid;

/* Counters and per-phase device times (CUDA events, milliseconds) of the last build. */
typedef struct crg_build_stats {
    int64_t n_dst, n_src;
    int64_t n_candidates;   /* pairs written by the broad phase and clipped                  */
    int64_t nnz;            /* pairs with area > threshold (after duplicate summation)        */
    int64_t n_bins, n_bin_entries, n_big_dst, n_big_src;
    double ms_total;        /* whole crg_build call on the host clock (incl. H2D, allocation) */
    double ms_h2d;          /* host clock spent staging host inputs on the device             */
    double ms_device;       /* first kernel to last kernel, CUDA events                       */
    double ms_bounds, ms_bin, ms_query, ms_clip, ms_sort_csr, ms_sort_csc, ms_areas, ms_finish;
    double bin_size;        /* broad-phase bin size (radians on the sphere)                   */
    int32_t sort_passes_csr, sort_passes_csc;
} crg_build_stats;

int crg_options_init(crg_options *opts);

int crg_build(const crg_options *opts, const crg_cells *dst, const crg_cells *src,
              crg_regridder **out);

/* Same as crg_build with grids that may be described instead of listed (spherical manifold). */
int crg_build_grids(const crg_options *opts, const crg_grid *dst, const crg_grid *src, crg_regridder **out);

/* Number of cells of a grid; and its cell vertices ([ncells][4][3] doubles, host or device `verts`)
 * generated on `device` -- what crg_build_grids feeds to the build (tests / export).             */
int crg_grid_ncells(const crg_grid *g, int64_t *ncells);
int crg_grid_cells(const crg_grid *g, int32_t device, double *verts);

/* areas(manifold, x, tree) = [GO.area(manifold, cell) for cell in getcell(tree)] for ONE grid
 * (src/regridder/regridder.jl:165-178): geometric cell areas * radius^2, ncells doubles (host or device).
 * (A sharded regridder computes the areas of the replicated side in equal shares, one per rank.)               */
int crg_grid_areas(const crg_options *opts, const crg_grid *g, double *areas);

/* Assemble from explicit (dst_idx, src_idx, area) triples (0-based; duplicates are summed;
 * non-positive areas must already be dropped by the caller, intersection_areas.jl:24).
 * dst_areas/src_areas are the geometric cell areas (copied).                                */
int crg_build_from_coo(const crg_options *opts, int64_t n_dst, int64_t n_src, int64_t nnz,
                       const int64_t *dst_idx, const int64_t *src_idx, const double *area,
                       const double *dst_areas, const double *src_areas, crg_regridder **out);

/* compute_intersection_areas (src/regridder/intersection_areas.jl:4-32) with the default operator
 * (regridder.jl:87-103) for an explicit list of 0-based (src, dst) cell pairs, run by the kernels of
 * crg_build: area_out[k] = area(src[src_idx[k]] n dst[dst_idx[k]]) * radius^2, or 0 when the pair does
 * not survive `area > threshold`.  Index and output arrays may be host or device pointers.  (The
 * parity harness checks these per-pair values against 50-digit arithmetic.)                          */
int crg_clip_pairs(const crg_options *opts, const crg_cells *dst, const crg_cells *src, int64_t n_pairs,
                   const int64_t *src_idx, const int64_t *dst_idx, double *area_out);

/* Releases the handle's device memory, stream-ordered on the stream the handle works on (crg_options.stream /
 * crg_set_stream, else the library's): a caller-provided stream must still exist when crg_free is called. */
int crg_free(crg_regridder *r);

int crg_dims(const crg_regridder *r, int64_t *n_dst, int64_t *n_src, int64_t *nnz);
int crg_stats(const crg_regridder *r, crg_build_stats *stats);

/* Geometric cell areas (already divided by maximum(A) when normalised). Either may be NULL. */
int crg_areas(const crg_regridder *r, double *dst_areas, double *src_areas);

/* Overwrite the area vectors the fused division of crg_apply uses (regrid! divides by
 * `regridder.dst_areas`, regrid.jl:104-118, which a user may edit in place -- masking, custom
 * normalisation).  Either may be NULL (left as is); host or device pointers.                     */
int crg_set_areas(crg_regridder *r, const double *dst_areas, const double *src_areas);

/* R.intersections as CSC (n_dst x n_src): colptr[n_src+1], rowval[nnz], nzval[nnz], rows
 * sorted within each column.  index_base = 1 for Julia.  Host pointers only.  Any NULL is
 * skipped.  crg_export_csr: the same matrix by rows (rowptr[n_dst+1], colval, nzval).        */
int crg_export_csc(const crg_regridder *r, int32_t index_base, int64_t *colptr, int64_t *rowval,
                   double *nzval);
int crg_export_csr(const crg_regridder *r, int32_t index_base, int64_t *rowptr, int64_t *colval,
                   double *nzval);

/* Broad-phase pair list (only if opts.keep_candidates): n_candidates (src, dst) pairs.       */
int crg_candidates(const crg_regridder *r, int64_t *src_idx, int64_t *dst_idx);

int crg_normalize(crg_regridder *r);

/* The two halves of normalize! for destination-sharded regridders (regridder.jl:54-62 applied to a
 * row block): crg_maximum = maximum(A) of this handle's block (0 for an empty block); crg_scale
 * divides A, A^T and both area vectors by `divisor` (> 0) -- the maximum over all blocks.        */
int crg_maximum(crg_regridder *r, double *out);
int crg_scale(crg_regridder *r, double divisor);

/* dst = A * src (transpose = 0) or A^T * src (transpose = 1), then divided element-wise by
 * the output grid's areas when divide_by_area != 0 (the `normalize` kw of regrid!,
 * regrid.jl:104-118), for K right-hand sides at once.
 *   level_fastest = 0: field k occupies src[k*ld_src + cell], dst[k*ld_dst + cell]
 *                      (Julia `dims = 1` of an (ncells, K) array; ld >= ncells)
 *   level_fastest = 1: src[cell*ld_src + k], dst[cell*ld_dst + k]  (`dims = 2` of (K, ncells))
 * src and dst must not alias.                                                               */
int crg_apply(crg_regridder *r, int32_t transpose, int32_t divide_by_area, double *dst,
              const double *src, int64_t K, int64_t ld_dst, int64_t ld_src, int32_t level_fastest);

/* Same, device pointers only, enqueued on the handle's stream without synchronising. */
int crg_apply_async(crg_regridder *r, int32_t transpose, int32_t divide_by_area, double *dst,
                    const double *src, int64_t K, int64_t ld_dst, int64_t ld_src,
                    int32_t level_fastest);

/* mirror_fold_partners! (ext/ConservativeRegriddingOceananigansExt.jl:216-240): on a tripolar grid with a
 * RightCenterFolded north row (nx x ny cells, field index i + j * nx) every physical cell of the last row shows
 * up at two field slots; one carries the polygon, its partner is a zero-area ghost whose matrix row is empty.
 * After regrid! the value of every primary slot r (0-based: 0 .. nx/4-1 and nx/2 .. nx/2+nx/4-1) is copied into
 * its partner nx-1-r.  `field`: K fields laid out like crg_apply's dst (ld, level_fastest); device pointer: a
 * kernel on `stream` (NULL = the library stream of `device`, synchronised before returning); host pointer: copied
 * in place on the host.                                                                                       */
int crg_mirror_fold_partners(double *field, int64_t nx, int64_t ny, int64_t K, int64_t ld, int32_t level_fastest,
                             int32_t device, void *stream);

/* Run the handle's work on a caller-owned cudaStream_t (e.g. torch's current stream).
 * NULL restores the handle's own (non-blocking) stream; the legacy default stream is named
 * explicitly as cudaStreamLegacy ((cudaStream_t)0x1), the per-thread one as cudaStreamPerThread
 * ((cudaStream_t)0x2).  Does not wait on the host: work already enqueued on the previous stream is
 * ordered before what follows on the new one with an event.  The library orders its own work on the
 * handle's stream only: producers of `src` and consumers of `dst` on OTHER streams are the caller's
 * to order (or pass their stream here, which is what the Python/torch front end does).            */
int crg_set_stream(crg_regridder *r, void *cuda_stream);
int crg_synchronize(crg_regridder *r);

/* Algorithmic bytes moved by one crg_apply call:
 * 12*nnz + 4*(n_out+1) + 8*n_out*[divide] + 8*K*(n_in + n_out).                             */
int crg_apply_bytes(const crg_regridder *r, int32_t transpose, int32_t divide_by_area, int64_t K,
                    int64_t *bytes);

/* Number of CUDA kernels this library has launched in the calling process so far. */
int crg_launch_count(uint64_t *count);

const char *crg_last_error(void);
int crg_device_count(int32_t *count);
const char *crg_version(void);

/* FP64 FMA micro-benchmark on `device` (roofline denominator of the clip kernel): sustained
 * DFMA throughput in TFLOP/s (2 flops per FMA).                                              */
int crg_fp64_peak(int32_t device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* CRG_B200_H */

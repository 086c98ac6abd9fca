// broadphase.cuh -- device-side candidate search (replaces the dual-tree DFS of
// /root/reference/src/utils/MultithreadedDualDepthFirstSearch.jl:12-65 and the bounding-cap
// machinery of src/trees/grids.jl:245-287).
//
// Spherical cells are binned on the six faces of a cube in *equiangular gnomonic*
// coordinates (alpha, beta) = (f(b/a), f(c/a)), f(u) = u/sqrt(1+u^2): the gnomonic map sends great-circle
// arcs to straight segments, so the exact bounding box of a great-circle polygon on a face
// is the bounding box of its projected vertices -- no arc-bulge terms, no poles, no date
// line.  Each face's bin domain is extended by a margin >= the largest destination cell, so
// a destination cell queries ONLY its home face (the face of its vertex mean) and every
// (src, dst) pair is produced on exactly one face; inside the face the pair is reported only
// in the first bin common to both boxes.  Planar cells use one "face" with (alpha, beta) =
// (x, y).  Cells too large for this scheme (angular diameter >= 0.2 rad, or covering more
// than BP_MAX_COVER bins, or not inside their home face's domain) are paired by brute force
// -- they only occur in toy grids.
//
// Pipeline: bounds -> [host picks the bin size] -> count -> scan -> fill   (source cells)
//           query-count -> scan -> query-fill                              (destination cells)
// Entries carry the source cell's box quantised to 1/16 bin, so the query rejects
// non-overlapping boxes without touching the source vertices.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "geom.cuh"

namespace crg {

constexpr int BP_SUB = 16;              // quantisation: sub-bins per bin
constexpr int BP_MAX_BINS_1D = 4095;    // so that quantised coordinates fit 16 bits
constexpr int BP_MAX_COVER = 4096;      // a source cell covering more bins is "big"
constexpr int BP_MAX_QUERY = 1 << 18;   // a destination cell covering more bins is "big"
// CTAs of 128 threads per SM the query kernels are compiled for (register cap 48).  Measured on cfg5,
// query phase: 8 -> 0.55 ms, 10 -> 0.42 ms, 12 -> 0.51 ms, 16 -> 0.91 ms (spills + L1 thrashing).
#ifndef BP_QUERY_MINB
#define BP_QUERY_MINB 10
#endif
constexpr int BP_HEAVY = 256;           // a destination cell whose bins hold more entries than this is traversed by a whole warp
constexpr int BP_SLAB = 32;             // candidates per destination cell kept by the count pass (slab[k][cell])
constexpr int BP_SHORT_ROW = 48;        // rows with at most this many candidates are column-sorted by one thread
constexpr double BP_BIG_ANGLE = 0.2;    // rad; larger spherical cells are "big"
constexpr double BP_MIN_W = 0.05;       // a face is usable only if all vertices have w > this

struct BPStats {
    double sum_diam;
    unsigned long long count;
    unsigned long long lo[3], hi[3];   // order-preserving encodings of the bounding box of the vertices
                                       // (planar: x, y of every cell; sphere: x, y, z of the non-big cells)
    float max_diam;
    int pad;
};

struct BPParams {
    int dim;                 // 2 planar, 3 spherical
    int nfaces;              // 1 or 6
    int nbx, nby;            // bins per face
    double ox, oy;           // domain origin
    double hx, hy;           // domain upper corner
    double inv_hq;           // BP_SUB / bin size
    double eps;              // box inflation
    double big_chord;        // diameter threshold of "big" cells (chord length / planar length)
    float u_reject;          // a face cannot see a (non-big) cell whose first vertex has |u| or |v| above this
    float cull_lo[3], cull_hi[3];   // sphere: source cells whose first vertex lies outside this box cannot meet
                                    // any destination cell (destination-sharded builds see a slab of the globe)
};

struct QBox { int x0, x1, y0, y1; };

__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
inline double ordered_bits_to_double(unsigned long long b) {
    b = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// chord-length (sphere) / Euclidean (plane) diameter of a convex cell = max vertex distance
template <int DIM>
__device__ __forceinline__ float cell_diameter(const double *p, int n) {
    double best = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) { const double d = p[i * DIM + k] - p[j * DIM + k]; s += d * d; }
            best = s > best ? s : best;
        }
    return (float)(sqrt(best) * (1.0 + 1e-6));
}

// Quantised bounding box of a cell on `face`.  Returns false when the face cannot see the
// cell (a vertex behind/near the horizon, or the box misses the face's domain).  *clamped is
// set when part of the box lies outside the domain.
template <int DIM>
__device__ __forceinline__ bool cell_face_qbox(const double *p, int n, int face, const BPParams &P, QBox *q,
                                               bool *clamped) {
    double a0 = 1e300, a1 = -1e300, b0 = 1e300, b1 = -1e300;
    if (DIM == 3) {
        // face coordinates f(u) = u / sqrt(1 + u^2) = sin(atan u) of the gnomonic u = b/a, v = c/a: any
        // monotone map of u keeps "box of the projected vertices = box of the polygon"; this one is a
        // few FP32 instructions (the FP64 atan version made the binning kernels FP64-bound).  FP32
        // round-off (< 3e-7) is covered by the box inflation P.eps = 1e-6.
        const int ax = face >> 1;
        const float sg = (face & 1) ? -1.f : 1.f;
        const int bx = ax == 2 ? 0 : ax + 1, cx = bx == 2 ? 0 : bx + 1;
        {   // quick reject on the first vertex (three of the six faces fail the sign test, most of the
            // others this one): whenever the box test below accepts, every vertex lies within
            // 4 cell diameters of azimuth of the extended face domain
            const float w0 = sg * (float)p[ax];
            if (!(w0 > (float)BP_MIN_W)) return false;
            const float lim = P.u_reject * w0;
            if (fabsf((float)p[bx]) > lim || fabsf((float)p[cx]) > lim) return false;
        }
        float fa0 = 2.f, fa1 = -2.f, fb0 = 2.f, fb1 = -2.f;
        for (int i = 0; i < n; ++i) {
            const float w = sg * (float)p[3 * i + ax];
            if (!(w > (float)BP_MIN_W)) return false;
            const float u = (float)p[3 * i + bx] / w, v = (float)p[3 * i + cx] / w;
            const float al = u * rsqrtf(fmaf(u, u, 1.f)), be = v * rsqrtf(fmaf(v, v, 1.f));
            fa0 = fminf(fa0, al); fa1 = fmaxf(fa1, al);
            fb0 = fminf(fb0, be); fb1 = fmaxf(fb1, be);
        }
        a0 = fa0; a1 = fa1; b0 = fb0; b1 = fb1;
    } else {
        for (int i = 0; i < n; ++i) {
            const double al = p[2 * i], be = p[2 * i + 1];
            a0 = fmin(a0, al); a1 = fmax(a1, al);
            b0 = fmin(b0, be); b1 = fmax(b1, be);
        }
    }
    a0 -= P.eps; a1 += P.eps; b0 -= P.eps; b1 += P.eps;
    if (a1 < P.ox || a0 > P.hx || b1 < P.oy || b0 > P.hy) return false;
    *clamped = a0 < P.ox || a1 > P.hx || b0 < P.oy || b1 > P.hy;
    const int qxmax = P.nbx * BP_SUB - 1, qymax = P.nby * BP_SUB - 1;
    q->x0 = max(0, min(qxmax, (int)floor((a0 - P.ox) * P.inv_hq)));
    q->x1 = max(0, min(qxmax, (int)floor((a1 - P.ox) * P.inv_hq)));
    q->y0 = max(0, min(qymax, (int)floor((b0 - P.oy) * P.inv_hq)));
    q->y1 = max(0, min(qymax, (int)floor((b1 - P.oy) * P.inv_hq)));
    return true;
}

template <int DIM>
__device__ __forceinline__ const double *cell_ptr(const CellsView &g, int64_t c, int *n) {
    int64_t f;
    *n = cell_nverts<DIM>(g, c, &f);
    return g.verts + f * DIM;
}

// Cell access for the streaming kernels (one thread per cell).  A thread reading its own 96-byte
// quad touches 12 scattered 8-byte words per warp instruction; for fixed-stride quads each warp
// instead copies its 32 consecutive cells (3 KB, contiguous) with 16-byte coalesced loads into a
// padded shared-memory tile and every thread reads its cell from there.  Must be called by all
// threads of the block (blockDim.x <= 256); returns this thread's cell (shared or global memory).
constexpr int STAGE_THREADS = 256;
template <int DIM, int NT = STAGE_THREADS>
struct CellStage {
    static constexpr int CELL = 4 * DIM;          // doubles per quad
    static constexpr int PAD = CELL + 1;          // padded stride: conflict-free 8-byte reads
    double tile[NT * PAD];
};
template <int DIM, int NT>
__device__ __forceinline__ const double *stage_cell(const CellsView &g, int64_t c, int *n, CellStage<DIM, NT> &S) {
    constexpr int CELL = CellStage<DIM, NT>::CELL, PAD = CellStage<DIM, NT>::PAD, CHUNKS = CELL / 2;
    const bool fast = !g.off && g.nv == 4 && ((uintptr_t)g.verts % 16 == 0);     // block-uniform
    if (!fast) {
        if (c >= g.ncells) { *n = 0; return g.verts; }
        return cell_ptr<DIM>(g, c, n);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t c0 = c - lane;                  // first cell of this warp
    const int64_t nvalid = g.ncells - c0;         // cells of this warp inside the grid (may be <= 0)
    double *w = S.tile + wid * 32 * PAD;
    const double2 *src = reinterpret_cast<const double2 *>(g.verts + c0 * CELL);
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
        const int q = i * 32 + lane;              // 16-byte chunk index inside the warp's 32 cells
        const int cell = q / CHUNKS, part = q % CHUNKS;
        if (cell < nvalid) {
            const double2 v = __ldg(src + q);
            w[cell * PAD + 2 * part] = v.x;
            w[cell * PAD + 2 * part + 1] = v.y;
        }
    }
    __syncwarp();
    *n = c < g.ncells ? 4 : 0;
    return w + lane * PAD;
}

// --- bounds: per-cell diameter + grid statistics -----------------------------------------
// K4 + K1 in one pass over the vertices: geometric area (x scale) and orientation flag of every cell
// (regridder.jl:165-178), its diameter, and the grid statistics the bin grid is chosen from.
// Convexity of a ring (both clip operators here are convex-convex Sutherland-Hodgman, like the reference's
// ConvexConvexSutherlandHodgman on the sphere, regridder.jl:96-103; the reference's planar operator is
// Foster-Hormann, :87-94, which also takes non-convex rings -- those are DETECTED here so that the build
// fails with CRG_ERR_UNSUPPORTED instead of returning a wrong area).  Every vertex must lie on the inner
// side of every edge, up to round-off: h_e(w) * orientation >= -tol * |n_e| * diameter.
template <int DIM>
__device__ __forceinline__ bool ring_is_convex(const double *p, int n, bool clockwise, double diam) {
    if (n <= 3) return true;
    const double sg = clockwise ? -1.0 : 1.0;
    // A quadrilateral is convex iff its four turns have the same sign: vertex e + 2 against edge e is enough
    // (4 tests instead of 16); longer rings are tested vertex by vertex (same-sign turns could wind twice).
    const int ntest = n == 4 ? 1 : n;
    for (int e = 0; e < n; ++e) {
        const double *u = p + DIM * e, *v = p + DIM * (e + 1 == n ? 0 : e + 1);
        double nx, ny, nz = 0.0, h0 = 0.0;
        if (DIM == 3) {
            edge_normal(u, v, nx, ny, nz);
        } else {
            nx = -(v[1] - u[1]); ny = v[0] - u[0];
            h0 = -(nx * u[0] + ny * u[1]);
        }
        // tolerance: 1e-9 of the cell size + the round-off of h itself (~eps for unit vectors, ~eps |n| max|coordinate|
        // in the plane); the length of n in single precision is plenty for that
        const double nn = (double)sqrtf((float)(nx * nx + ny * ny + nz * nz)) * 1.000001;
        const double tol = DIM == 3 ? 1e-9 * nn * diam + 1e-14
                                    : nn * (1e-9 * diam + 1e-14 * (fabs(u[0]) + fabs(u[1]) + fabs(v[0]) + fabs(v[1])));
        for (int t = 0; t < ntest; ++t) {
            const int i = n == 4 ? ((e + 2) & 3) : t;
            const double *w = p + DIM * i;
            const double h = DIM == 3 ? nx * w[0] + ny * w[1] + nz * w[2] : nx * w[0] + ny * w[1] + h0;
            if (sg * h < -tol) return false;
        }
    }
    return true;
}

template <int DIM>
__global__ void __launch_bounds__(256) bp_bounds_kernel(CellsView g, float *__restrict__ diam, BPStats *st,
                                                        float big_chord, double scale, double *__restrict__ areas,
                                                        uint8_t *__restrict__ flip, unsigned int *__restrict__ nflip,
                                                        double *__restrict__ normals = nullptr) {
    __shared__ CellStage<DIM> stage;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // bounding box of the vertices: exact (double) in the plane, where it becomes the bin-grid domain;
    // single precision on the sphere, where it only feeds the (inflated) culling box
    using BT = typename std::conditional<DIM == 2, double, float>::type;
    double sum = 0.0;
    BT lo[3] = {(BT)1e30, (BT)1e30, (BT)1e30}, hi[3] = {(BT)-1e30, (BT)-1e30, (BT)-1e30};
    float mx = 0.f;
    unsigned cnt = 0;
    int n;
    const double *p = stage_cell<DIM>(g, c, &n, stage);
    if (c < g.ncells) {
        // pp, nn: the cell -- for quadrilaterals a register copy and a literal 4, so that the loops of the helpers below
        // unroll and every vertex is read from the shared-memory stage once (with the run-time vertex count they re-read
        // it: 1 220 thread instructions per cell, shared-memory wavefronts 68 % of peak, 156 us for cfg5's 3.1 M cells)
        auto body = [&](const double *pp, const int nn) {
            const double a = polygon_signed_area<DIM>(pp, nn);
            areas[c] = fabs(a) * scale;
            const bool f = a < 0.0;
            flip[c] = f ? 1 : 0;
            if (f) atomicAdd(nflip, 1u);
            const float d = cell_diameter<DIM>(pp, nn);
            diam[c] = d;
            if (!ring_is_convex<DIM>(pp, nn, f, (double)d)) atomicAdd(nflip + 2, 1u);   // nflip[2..3]: non-convex cells
            if (DIM == 3 && normals && nn == 4) {
                // Edge-plane normals of a spherical quadrilateral for the clip kernel (the grid that clips: 12 doubles per cell,
                // negated for a clockwise cell so that "inside" is n . x >= 0): the same edge_normal the clip would evaluate per
                // candidate pair -- 36 FP64 operations per pair there, once per cell here (the convexity test needs them anyway).
                double en[12];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    edge_normal(pp + 3 * e, pp + 3 * ((e + 1) & 3), en[3 * e], en[3 * e + 1], en[3 * e + 2]);
                    if (f) { en[3 * e] = -en[3 * e]; en[3 * e + 1] = -en[3 * e + 1]; en[3 * e + 2] = -en[3 * e + 2]; }
                }
                double2 *o = reinterpret_cast<double2 *>(normals + c * 12);
#pragma unroll
                for (int i = 0; i < 6; ++i) o[i] = make_double2(en[2 * i], en[2 * i + 1]);
            }
            if (DIM == 2 || d < big_chord) {
                sum = d; mx = d; cnt = 1;
                for (int i = 0; i < nn; ++i)
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const BT x = (BT)pp[DIM * i + k];
                        lo[k] = x < lo[k] ? x : lo[k];
                        hi[k] = x > hi[k] ? x : hi[k];
                    }
            }
        };
        if (n == 4) {
            double q[4 * DIM];
#pragma unroll
            for (int k = 0; k < 4 * DIM; ++k) q[k] = p[k];
            body(q, 4);
        } else {
            body(p, n);
        }
    }
    sum = warp_sum(sum); mx = warp_max(mx); cnt = warp_sum(cnt);
#pragma unroll
    for (int k = 0; k < DIM; ++k) { lo[k] = warp_min(lo[k]); hi[k] = warp_max(hi[k]); }
    // one set of atomics per block (same-address atomics from every warp serialised at the L2)
    __shared__ double r_sum[8];
    __shared__ BT r_lo[3][8], r_hi[3][8];
    __shared__ float r_mx[8];
    __shared__ unsigned r_cnt[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        r_sum[wid] = sum; r_mx[wid] = mx; r_cnt[wid] = cnt;
#pragma unroll
        for (int k = 0; k < DIM; ++k) { r_lo[k][wid] = lo[k]; r_hi[k][wid] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; ++w) {
            sum += r_sum[w]; mx = fmaxf(mx, r_mx[w]); cnt += r_cnt[w];
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                lo[k] = r_lo[k][w] < lo[k] ? r_lo[k][w] : lo[k];
                hi[k] = r_hi[k][w] > hi[k] ? r_hi[k][w] : hi[k];
            }
        }
        if (cnt) {
            atomicAdd(&st->sum_diam, sum);
            atomicAdd(&st->count, (unsigned long long)cnt);
            atomic_max_pos_float(&st->max_diam, mx);
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                atomicMin(&st->lo[k], ordered_bits((double)lo[k]));
                atomicMax(&st->hi[k], ordered_bits((double)hi[k]));
            }
        }
    }
}

// --- source side: count / fill bins ----------------------------------------------------
// entry = {cell id, x0 | x1 << 16, y0 | y1 << 16, 0}
// The count pass (FILL = false) also leaves a 16-byte record per cell, {x0 | x1 << 16, y0 | y1 << 16, face | nfaces << 8, 0}:
// the quantised box on the one face that sees the cell (nfaces = 0: the cell is not binned at all -- ghost, culled,
// big; nfaces >= 2: a cell near a cube edge, its boxes are recomputed).  bp_bin_fill_kernel inserts from the records
// without touching the vertices again (the projection of 3.1 M source cells a second time was 183 us of cfg5's build).
template <int DIM, bool FILL>
__global__ void __launch_bounds__(256) bp_bin_kernel(CellsView g, const float *__restrict__ diam, BPParams P,
                                                     uint32_t *__restrict__ bin_count /* or cursor */,
                                                     const uint32_t *__restrict__ bin_start,
                                                     int4 *__restrict__ entries, int32_t *__restrict__ big_list,
                                                     uint32_t *__restrict__ big_counter, int4 *__restrict__ rec = nullptr) {
    __shared__ CellStage<DIM> stage;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n;
    const double *p = stage_cell<DIM>(g, c, &n, stage);
    if (c >= g.ncells) return;
    if (!FILL && rec) rec[c] = make_int4(0, 0, 0, 0);
    // a cell whose vertices all coincide is a GHOST (the padding polygon of a tripolar fold row): it has no area and
    // is never a candidate, like the cells the reference keeps out of its tree (OceananigansExt.jl:66-75,141-160)
    if (diam[c] == 0.f) return;
    bool big = DIM == 3 && !(diam[c] < (float)P.big_chord);
    if (DIM == 3 && !big) {
        const float x = (float)p[0], y = (float)p[1], z = (float)p[2];
        if (x < P.cull_lo[0] || x > P.cull_hi[0] || y < P.cull_lo[1] || y > P.cull_hi[1] || z < P.cull_lo[2] || z > P.cull_hi[2])
            return;
    }
    unsigned faces = 0;          // faces that see the cell
    QBox b1;                     // the box on the (last) face that sees it: kept when it is the only one
    if (!big) {     // total number of bins the cell would be inserted in (with two or three faces the boxes
        int cover = 0;  // are recomputed below: cheaper than keeping them in local memory)
        for (int f = 0; f < P.nfaces; ++f) {
            QBox b;
            bool cl;
            if (cell_face_qbox<DIM>(p, n, f, P, &b, &cl)) {
                cover += ((b.x1 >> 4) - (b.x0 >> 4) + 1) * ((b.y1 >> 4) - (b.y0 >> 4) + 1);
                faces |= 1u << f;
                b1 = b;
            }
        }
        big = cover > BP_MAX_COVER;
    }
    if (big) {
        if (!FILL) big_list[atomicAdd(big_counter, 1u)] = (int32_t)c;
        return;
    }
    const int nfaces = __popc(faces);
    while (faces) {
        const int f = __ffs(faces) - 1;
        faces &= faces - 1u;
        QBox b = b1;
        bool cl;
        if (nfaces > 1 && !cell_face_qbox<DIM>(p, n, f, P, &b, &cl)) continue;
        const int4 e = make_int4((int)c, b.x0 | (b.x1 << 16), b.y0 | (b.y1 << 16), 0);
        if (!FILL && rec) rec[c] = make_int4(e.y, e.z, f | (nfaces << 8), 0);
        for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by)
            for (int bx = b.x0 >> 4; bx <= (b.x1 >> 4); ++bx) {
                const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
                const uint32_t k = atomicAdd(&bin_count[bin], 1u);
                if (FILL) entries[bin_start[bin] + k] = e;
            }
    }
}

// Fill pass from the records of the count pass.
template <int DIM>
__global__ void __launch_bounds__(256) bp_bin_fill_kernel(CellsView g, BPParams P, const int4 *__restrict__ rec,
                                                          uint32_t *__restrict__ cursor, const uint32_t *__restrict__ bin_start,
                                                          int4 *__restrict__ entries) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    const int4 r = rec[c];
    const int nfaces = (r.z >> 8) & 0xff;
    if (nfaces == 0) return;
    if (nfaces == 1) {
        const int f = r.z & 0xff;
        const int x0 = r.x & 0xffff, x1 = (r.x >> 16) & 0xffff, y0 = r.y & 0xffff, y1 = (r.y >> 16) & 0xffff;
        const int4 e = make_int4((int)c, r.x, r.y, 0);
        for (int by = y0 >> 4; by <= (y1 >> 4); ++by)
            for (int bx = x0 >> 4; bx <= (x1 >> 4); ++bx) {
                const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
                entries[bin_start[bin] + atomicAdd(&cursor[bin], 1u)] = e;
            }
        return;
    }
    int n;                                     // a cell near a cube edge: its boxes again, from the vertices
    const double *p = cell_ptr<DIM>(g, c, &n);
    for (int f = 0; f < P.nfaces; ++f) {
        QBox b;
        bool cl;
        if (!cell_face_qbox<DIM>(p, n, f, P, &b, &cl)) continue;
        const int4 e = make_int4((int)c, b.x0 | (b.x1 << 16), b.y0 | (b.y1 << 16), 0);
        for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by)
            for (int bx = b.x0 >> 4; bx <= (b.x1 >> 4); ++bx) {
                const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
                entries[bin_start[bin] + atomicAdd(&cursor[bin], 1u)] = e;
            }
    }
}

// --- destination side: count / fill candidate pairs ----------------------------------------
__device__ __forceinline__ int home_face(const double *p, int n) {
    double cx = 0, cy = 0, cz = 0;
    for (int i = 0; i < n; ++i) { cx += p[3 * i]; cy += p[3 * i + 1]; cz += p[3 * i + 2]; }
    const double ax = fabs(cx), ay = fabs(cy), az = fabs(cz);
    if (ax >= ay && ax >= az) return cx >= 0 ? 0 : 1;
    if (ay >= az) return cy >= 0 ? 2 : 3;
    return cz >= 0 ? 4 : 5;
}

// FILL = false: cand_count[d] = number of candidates of destination cell d; big destination
//               cells (count = n_src) are appended to big_dst; big_dst_counter[1] is set when some
//               cell has more than BP_SHORT_ROW candidates, big_dst_counter[2] when some cell has more
//               than BP_SLAB (the first BP_SLAB candidates of every cell are kept in `slab`).
// FILL = true : pairs[cand_off[d] + k] = (src, d); big destination cells are skipped (filled by
//               bp_fill_big_dst_kernel).
template <int DIM, bool FILL>
__global__ void __launch_bounds__(128, BP_QUERY_MINB) bp_query_kernel(CellsView g, const float *__restrict__ diam, BPParams P,
                                                       const uint32_t *__restrict__ bin_start,
                                                       const int4 *__restrict__ entries,
                                                       const int32_t *__restrict__ big_src, int n_big_src,
                                                       int64_t n_src, uint32_t *__restrict__ cand_count,
                                                       const int64_t *__restrict__ cand_off,
                                                       int2 *__restrict__ pairs, int32_t *__restrict__ big_dst,
                                                       uint32_t *__restrict__ big_dst_counter,
                                                       int32_t *__restrict__ slab, int32_t *__restrict__ heavy_list,
                                                       uint32_t *__restrict__ heavy_counter) {
    __shared__ CellStage<DIM, 128> stage;
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n;
    const double *p = stage_cell<DIM>(g, d, &n, stage);
    if (d >= g.ncells) return;
    if (diam[d] == 0.f) {                    // ghost destination cell (see bp_bin_kernel): no candidates
        if (!FILL) cand_count[d] = 0u;
        return;
    }
    bool big = DIM == 3 && !(diam[d] < (float)P.big_chord);
    QBox b;
    int f = 0;
    if (!big) {
        f = DIM == 3 ? home_face(p, n) : 0;
        bool cl = false;
        const bool ok = cell_face_qbox<DIM>(p, n, f, P, &b, &cl);
        big = !ok || cl;
        if (!big) {
            const int64_t cover = (int64_t)((b.x1 >> 4) - (b.x0 >> 4) + 1) * ((b.y1 >> 4) - (b.y0 >> 4) + 1);
            big = cover > BP_MAX_QUERY;
        }
    }
    if (big) {
        if (!FILL) {
            cand_count[d] = (uint32_t)n_src;
            big_dst[atomicAdd(big_dst_counter, 1u)] = (int32_t)d;
            if (n_src > BP_SHORT_ROW) big_dst_counter[1] = 1u;
            if (slab) slab[d] = -1;                           // marker: filled by bp_fill_big_dst_kernel
        }
        return;
    }
    const int64_t nd = g.ncells;
    // Where hundreds of source cells share a bin (the pole of a lon-lat source), one thread walking them all is the
    // tail of the whole kernel (config 1: 0.57 of the 1.0 ms build): such cells go to a list and are traversed by a
    // warp each in bp_query_heavy_kernel (the count pass decides; the fill pass skips them the same way).
    if (heavy_list) {
        uint32_t work = 0;
        for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by) {
            const size_t row = ((size_t)f * P.nby + by) * P.nbx;
            work += bin_start[row + (b.x1 >> 4) + 1] - bin_start[row + (b.x0 >> 4)];      // bins of a row are consecutive
        }
        if (work > (uint32_t)BP_HEAVY) {
            if (!FILL) heavy_list[atomicAdd(heavy_counter, 1u)] = (int32_t)d;
            return;
        }
    }
    uint32_t cnt = 0;
    int2 *out = FILL ? pairs + cand_off[d] : nullptr;
    for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by)
        for (int bx = b.x0 >> 4; bx <= (b.x1 >> 4); ++bx) {
            const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
            const uint32_t lo = bin_start[bin], hi = bin_start[bin + 1];
            for (uint32_t k = lo; k < hi; ++k) {
                const int4 e = __ldg(&entries[k]);
                const int sx0 = e.y & 0xffff, sx1 = (e.y >> 16) & 0xffff;
                const int sy0 = e.z & 0xffff, sy1 = (e.z >> 16) & 0xffff;
                if (sx0 > b.x1 || b.x0 > sx1 || sy0 > b.y1 || b.y0 > sy1) continue;
                // report the pair only in the first bin common to both boxes
                if ((max(sx0, b.x0) >> 4) != bx || (max(sy0, b.y0) >> 4) != by) continue;
                if (FILL) out[cnt] = make_int2(e.x, (int)d);
                else if (slab && cnt < BP_SLAB) slab[(size_t)cnt * nd + d] = e.x;
                ++cnt;
            }
        }
    for (int k = 0; k < n_big_src; ++k) {
        if (FILL) out[cnt] = make_int2(big_src[k], (int)d);
        else if (slab && cnt < BP_SLAB) slab[(size_t)cnt * nd + d] = big_src[k];
        ++cnt;
    }
    if (!FILL) {
        cand_count[d] = cnt;
        if (cnt > BP_SHORT_ROW) big_dst_counter[1] = 1u;      // flag: some row is long (benign race, same value)
        if (cnt > BP_SLAB) big_dst_counter[2] = 1u;           // flag: the slab does not hold every candidate
    }
}

// One WARP per heavy destination cell (see bp_query_kernel): the lanes stride over the entries of its bins; a ballot
// ranks the candidates, so the list order is the order of the entries (deterministic).
template <int DIM, bool FILL>
__global__ void __launch_bounds__(128) bp_query_heavy_kernel(CellsView g, BPParams P, const uint32_t *__restrict__ bin_start,
                                                             const int4 *__restrict__ entries,
                                                             const int32_t *__restrict__ big_src, int n_big_src,
                                                             uint32_t *__restrict__ cand_count,
                                                             const int64_t *__restrict__ cand_off, int2 *__restrict__ pairs,
                                                             uint32_t *__restrict__ flags /* [1] long row, [2] slab overflow */,
                                                             int32_t *__restrict__ slab, const int32_t *__restrict__ heavy_list,
                                                             const uint32_t *__restrict__ heavy_counter) {
    const int lane = threadIdx.x & 31;
    const uint32_t nheavy = *heavy_counter;
    const int64_t nd = g.ncells;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); h < nheavy; h += gridDim.x * (blockDim.x >> 5)) {
        const int64_t d = heavy_list[h];
        int n;
        const double *p = cell_ptr<DIM>(g, d, &n);
        const int f = DIM == 3 ? home_face(p, n) : 0;
        QBox b;
        bool cl = false;
        cell_face_qbox<DIM>(p, n, f, P, &b, &cl);            // (accepted by bp_query_kernel already)
        uint32_t cnt = 0;
        int2 *out = FILL ? pairs + cand_off[d] : nullptr;
        for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by)
            for (int bx = b.x0 >> 4; bx <= (b.x1 >> 4); ++bx) {
                const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
                const uint32_t lo = bin_start[bin], hi = bin_start[bin + 1];
                for (uint32_t k0 = lo; k0 < hi; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    bool hit = false;
                    int src = 0;
                    if (k < hi) {
                        const int4 e = __ldg(&entries[k]);
                        const int sx0 = e.y & 0xffff, sx1 = (e.y >> 16) & 0xffff;
                        const int sy0 = e.z & 0xffff, sy1 = (e.z >> 16) & 0xffff;
                        hit = !(sx0 > b.x1 || b.x0 > sx1 || sy0 > b.y1 || b.y0 > sy1) &&
                              (max(sx0, b.x0) >> 4) == bx && (max(sy0, b.y0) >> 4) == by;
                        src = e.x;
                    }
                    const unsigned m = __ballot_sync(CRG_FULL, hit);
                    if (hit) {
                        const uint32_t pos = cnt + __popc(m & lt);
                        if (FILL) out[pos] = make_int2(src, (int)d);
                        else if (slab && pos < (uint32_t)BP_SLAB) slab[(size_t)pos * nd + d] = src;
                    }
                    cnt += __popc(m);
                }
            }
        for (int k0 = 0; k0 < n_big_src; k0 += 32) {
            const int k = k0 + lane;
            if (k < n_big_src) {
                if (FILL) out[cnt + k] = make_int2(big_src[k], (int)d);
                else if (slab && cnt + k < (uint32_t)BP_SLAB) slab[(size_t)(cnt + k) * nd + d] = big_src[k];
            }
        }
        cnt += (uint32_t)n_big_src;
        if (!FILL && lane == 0) {
            cand_count[d] = cnt;
            if (cnt > BP_SHORT_ROW) flags[1] = 1u;
            if (cnt > BP_SLAB) flags[2] = 1u;
        }
    }
}

// The count pass keeps the first BP_SLAB candidates of every destination cell in slab[k][cell]
// (coalesced across cells); when no cell has more, the pair list is a copy of the slab and the
// second traversal of the bins is skipped.
// A warp takes 32 destination cells: their slab columns are read coalesced (slot by slot) into a shared-memory tile, then
// the run of every cell is written by the whole warp -- consecutive lanes, consecutive pairs (one thread writing its own
// run of ~13 pairs touched 32 sectors per store instruction: 90 us on cfg5 for 160 MB).
__global__ void __launch_bounds__(256) bp_fill_slab_kernel(const int32_t *__restrict__ slab, int64_t nd,
                                                           const uint32_t *__restrict__ cand_count,
                                                           const int64_t *__restrict__ cand_off,
                                                           int2 *__restrict__ pairs) {
    __shared__ int32_t s_tile[8][BP_SLAB][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = d < nd ? cand_count[d] : 0u;
    if (cnt && slab[d] < 0) cnt = 0;                          // big destination cell (filled by bp_fill_big_dst_kernel)
    const int64_t off = d < nd ? cand_off[d] : 0;
    const uint32_t maxcnt = warp_max(cnt);
    for (uint32_t k = 0; k < maxcnt; ++k)
        if (k < cnt) s_tile[wid][k][lane] = slab[(size_t)k * nd + d];
    __syncwarp();
    for (int L = 0; L < 32; ++L) {
        const uint32_t c = __shfl_sync(CRG_FULL, cnt, L);
        if (c == 0) continue;                                 // (warp-uniform)
        const int64_t o = __shfl_sync(CRG_FULL, off, L);
        const int dd = (int)(d - lane + L);
        if ((uint32_t)lane < c) pairs[o + lane] = make_int2(s_tile[wid][lane][L], dd);
    }
}

__global__ void __launch_bounds__(256) bp_fill_big_dst_kernel(const int32_t *__restrict__ big_dst,
                                                              const int64_t *__restrict__ cand_off,
                                                              int64_t n_src, int2 *__restrict__ pairs) {
    const int d = big_dst[blockIdx.y];
    int2 *out = pairs + cand_off[d];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_src; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = make_int2((int)j, d);
}

}  // namespace crg

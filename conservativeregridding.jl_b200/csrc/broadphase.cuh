// broadphase.cuh -- device-side candidate search (replaces the dual-tree DFS of
// /root/reference/src/utils/MultithreadedDualDepthFirstSearch.jl:12-65 and the bounding-cap
// machinery of src/trees/grids.jl:245-287).
//
// Spherical cells are binned on the six faces of a cube in *equiangular gnomonic*
// coordinates (alpha, beta) = (atan(b/a), atan(c/a)): the gnomonic map sends great-circle
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
#include "common.cuh"
#include "geom.cuh"

namespace crg {

constexpr int BP_SUB = 16;              // quantisation: sub-bins per bin
constexpr int BP_MAX_BINS_1D = 4095;    // so that quantised coordinates fit 16 bits
constexpr int BP_MAX_COVER = 4096;      // a source cell covering more bins is "big"
constexpr int BP_MAX_QUERY = 1 << 18;   // a destination cell covering more bins is "big"
constexpr double BP_BIG_ANGLE = 0.2;    // rad; larger spherical cells are "big"
constexpr double BP_MIN_W = 0.05;       // a face is usable only if all vertices have w > this

struct BPStats {
    double sum_diam;
    unsigned long long count;
    unsigned long long lo[2], hi[2];   // order-preserving encodings of the planar bounding box
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
        const int ax = face >> 1;
        const double sg = (face & 1) ? -1.0 : 1.0;
        const int bx = ax == 2 ? 0 : ax + 1, cx = bx == 2 ? 0 : bx + 1;
        for (int i = 0; i < n; ++i) {
            const double w = sg * p[3 * i + ax];
            if (!(w > BP_MIN_W)) return false;
            const double iw = 1.0 / w;
            const double al = atan(p[3 * i + bx] * iw), be = atan(p[3 * i + cx] * iw);
            a0 = fmin(a0, al); a1 = fmax(a1, al);
            b0 = fmin(b0, be); b1 = fmax(b1, be);
        }
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

// --- bounds: per-cell diameter + grid statistics -----------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) bp_bounds_kernel(CellsView g, float *__restrict__ diam, BPStats *st,
                                                        float big_chord) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sum = 0.0, lo0 = 1e300, lo1 = 1e300, hi0 = -1e300, hi1 = -1e300;
    float mx = 0.f;
    unsigned cnt = 0;
    if (c < g.ncells) {
        int n;
        const double *p = cell_ptr<DIM>(g, c, &n);
        const float d = cell_diameter<DIM>(p, n);
        diam[c] = d;
        if (DIM == 2 || d < big_chord) { sum = d; mx = d; cnt = 1; }
        if (DIM == 2)
            for (int i = 0; i < n; ++i) {
                lo0 = fmin(lo0, p[2 * i]); hi0 = fmax(hi0, p[2 * i]);
                lo1 = fmin(lo1, p[2 * i + 1]); hi1 = fmax(hi1, p[2 * i + 1]);
            }
    }
    sum = warp_sum(sum); mx = warp_max(mx); cnt = warp_sum(cnt);
    if (DIM == 2) { lo0 = warp_min(lo0); lo1 = warp_min(lo1); hi0 = warp_max(hi0); hi1 = warp_max(hi1); }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&st->sum_diam, sum);
        atomicAdd(&st->count, (unsigned long long)cnt);
        atomic_max_pos_float(&st->max_diam, mx);
        if (DIM == 2) {
            atomicMin(&st->lo[0], ordered_bits(lo0)); atomicMin(&st->lo[1], ordered_bits(lo1));
            atomicMax(&st->hi[0], ordered_bits(hi0)); atomicMax(&st->hi[1], ordered_bits(hi1));
        }
    }
}

// --- source side: count / fill bins ----------------------------------------------------
// entry = {cell id, x0 | x1 << 16, y0 | y1 << 16, 0}
template <int DIM, bool FILL>
__global__ void __launch_bounds__(256) bp_bin_kernel(CellsView g, const float *__restrict__ diam, BPParams P,
                                                     uint32_t *__restrict__ bin_count /* or cursor */,
                                                     const uint32_t *__restrict__ bin_start,
                                                     int4 *__restrict__ entries, int32_t *__restrict__ big_list,
                                                     uint32_t *__restrict__ big_counter) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    int n;
    const double *p = cell_ptr<DIM>(g, c, &n);
    bool big = DIM == 3 && !(diam[c] < (float)P.big_chord);
    QBox box[6];
    bool ok[6];
    if (!big) {
        int cover = 0;
        for (int f = 0; f < P.nfaces; ++f) {
            bool cl;
            ok[f] = cell_face_qbox<DIM>(p, n, f, P, &box[f], &cl);
            if (ok[f]) cover += ((box[f].x1 >> 4) - (box[f].x0 >> 4) + 1) * ((box[f].y1 >> 4) - (box[f].y0 >> 4) + 1);
        }
        big = cover > BP_MAX_COVER;
    }
    if (big) {
        if (!FILL) big_list[atomicAdd(big_counter, 1u)] = (int32_t)c;
        return;
    }
    for (int f = 0; f < P.nfaces; ++f) {
        if (!ok[f]) continue;
        const QBox b = box[f];
        const int4 e = make_int4((int)c, b.x0 | (b.x1 << 16), b.y0 | (b.y1 << 16), 0);
        for (int by = b.y0 >> 4; by <= (b.y1 >> 4); ++by)
            for (int bx = b.x0 >> 4; bx <= (b.x1 >> 4); ++bx) {
                const size_t bin = ((size_t)f * P.nby + by) * P.nbx + bx;
                const uint32_t k = atomicAdd(&bin_count[bin], 1u);
                if (FILL) entries[bin_start[bin] + k] = e;
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
//               cells (count = n_src) are appended to big_dst.
// FILL = true : pairs[cand_off[d] + k] = (src, d); big destination cells are skipped (filled by
//               bp_fill_big_dst_kernel).
template <int DIM, bool FILL>
__global__ void __launch_bounds__(128) bp_query_kernel(CellsView g, const float *__restrict__ diam, BPParams P,
                                                       const uint32_t *__restrict__ bin_start,
                                                       const int4 *__restrict__ entries,
                                                       const int32_t *__restrict__ big_src, int n_big_src,
                                                       int64_t n_src, uint32_t *__restrict__ cand_count,
                                                       const int64_t *__restrict__ cand_off,
                                                       int2 *__restrict__ pairs, int32_t *__restrict__ big_dst,
                                                       uint32_t *__restrict__ big_dst_counter) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= g.ncells) return;
    int n;
    const double *p = cell_ptr<DIM>(g, d, &n);
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
        }
        return;
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
                ++cnt;
            }
        }
    for (int k = 0; k < n_big_src; ++k) {
        if (FILL) out[cnt] = make_int2(big_src[k], (int)d);
        ++cnt;
    }
    if (!FILL) cand_count[d] = cnt;
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

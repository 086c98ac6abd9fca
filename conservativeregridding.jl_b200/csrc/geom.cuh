// geom.cuh -- FP64 polygon geometry on the unit sphere / plane (device).
//
// Replaces, per candidate pair, DefaultIntersectionOperator
// (/root/reference/src/regridder/regridder.jl:87-103): GO.intersection with
// ConvexConvexSutherlandHodgman (great-circle half-space cuts) followed by GO.area, and per
// cell GO.area (regridder.jl:165-178).  The arithmetic itself lives in GeometryOps.jl (not
// in the reference tree); this is the published algorithm in two device forms: clip_pair_area
// (any convex rings, working polygon ping-ponged through shared memory) and quad_prepass /
// quad_cut_area (quadrilaterals: symbolic polygon over a per-thread point table).
#pragma once
#include "common.cuh"

namespace crg {

struct CellsView {
    const double *verts;     // [ncells][nv][DIM] or ragged
    const int32_t *off;      // nullptr => fixed nv
    const uint8_t *flip;     // nullptr => every ring already CCW; else 1 = stored clockwise
    int64_t ncells;
    int nv;
};

struct d3 { double x, y, z; };

__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ d3 cross(d3 a, d3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Normal u x v of the great circle through an edge, WITHOUT fused multiply-adds: with contraction,
// fma(u1, v2, -rn(u2 v1)) of a zero-length edge (u == v bit for bit: the pole corners of a lon-lat cell) is the
// rounding error of a product -- a 1e-17 vector of arbitrary direction that then "separates" or "cuts" cells --
// unless the coordinates happen to be exact (unrotated grids).  Products rounded separately cancel exactly.
__device__ __forceinline__ void edge_normal(const double *u, const double *v, double &nx, double &ny, double &nz) {
    nx = __dsub_rn(__dmul_rn(u[1], v[2]), __dmul_rn(u[2], v[1]));
    ny = __dsub_rn(__dmul_rn(u[2], v[0]), __dmul_rn(u[0], v[2]));
    nz = __dsub_rn(__dmul_rn(u[0], v[1]), __dmul_rn(u[1], v[0]));
}

// Accumulates spherical-excess half-angles as a complex product: the triangle (a, b, c)
// contributes the angle of z = (1 + a.b + b.c + c.a) + i a.((b-a) x (c-a)), i.e. E/2, and
// sum of angles = angle of the product -- one atan2 per polygon instead of one per
// triangle.  The product is flushed through atan2 whenever its angle could leave
// (-pi/2, pi/2), so arbitrarily large polygons stay exact.
// atan2(im, re).  Grid cells are small: the accumulated half-excess is almost always a tiny angle, for
// which the odd series through x^13 is exact to double precision (|x| <= 1/32: next term < 2^-73 x);
// everything else takes the library routine.
__device__ __forceinline__ double atan2_general(double im, double re) { return atan2(im, re); }
__device__ __forceinline__ double atan2_small(double im, double re) {
    if (re > 0.0 && fabs(im) <= 0.03125 * re) {
        const double x = im / re, t = x * x;
        double p = fma(t, 1.0 / 13.0, -1.0 / 11.0);
        p = fma(p, t, 1.0 / 9.0);
        p = fma(p, t, -1.0 / 7.0);
        p = fma(p, t, 1.0 / 5.0);
        p = fma(p, t, -1.0 / 3.0);
        return fma(p * t, x, x);
    }
    return atan2_general(im, re);
}

struct ExcessAcc {
    double re = 1.0, im = 0.0, total = 0.0;
    __device__ __forceinline__ void flush() {
        total += atan2_small(im, re);
        re = 1.0; im = 0.0;
    }
    __device__ __forceinline__ void add_triangle(d3 a, d3 b, d3 c) {
        const double det = dot(a, cross(b - a, c - a));
        const double den = 1.0 + dot(a, b) + dot(b, c) + dot(c, a);
        if (!(den > 0.0)) {          // huge triangle: angle outside (-pi/2, pi/2)
            flush();
            total += atan2_general(det, den);
            return;
        }
        const double nre = re * den - im * det;
        const double nim = re * det + im * den;
        re = nre; im = nim;
        if (!(re > 0.0) || re > 1e200) flush();
    }
    __device__ __forceinline__ double area() {   // signed: CCW seen from outside positive
        flush();
        return 2.0 * total;
    }
};

// ------------------------------------------------------------------------------------
// cell access
// ------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ int cell_nverts(const CellsView &g, int64_t c, int64_t *first) {
    if (g.off) { *first = g.off[c]; return g.off[c + 1] - g.off[c]; }
    *first = c * g.nv;
    return g.nv;
}

// signed area of a ring as stored (no orientation fix-up)
template <int DIM>
__device__ double polygon_signed_area(const double *p, int n) {
    if (DIM == 3) {
        ExcessAcc acc;
        const d3 a = {p[0], p[1], p[2]};
        d3 b = {p[3], p[4], p[5]};
        for (int i = 2; i < n; ++i) {
            const d3 cc = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
            acc.add_triangle(a, b, cc);
            b = cc;
        }
        return acc.area();
    } else {
        double s = 0.0;
        const double x0 = p[0], y0 = p[1];
        for (int i = 1; i + 1 < n; ++i) {
            const double ax = p[2 * i] - x0, ay = p[2 * i + 1] - y0;
            const double bx = p[2 * i + 2] - x0, by = p[2 * i + 3] - y0;
            s += ax * by - ay * bx;
        }
        return 0.5 * s;
    }
}

// ------------------------------------------------------------------------------------
// Sutherland-Hodgman clip + area of one (subject, clip) pair.
//
// Shared-memory working polygon: slot v, coordinate c of thread t lives at
//   buf[((v * DIM) + c) * NT + t]            (NT = threads per block)
// so a warp touches consecutive doubles -- conflict-free -- and nothing spills to local
// memory.  MAXW = capacity of the working polygon (subject + clip vertex counts).
// ------------------------------------------------------------------------------------
template <int DIM, int NT, int MAXW>
struct PolyBuf {
    double *b;   // this thread's column: b[(v*DIM + c) * NT]
    __device__ __forceinline__ double get(int v, int c) const { return b[(v * DIM + c) * NT]; }
    __device__ __forceinline__ void set(int v, int c, double x) { b[(v * DIM + c) * NT] = x; }
};

// Returns the (non-negative up to round-off) area of subject ∩ clip on the unit sphere
// (DIM == 3) or in the plane (DIM == 2); 0 when the intersection has fewer than 3 vertices.
template <int DIM, int NT, int MAXW>
__device__ double clip_pair_area(const CellsView &gs, int64_t s, const CellsView &gc, int64_t c,
                                 double *smem /* 2 * MAXW * DIM * NT doubles */) {
    PolyBuf<DIM, NT, MAXW> cur{smem + threadIdx.x};
    PolyBuf<DIM, NT, MAXW> nxt{smem + MAXW * DIM * NT + threadIdx.x};

    int64_t sf, cf;
    const int ns = cell_nverts<DIM>(gs, s, &sf);
    const int nc = cell_nverts<DIM>(gc, c, &cf);
    const double *sp = gs.verts + sf * DIM;
    const double *cp = gc.verts + cf * DIM;
    const bool sflip = gs.flip && gs.flip[s];
    const bool cflip = gc.flip && gc.flip[c];

    // subject ring -> shared memory, oriented CCW
    for (int i = 0; i < ns; ++i) {
        const double *q = sp + (sflip ? (ns - 1 - i) : i) * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) cur.set(i, k, q[k]);
    }
    int m = ns;

    for (int e = 0; e < nc && m > 0; ++e) {
        const int iu = cflip ? (nc - 1 - e) : e;
        const int iv = cflip ? (iu == 0 ? nc - 1 : iu - 1) : (e + 1 == nc ? 0 : e + 1);
        const double *u = cp + iu * DIM, *v = cp + iv * DIM;
        // half-space of the directed edge u -> v: inside <=> h(p) >= 0
        double nx, ny, nz = 0.0, h0 = 0.0;
        if (DIM == 3) {
            edge_normal(u, v, nx, ny, nz);
            if (nx == 0.0 && ny == 0.0 && nz == 0.0) continue;   // zero-length edge (pole cells)
        } else {
            const double ex = v[0] - u[0], ey = v[1] - u[1];
            if (ex == 0.0 && ey == 0.0) continue;
            nx = -ey; ny = ex;                    // h(p) = ex*(py-uy) - ey*(px-ux)
            h0 = -(nx * u[0] + ny * u[1]);
        }
        auto hval = [&](int i) -> double {
            if (DIM == 3) return fma(nx, cur.get(i, 0), fma(ny, cur.get(i, 1), nz * cur.get(i, 2)));
            return fma(nx, cur.get(i, 0), fma(ny, cur.get(i, 1), h0));
        };
        int mo = 0;
        double dp = hval(m - 1);
        double px = cur.get(m - 1, 0), py = cur.get(m - 1, 1), pz = DIM == 3 ? cur.get(m - 1, 2) : 0.0;
        for (int i = 0; i < m; ++i) {
            const double qx = cur.get(i, 0), qy = cur.get(i, 1), qz = DIM == 3 ? cur.get(i, 2) : 0.0;
            const double dq = hval(i);
            const bool in_p = dp >= 0.0, in_q = dq >= 0.0;
            if (in_p != in_q) {
                const double t = dp / (dp - dq);
                double rx = fma(t, qx - px, px), ry = fma(t, qy - py, py), rz = fma(t, qz - pz, pz);
                if (DIM == 3) {
                    const double inv = rsqrt(rx * rx + ry * ry + rz * rz);
                    rx *= inv; ry *= inv; rz *= inv;
                }
                if (mo < MAXW) {
                    nxt.set(mo, 0, rx); nxt.set(mo, 1, ry);
                    if (DIM == 3) nxt.set(mo, 2, rz);
                    ++mo;
                }
            }
            if (in_q && mo < MAXW) {
                nxt.set(mo, 0, qx); nxt.set(mo, 1, qy);
                if (DIM == 3) nxt.set(mo, 2, qz);
                ++mo;
            }
            dp = dq; px = qx; py = qy; pz = qz;
        }
        m = mo;
        double *t = cur.b; cur.b = nxt.b; nxt.b = t;
    }
    if (m < 3) return 0.0;

    if (DIM == 3) {
        ExcessAcc acc;
        const d3 a = {cur.get(0, 0), cur.get(0, 1), cur.get(0, 2)};
        d3 b = {cur.get(1, 0), cur.get(1, 1), cur.get(1, 2)};
        for (int i = 2; i < m; ++i) {
            const d3 cc = {cur.get(i, 0), cur.get(i, 1), cur.get(i, 2)};
            acc.add_triangle(a, b, cc);
            b = cc;
        }
        return acc.area();
    } else {
        double sarea = 0.0;
        const double x0 = cur.get(0, 0), y0 = cur.get(0, 1);
        double ax = cur.get(1, 0) - x0, ay = cur.get(1, 1) - y0;
        for (int i = 2; i < m; ++i) {
            const double bx = cur.get(i, 0) - x0, by = cur.get(i, 1) - y0;
            sarea += ax * by - ay * bx;
            ax = bx; ay = by;
        }
        return 0.5 * sarea;
    }
}


// ------------------------------------------------------------------------------------
// Fast path: quadrilateral x quadrilateral (every structured grid of BASELINE.json): quad_prepass +
// quad_cut_area below, driven by clip_quad_kernel (kernels.cuh).  Cells are fetched with 16-byte loads.
// ------------------------------------------------------------------------------------
// 256-bit loads (sm_100: LDG.256, three per spherical quadrilateral instead of six 128-bit ones -- a gather at a 96-byte
// stride costs one L1 wavefront per lane and instruction, and the L1 data pipe is the clip kernel's busiest unit);
// p must be 32-byte aligned.
__device__ __forceinline__ void load_quad_wide(const double *__restrict__ p, double (&v)[4][3]) {
    double t[12];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(t[4 * i]), "=d"(t[4 * i + 1]), "=d"(t[4 * i + 2]), "=d"(t[4 * i + 3]) : "l"(p + 4 * i));
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) v[i][k] = t[i * 3 + k];
}

template <int DIM>
__device__ __forceinline__ void load_quad(const double *__restrict__ p, bool flip, double (&v)[4][DIM]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double t[4 * DIM];
#pragma unroll
    for (int i = 0; i < 2 * DIM; ++i) { const double2 d = __ldg(q + i); t[2 * i] = d.x; t[2 * i + 1] = d.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < DIM; ++k) v[i][k] = flip ? t[(3 - i) * DIM + k] : t[i * DIM + k];
}

// One edge (p -> q) of the subject polygon that crosses the clip line, with the signed distances
// (dp, dq) of its ends.  point(): where it crosses.
//   planar : p + t (q - p), t = dp / (dp - dq)                       (Sutherland-Hodgman)
//   sphere : the same chord point pushed back to the unit sphere.  (dp - dq) (p + t (q - p)) =
//            dp q - dq p, so the direction needs no division: normalise sign(dp - dq) (dp q - dq p).
//            (Differs from "divide, interpolate, normalise" by rounding only, ~1e-16 relative.)
constexpr double CLIP_SNAP = 4e-15;   // relative position along the edge below which a crossing is an end vertex

template <int DIM>
struct Crossing {
    double px = 1.0, py = 0.0, pz = 0.0, qx = 0.0, qy = 1.0, qz = 0.0, dp = 1.0, dq = -1.0;
    __device__ __forceinline__ void point(double &rx, double &ry, double &rz) const {
        if (DIM == 3) {
            rx = fma(dp, qx, -(dq * px));
            ry = fma(dp, qy, -(dq * py));
            rz = fma(dp, qz, -(dq * pz));
            double inv = rsqrt(fma(rx, rx, fma(ry, ry, rz * rz)));
            if (dp < dq) inv = -inv;
            rx *= inv; ry *= inv; rz *= inv;
        } else {
            const double t = dp / (dp - dq);
            rx = fma(t, qx - px, px); ry = fma(t, qy - py, py); rz = 0.0;
        }
    }
};

// The working polygon is SYMBOLIC: a 32-bit word of 4-bit vertex ids (up to 8 vertices) into a
// per-thread table of 12 points in shared memory -- ids 0..3 are the subject cell's corners, every
// clip pass appends at most two crossing points.  A pass computes the signed distances of the
// current vertices (static unrolled reads), and
//   * all inside  -> nothing to do (no copy),          * all outside -> empty,
//   * one inside run (the convex case) -> the new polygon is that run followed by the exit and the
//     enter crossing: a rotate/mask of the id word, and BOTH crossing points are computed in one
//     straight-line block, so the lanes of a warp do the expensive arithmetic together,
//   * several runs (round-off on degenerate input) -> sequential Sutherland-Hodgman over the ids.
// A crossing within CLIP_SNAP of an end of its edge reuses that vertex's id (exact coordinates).
constexpr int QUAD_SLOTS = 12;

template <int DIM, int NT>
struct PointTable {
    double *b;   // this thread's column: b[(id * DIM + c) * NT]
    __device__ __forceinline__ double get(int id, int c) const { return b[(id * DIM + c) * NT]; }
    __device__ __forceinline__ void set(int id, int c, double x) { b[(id * DIM + c) * NT] = x; }
};

// Which vertex id stands for the crossing of edge p -> q (distances dp, dq of different sign)?
// -1: none (the crossing is a vertex that is part of the inside run anyway), idp / idq: that end
// vertex itself (kept as the boundary point), -2: a new point.
__device__ __forceinline__ int crossing_kind(double dp, double dq, bool p_inside, int idp, int idq) {
    const double adp = fabs(dp), adq = fabs(dq);
    if (adq <= CLIP_SNAP * adp) return p_inside ? idq : -1;     // at q
    if (adp <= CLIP_SNAP * adq) return p_inside ? -1 : idp;     // at p
    return -2;
}

// Pre-pass over the ORIGINAL corners (static code, every lane busy): a clip edge that leaves all
// four subject corners inside can never cut (the working polygon only shrinks); one that leaves all
// four outside empties the intersection (most false candidates end here).
// Returns -1 (empty) or the 4-bit set of clip edges that cut.
// clip_nrm (sphere only, may be null): the clip grid's edge-plane normals, 12 doubles per cell, orientation folded in
// (bp_bounds_kernel writes them with the same edge_normal, so the distances are the same bits either way).
template <int DIM, bool WIDE = false>
__device__ __forceinline__ int quad_prepass(const CellsView &gs, int64_t s, const CellsView &gc, int64_t c,
                                            const double *__restrict__ clip_nrm = nullptr) {
    // Cells stored clockwise are NOT reversed: a clockwise clip cell has its interior on the other side
    // of every edge (the signed distances change sign); a clockwise subject only changes the sign of the
    // final area.
    double sv[4][DIM], cv[4][DIM];
    const bool have_n = DIM == 3 && clip_nrm;
    if constexpr (WIDE && DIM == 3) {
        load_quad_wide(gs.verts + s * 12, sv);
        load_quad_wide(have_n ? clip_nrm + c * 12 : gc.verts + c * 12, cv);
    } else {
        load_quad<DIM>(gs.verts + s * 4 * DIM, false, sv);
        load_quad<DIM>(have_n ? clip_nrm + c * 4 * DIM : gc.verts + c * 4 * DIM, false, cv);
    }
    const bool cflip = !have_n && gc.flip && gc.flip[c];
    uint32_t cut = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double *u = cv[e], *v = cv[(e + 1) & 3];
        double nx, ny, nz = 0.0, h0 = 0.0;
        if (DIM == 3) {
            if (have_n) { nx = u[0]; ny = u[1]; nz = u[2]; }
            else edge_normal(u, v, nx, ny, nz);
        } else {
            nx = -(v[1] - u[1]); ny = v[0] - u[0];
            h0 = -(nx * u[0] + ny * u[1]);
        }
        uint32_t me = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double d = DIM == 3 ? fma(nx, sv[i][0], fma(ny, sv[i][1], nz * sv[i][2]))
                                : fma(nx, sv[i][0], fma(ny, sv[i][1], h0));
            if (cflip) d = -d;
            if (d >= 0.0) me |= 1u << i;
        }
        if (me == 0u) return -1;
        if (me != 15u) cut |= 1u << e;       // (a zero-length edge has n = 0: every d = 0, never cuts)
    }
    return (int)cut;
}

// Area of subject ∩ clip given the set of cutting clip edges (quad_prepass).  A lane only visits ITS
// cutting edges, so the lanes of a warp meet in the cut code even when different edges cut them.
template <int DIM, int NT, bool WIDE = false>
__device__ double quad_cut_area(const CellsView &gs, int64_t s, const CellsView &gc, int64_t c, uint32_t cut,
                                double *smem /* QUAD_SLOTS * DIM * NT doubles */, const double *__restrict__ clip_nrm = nullptr) {
    PointTable<DIM, NT> tab{smem + threadIdx.x};
    const bool have_n = DIM == 3 && clip_nrm;
    const double *cbase = have_n ? clip_nrm + c * 4 * DIM : gc.verts + c * 4 * DIM;
    const double sc = (!have_n && gc.flip && gc.flip[c]) ? -1.0 : 1.0;        // clockwise clip cell: normals negated
    const double ss = (gs.flip && gs.flip[s]) ? -1.0 : 1.0;        // clockwise subject: area negated
    {
        double sv[4][DIM];
        if constexpr (WIDE && DIM == 3) load_quad_wide(gs.verts + s * 12, sv);
        else load_quad<DIM>(gs.verts + s * 4 * DIM, false, sv);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < DIM; ++k) tab.set(i, k, sv[i][k]);
    }
    uint32_t poly = 0x3210u;     // vertex i = nibble i
    int m = 4, nv = 4;
    while (cut) {
        const int e = __ffs(cut) - 1;
        cut &= cut - 1u;
        const int iu = e, iv = (e + 1) & 3;
        double u[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) { u[k] = __ldg(cbase + iu * DIM + k); if (!have_n) v[k] = __ldg(cbase + iv * DIM + k); }
        double nx, ny, nz = 0.0, h0 = 0.0;
        if (DIM == 3) {
            if (have_n) { nx = u[0]; ny = u[1]; nz = u[2]; }
            else {
                edge_normal(u, v, nx, ny, nz);
                nx *= sc; ny *= sc; nz *= sc;
            }
        } else {
            nx = -sc * (v[1] - u[1]); ny = sc * (v[0] - u[0]);
            h0 = -(nx * u[0] + ny * u[1]);
        }
        auto dist = [&](int id) -> double {
            if (DIM == 3) return fma(nx, tab.get(id, 0), fma(ny, tab.get(id, 1), nz * tab.get(id, 2)));
            return fma(nx, tab.get(id, 0), fma(ny, tab.get(id, 1), h0));
        };
        uint32_t mask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < m && dist((poly >> (4 * i)) & 15u) >= 0.0) mask |= 1u << i;
        const uint32_t full = (1u << m) - 1u;
        if (mask == full) continue;
        if (mask == 0u) return 0.0;
        const uint32_t next_in = ((mask >> 1) | (mask << (m - 1))) & full;    // bit i: vertex i+1 is inside
        const uint32_t exits = mask & ~next_in;                                 // i inside, i+1 outside
        const uint32_t enters = ~mask & next_in & full;                         // i outside, i+1 inside
        auto nib = [&](int i) -> int { return (int)((poly >> (4 * i)) & 15u); };
        if (__popc(exits) == 1) {
            const int i_out = __ffs(exits) - 1, i_in = __ffs(enters) - 1;
            int k = __popc(mask);                                               // 1 <= k <= m - 1 <= 7
            const int rot = i_in + 1 == m ? 0 : i_in + 1;                       // first vertex of the inside run
            uint32_t np = rot ? ((poly >> (4 * rot)) | (poly << (4 * (m - rot)))) : poly;
            np &= (1u << (4 * k)) - 1u;
            const int p0 = nib(i_out), q0 = nib(i_out + 1 == m ? 0 : i_out + 1);   // exit edge (p inside)
            const int p1 = nib(i_in), q1 = rot ? nib(rot) : nib(0);               // enter edge (q inside)
            Crossing<DIM> X0, X1;
            X0.px = tab.get(p0, 0); X0.py = tab.get(p0, 1); X0.pz = DIM == 3 ? tab.get(p0, 2) : 0.0;
            X0.qx = tab.get(q0, 0); X0.qy = tab.get(q0, 1); X0.qz = DIM == 3 ? tab.get(q0, 2) : 0.0;
            X1.px = tab.get(p1, 0); X1.py = tab.get(p1, 1); X1.pz = DIM == 3 ? tab.get(p1, 2) : 0.0;
            X1.qx = tab.get(q1, 0); X1.qy = tab.get(q1, 1); X1.qz = DIM == 3 ? tab.get(q1, 2) : 0.0;
            X0.dp = DIM == 3 ? fma(nx, X0.px, fma(ny, X0.py, nz * X0.pz)) : fma(nx, X0.px, fma(ny, X0.py, h0));
            X0.dq = DIM == 3 ? fma(nx, X0.qx, fma(ny, X0.qy, nz * X0.qz)) : fma(nx, X0.qx, fma(ny, X0.qy, h0));
            X1.dp = DIM == 3 ? fma(nx, X1.px, fma(ny, X1.py, nz * X1.pz)) : fma(nx, X1.px, fma(ny, X1.py, h0));
            X1.dq = DIM == 3 ? fma(nx, X1.qx, fma(ny, X1.qy, nz * X1.qz)) : fma(nx, X1.qx, fma(ny, X1.qy, h0));
            double r0x, r0y, r0z, r1x, r1y, r1z;
            X0.point(r0x, r0y, r0z);
            X1.point(r1x, r1y, r1z);
            int id0 = crossing_kind(X0.dp, X0.dq, true, p0, q0);
            int id1 = crossing_kind(X1.dp, X1.dq, false, p1, q1);
            if (id0 == -2) {
                id0 = nv++;
                tab.set(id0, 0, r0x); tab.set(id0, 1, r0y);
                if (DIM == 3) tab.set(id0, 2, r0z);
            }
            if (id1 == -2) {
                id1 = nv++;
                tab.set(id1, 0, r1x); tab.set(id1, 1, r1y);
                if (DIM == 3) tab.set(id1, 2, r1z);
            }
            if (id0 >= 0) { np |= (uint32_t)id0 << (4 * k); ++k; }
            if (id1 >= 0 && id1 != id0) { np |= (uint32_t)id1 << (4 * k); ++k; }
            poly = np;
            m = k;
        } else {
            // several inside runs: sequential pass over the ids
            uint32_t np = 0;
            int mo = 0;
            for (int i = 0; i < m; ++i) {
                const int j = i + 1 == m ? 0 : i + 1;
                const bool in_p = (mask >> i) & 1u, in_q = (mask >> j) & 1u;
                const int idp = nib(i), idq = nib(j);
                if (in_p != in_q) {
                    Crossing<DIM> X;
                    X.px = tab.get(idp, 0); X.py = tab.get(idp, 1); X.pz = DIM == 3 ? tab.get(idp, 2) : 0.0;
                    X.qx = tab.get(idq, 0); X.qy = tab.get(idq, 1); X.qz = DIM == 3 ? tab.get(idq, 2) : 0.0;
                    X.dp = dist(idp); X.dq = dist(idq);
                    int id = crossing_kind(X.dp, X.dq, in_p, idp, idq);
                    if (id == -2 && nv < QUAD_SLOTS) {
                        double rx, ry, rz;
                        X.point(rx, ry, rz);
                        id = nv++;
                        tab.set(id, 0, rx); tab.set(id, 1, ry);
                        if (DIM == 3) tab.set(id, 2, rz);
                    }
                    if (id >= 0 && mo < 8) { np |= (uint32_t)id << (4 * mo); ++mo; }
                }
                if (in_q && mo < 8) { np |= (uint32_t)idq << (4 * mo); ++mo; }
            }
            poly = np;
            m = mo;
        }
        if (m < 3) return 0.0;
    }
    if (m < 3) return 0.0;
    const int ia = (int)(poly & 15u);
    if (DIM == 3) {
        ExcessAcc acc;
        const d3 a = {tab.get(ia, 0), tab.get(ia, 1), tab.get(ia, 2)};
        int ib = (int)((poly >> 4) & 15u);
        d3 b = {tab.get(ib, 0), tab.get(ib, 1), tab.get(ib, 2)};
        for (int i = 2; i < m; ++i) {
            const int ic = (int)((poly >> (4 * i)) & 15u);
            const d3 cc = {tab.get(ic, 0), tab.get(ic, 1), tab.get(ic, 2)};
            acc.add_triangle(a, b, cc);
            b = cc;
        }
        return ss * acc.area();
    } else {
        double sarea = 0.0;
        const double x0 = tab.get(ia, 0), y0 = tab.get(ia, 1);
        const int ib = (int)((poly >> 4) & 15u);
        double ax = tab.get(ib, 0) - x0, ay = tab.get(ib, 1) - y0;
        for (int i = 2; i < m; ++i) {
            const int ic = (int)((poly >> (4 * i)) & 15u);
            const double bx = tab.get(ic, 0) - x0, by = tab.get(ic, 1) - y0;
            sarea += ax * by - ay * bx;
            ax = bx; ay = by;
        }
        return 0.5 * ss * sarea;
    }
}

}  // namespace crg

// clipfast.cuh -- K3 fast path: area of (spherical quadrilateral) ∩ (spherical quadrilateral) without building the
// intersection polygon.
//
// Replaces, per candidate pair, DefaultIntersectionOperator{Spherical}
// (/root/reference/src/regridder/regridder.jl:96-103: ConvexConvexSutherlandHodgman clip, then GO.area) for the
// pairs in which the clip cell C cuts the subject cell S with ONE edge or with TWO ADJACENT edges -- 91 % of the
// surviving pairs of BASELINE config 5; everything else (three or four cutting edges, two opposite ones) goes to the
// symbolic Sutherland-Hodgman of geom.cuh through a job list.
//
// The area of a spherical polygon is the sum over its boundary arcs a -> b of the signed excess T(P; a, b) of the
// triangle (P, a, b), for ANY fixed point P.  The boundary of S ∩ H1 ∩ H2 (H = the half-spaces of the cutting clip
// edges; every other clip edge leaves all of S inside) consists of
//   (A) the parts of S's four edges that lie inside H1 ∩ H2 -- per edge ONE parametric interval [t_in, t_out] of the
//       chord p + t (q - p), because the signed distance to a great-circle plane is linear along the chord
//       (Liang-Barsky), and
//   (B) parts of the two cutting great circles.
// With P = the point where the planes of the two cutting edges meet -- the clip cell's corner as the COMPUTED planes
// define it (kernels.cuh: quad_normals_kernel); any point of the plane when only one edge cuts -- P lies on both
// cutting circles, every (B) triangle is degenerate, and
//       area(S ∩ C) = sum over the 4 edges of S of T(P; a_i, b_i).
// Straight-line code: no working polygon, no vertex table, no dependence between the four edges, and no cascade of
// classifications of computed points -- a subject edge that lies ON a cutting circle (nested / identical grids)
// contributes ~0 whichever way its round-off distances fall, so no vertex snapping is needed.
// Crossing points are the reference clip's: the chord point where the signed distance vanishes, pushed back to the
// sphere, in the division-free form normalise(sign(dp - dq) (dp q - dq p)).
//
// The arithmetic is written once for the device and the host (CRG_HD): tests/test_clipfast_host.py compiles it with
// g++ and checks it pair by pair against the scalar Sutherland-Hodgman restatement of the reference, on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CRG_HD __host__ __device__ __forceinline__
#else
#define CRG_HD inline
#endif

namespace crg {

CRG_HD double cf_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// Great-circle normals of the four edges of a quadrilateral, products rounded separately (geom.cuh: edge_normal --
// a zero-length edge gives an exact zero normal, which never cuts), negated for a cell stored clockwise so that
// "inside" is n . x >= 0 for every cell.
CRG_HD void cf_quad_normals(const double (&v)[4][3], bool flip, double (&n)[4][3]) {
    const double sg = flip ? -1.0 : 1.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int e = 0; e < 4; ++e) {
        const double *u = v[e], *w = v[(e + 1) & 3];
#if defined(__CUDA_ARCH__)
        n[e][0] = sg * __dsub_rn(__dmul_rn(u[1], w[2]), __dmul_rn(u[2], w[1]));
        n[e][1] = sg * __dsub_rn(__dmul_rn(u[2], w[0]), __dmul_rn(u[0], w[2]));
        n[e][2] = sg * __dsub_rn(__dmul_rn(u[0], w[1]), __dmul_rn(u[1], w[0]));
#else
        volatile double a0 = u[1] * w[2], a1 = u[2] * w[1], b0 = u[2] * w[0], b1 = u[0] * w[2], c0 = u[0] * w[1], c1 = u[1] * w[0];
        n[e][0] = sg * (a0 - a1); n[e][1] = sg * (b0 - b1); n[e][2] = sg * (c0 - c1);
#endif
    }
}

// The four corners of a cell AS ITS COMPUTED EDGE PLANES DEFINE THEM: corner j = normalise(n_{j-1} x n_j), on the side of
// the stored vertex.  The cross product of two nearly parallel unit vectors carries an absolute error of ~1e-16, so the
// computed great circle misses its own end points by ~1e-16 / (edge length); a polygon clipped against the computed
// planes (Sutherland-Hodgman) has its corner where the planes meet, and the wedge sum below must use the same point P --
// with the stored vertex instead the two differ by ~1e-16 * (polygon size / edge length): measured 2e-17 on BASELINE
// config 5, 5e-12 of the largest entry.  A zero-length edge (pole) or a straight angle: the vertex moved onto the plane of
// the longer edge.
CRG_HD void cf_quad_corners(const double (&v)[4][3], const double (&n)[4][3], double (&p)[4][3]) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) {
        const double *a = n[(j + 3) & 3], *b = n[j];
        const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
        const double cc = c0 * c0 + c1 * c1 + c2 * c2;
        const double na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], nb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        if (cc > 1e-24 * na * nb && cc > 0.0) {
            double inv = cf_rsqrt(cc);
            if (c0 * v[j][0] + c1 * v[j][1] + c2 * v[j][2] < 0.0) inv = -inv;
            p[j][0] = c0 * inv; p[j][1] = c1 * inv; p[j][2] = c2 * inv;
        } else {
            const double *m = na >= nb ? a : b;
            const double mm = na >= nb ? na : nb;
            const double t = mm > 0.0 ? (m[0] * v[j][0] + m[1] * v[j][1] + m[2] * v[j][2]) / mm : 0.0;
            p[j][0] = v[j][0] - t * m[0]; p[j][1] = v[j][1] - t * m[1]; p[j][2] = v[j][2] - t * m[2];
        }
    }
}

enum { CF_EMPTY = -1, CF_INSIDE = 0, CF_FAST = 1, CF_GENERAL = 2 };

// Signed distances of the subject's corners to the clip cell's edge planes and what follows from their signs:
//   CF_EMPTY   some clip edge has all four corners outside: no intersection (most false candidates end here);
//   CF_INSIDE  no clip edge cuts: the intersection is the subject cell;
//   CF_FAST    one edge or two adjacent edges cut: e1 = the (first) cutting edge, two = the next edge cuts as well;
//   CF_GENERAL anything else.
// cut = the set of cutting edges (bit e).
CRG_HD double cf_distance(const double *n, const double *x) { return fma(n[0], x[0], fma(n[1], x[1], n[2] * x[2])); }

CRG_HD int cf_classify(const double (&s)[4][3], const double (&n)[4][3], int &e1, bool &two, uint32_t &cut) {
    cut = 0;
    bool empty = false;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int e = 0; e < 4; ++e) {
        uint32_t me = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i)
            if (cf_distance(n[e], s[i]) >= 0.0) me |= 1u << i;
        empty = empty || me == 0u;
        if (me != 15u) cut |= 1u << e;       // (a zero normal: every distance is 0, never cuts)
    }
    e1 = 0; two = false;
    if (empty) return CF_EMPTY;
    if (cut == 0u) return CF_INSIDE;
    // single edges 1 2 4 8, adjacent pairs 3 6 12 9
    const bool single = (cut & (cut - 1u)) == 0u;
    const bool adjacent = cut == 3u || cut == 6u || cut == 12u || cut == 9u;
    if (!single && !adjacent) return CF_GENERAL;
    two = adjacent;
    e1 = cut == 9u ? 3 : (cut & 1u) ? 0 : (cut & 2u) ? 1 : (cut & 4u) ? 2 : 3;
    return CF_FAST;
}

// sum of T(P; a_i, b_i) over the four subject edges (see the banner); P = the clip cell's corner (e1 + 1) & 3.
// Returns the SIGNED area (negative for a subject stored clockwise).
CRG_HD double cf_wedge_area(const double (&s)[4][3], const double (&d1)[4], const double (&d2)[4], const double (&P)[3]) {
    double re = 1.0, im = 0.0, total = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        const double *p = s[i], *q = s[j];
        const double dp1 = d1[i], dq1 = d1[j], dp2 = d2[i], dq2 = d2[j];
        const bool ip1 = dp1 >= 0.0, iq1 = dq1 >= 0.0, ip2 = dp2 >= 0.0, iq2 = dq2 >= 0.0;
        const bool killed = (!ip1 && !iq1) || (!ip2 && !iq2);
        const bool en1 = !ip1 && iq1, en2 = !ip2 && iq2, ex1 = ip1 && !iq1, ex2 = ip2 && !iq2;
        // t_j = dp_j / (dp_j - dq_j); comparisons by cross-multiplication (denominators of two enters are both
        // negative, of two exits both positive, of an enter and an exit of opposite sign)
        const double D1 = dp1 - dq1, D2 = dp2 - dq2;
        const double m12 = dp1 * D2, m21 = dp2 * D1;
        const bool a2 = en2 && (!en1 || m21 > m12);         // both enter: the later one bounds the interval
        const bool b2 = ex2 && (!ex1 || m21 < m12);         // both exit: the earlier one
        const bool a_any = en1 || en2, b_any = ex1 || ex2;
        const double dpa = a2 ? dp2 : dp1, dqa = a2 ? dq2 : dq1;
        const double dpb = b2 ? dp2 : dp1, dqb = b2 ? dq2 : dq1;
        // an enter and an exit on the same edge (from different planes): the interval must not be empty
        const bool ok = dpa * (dpb - dqb) > dpb * (dpa - dqa);
        bool valid = !killed && (!(a_any && b_any) || ok);
        // the ends of the interval
        double a[3], b[3];
        {
            double r0 = fma(dpa, q[0], -(dqa * p[0])), r1 = fma(dpa, q[1], -(dqa * p[1])), r2 = fma(dpa, q[2], -(dqa * p[2]));
            const double rr = fma(r0, r0, fma(r1, r1, r2 * r2));
            double inv = cf_rsqrt(rr);
            if (dpa < dqa) inv = -inv;
            const bool use = a_any && rr > 1e-280;
            a[0] = use ? r0 * inv : p[0]; a[1] = use ? r1 * inv : p[1]; a[2] = use ? r2 * inv : p[2];
        }
        {
            double r0 = fma(dpb, q[0], -(dqb * p[0])), r1 = fma(dpb, q[1], -(dqb * p[1])), r2 = fma(dpb, q[2], -(dqb * p[2]));
            const double rr = fma(r0, r0, fma(r1, r1, r2 * r2));
            double inv = cf_rsqrt(rr);
            if (dpb < dqb) inv = -inv;
            const bool use = b_any && rr > 1e-280;
            b[0] = use ? r0 * inv : q[0]; b[1] = use ? r1 * inv : q[1]; b[2] = use ? r2 * inv : q[2];
        }
        // T(P; a, b): tan(E / 2) = P . ((a - P) x (b - P)) / (1 + P.a + a.b + b.P), accumulated as a complex product
        const double wa0 = a[0] - P[0], wa1 = a[1] - P[1], wa2 = a[2] - P[2];
        const double wb0 = b[0] - P[0], wb1 = b[1] - P[1], wb2 = b[2] - P[2];
        const double c0 = wa1 * wb2 - wa2 * wb1, c1 = wa2 * wb0 - wa0 * wb2, c2 = wa0 * wb1 - wa1 * wb0;
        double det = fma(P[0], c0, fma(P[1], c1, P[2] * c2));
        double den = 1.0 + fma(P[0], a[0], fma(P[1], a[1], P[2] * a[2])) + fma(a[0], b[0], fma(a[1], b[1], a[2] * b[2])) +
                     fma(b[0], P[0], fma(b[1], P[1], b[2] * P[2]));
        if (valid && !(den > 0.0)) {           // a triangle wider than a hemisphere (toy grids): its own atan2
            total += atan2(det, den);
            valid = false;
        }
        det = valid ? det : 0.0;               // (x 1 + 0 i: exactly nothing, so an empty intersection is exactly 0)
        den = valid ? den : 1.0;
        const double nre = re * den - im * det, nim = re * det + im * den;
        re = nre; im = nim;
    }
    // the accumulated half-excess is a small angle for grid cells: series, else the library routine
    double ang;
    if (re > 0.0 && fabs(im) <= 0.03125 * re) {
        const double x = im / re, t = x * x;
        double pl = fma(t, 1.0 / 13.0, -1.0 / 11.0);
        pl = fma(pl, t, 1.0 / 9.0);
        pl = fma(pl, t, -1.0 / 7.0);
        pl = fma(pl, t, 1.0 / 5.0);
        pl = fma(pl, t, -1.0 / 3.0);
        ang = fma(pl * t, x, x);
    } else {
        ang = atan2(im, re);
    }
    return 2.0 * (total + ang);
}

// One pair on the host (tests): *kind = CF_*; the area is meaningful for CF_FAST and CF_INSIDE (the subject's own area
// through the same sum with no cutting plane).  sflip: the subject is stored clockwise.  corners: the clip cell's
// corners where its computed edge planes meet (kernels.cuh: quad_normals_kernel).
CRG_HD double cf_pair_area(const double (&s)[4][3], bool sflip, const double (&n)[4][3], const double *corners, int *kind) {
    int e1; bool two;
    uint32_t cut;
    *kind = cf_classify(s, n, e1, two, cut);
    if (*kind == CF_EMPTY || *kind == CF_GENERAL) return 0.0;
    const int e2 = (e1 + 1) & 3;
    double d1[4], d2[4];
    for (int i = 0; i < 4; ++i) {
        d1[i] = *kind == CF_FAST ? cf_distance(n[e1], s[i]) : 1.0;
        d2[i] = two ? cf_distance(n[e2], s[i]) : 1.0;
    }
    const double P[3] = {corners[3 * e2], corners[3 * e2 + 1], corners[3 * e2 + 2]};
    const double a = cf_wedge_area(s, d1, d2, P);
    return sflip ? -a : a;
}

}  // namespace crg

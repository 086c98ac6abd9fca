/*
 * crg_oracle.c -- CPU ORACLE for the Regridder-build and regrid! hot paths.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (conservativeregridding.jl_b200/csrc) never calls into this file and has no CPU
 * fallback.
 *
 * It restates, in plain C (Float64, scalar, one pair at a time), the algorithm of
 * JuliaGeo/ConservativeRegridding.jl v0.2.5 (citations relative to /root/reference):
 *
 *   src/regridder/intersection_areas.jl:4-32    per-pair loop, keep area > 0
 *   src/regridder/intersection_areas.jl:67-122  candidates -> areas -> sparse(dst, src, area)
 *   src/regridder/regridder.jl:54-62            normalize!  (divide by maximum(A))
 *   src/regridder/regridder.jl:96-103           spherical operator: convex-convex
 *                                               Sutherland-Hodgman clip, then polygon area
 *   src/regridder/regridder.jl:87-94            planar operator (clip + planar area)
 *   src/regridder/regridder.jl:165-178          per-cell geometric areas
 *   src/regridder/regrid.jl:95-118              mul! then divide by dst_areas
 *   src/utils/MultithreadedDualDepthFirstSearch.jl:12-65, src/trees/grids.jl:245-287,
 *   src/trees/quadtree_cursors.jl:214-316       dual-tree candidate search over bounding
 *                                               caps (the CPU baseline's broad phase)
 *
 * PARITY STATUS: the per-pair arithmetic (GO.intersection / GO.area) lives in the
 * third-party package GeometryOps.jl (compat "0.1.33", no Manifest pinned), whose source
 * is not under /root/reference.  The clip and area below restate its *published
 * algorithm* (Sutherland-Hodgman against great-circle half-spaces; spherical excess);
 * they are pinned against the reference's own planar known-answer tests
 * (test/usecases/simple.jl:10-51, README.md:52-80, test/regridding.jl) and its
 * spherical invariants (row/col sums == geometric areas at rtol sqrt(eps),
 * test/sweat.jl:113-116; sum of areas = 4 pi R^2), plus 50-digit mpmath areas on
 * sampled pairs (oracle/highprec.py).  Per-entry spherical values of the Julia package
 * itself: PARITY UNPINNED (no golden matrices exist in the reference; Julia is not
 * installable here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXV 64 /* max vertices of a working polygon */

/* ------------------------------------------------------------------------- */
/* small vector helpers                                                       */
/* ------------------------------------------------------------------------- */
static inline double dot3(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* Signed area of the great-circle triangle (a, b, c) on the unit sphere, CCW seen
 * from outside positive.  tan(E/2) = a.(b x c) / (1 + a.b + b.c + c.a); the triple
 * product is taken in difference form a.((b-a) x (c-a)) (identical in exact
 * arithmetic) so that small triangles keep full relative accuracy. */
static double sph_triangle_area(const double *a, const double *b, const double *c) {
    double ba[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ca[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double x[3];
    cross3(ba, ca, x);
    double det = dot3(a, x);
    double den = 1.0 + dot3(a, b) + dot3(b, c) + dot3(c, a);
    return 2.0 * atan2(det, den);
}

/* Signed spherical-excess area of a polygon (unit sphere): fan from vertex 0.
 * GO.area(Spherical, poly) / radius^2  (regridder.jl:102,173-177). */
double orc_sph_polygon_area(const double *p, int n) {
    double s = 0.0;
    for (int i = 1; i + 1 < n; ++i) s += sph_triangle_area(p, p + 3 * i, p + 3 * (i + 1));
    return s;
}

/* Signed planar polygon area (shoelace about vertex 0). GO.area(Planar, poly). */
double orc_planar_polygon_area(const double *p, int n) {
    double s = 0.0;
    for (int i = 1; i + 1 < n; ++i) {
        double ax = p[2 * i] - p[0], ay = p[2 * i + 1] - p[1];
        double bx = p[2 * (i + 1)] - p[0], by = p[2 * (i + 1) + 1] - p[1];
        s += ax * by - ay * bx;
    }
    return 0.5 * s;
}

/* ------------------------------------------------------------------------- */
/* Sutherland-Hodgman clipping                                                */
/* ------------------------------------------------------------------------- */

/* Clip spherical polygon `subj` (ns vertices, unit vectors) against the convex CCW
 * polygon `clip` (nc vertices): successive half-space cuts by the great circles
 * through the clip edges; inside <=> (u x v).p >= 0; crossing = chord interpolation
 * projected back to the sphere (exact for great-circle arcs shorter than pi).
 * Writes the result to `out` (capacity ORC_MAXV) and returns its vertex count. */
int orc_sph_clip(const double *subj, int ns, const double *clip, int nc, double *out) {
    double bufA[ORC_MAXV * 3], bufB[ORC_MAXV * 3], d[ORC_MAXV];
    double *cur = bufA, *nxt = bufB;
    int m = ns;
    memcpy(cur, subj, sizeof(double) * 3 * ns);
    for (int k = 0; k < nc && m > 0; ++k) {
        const double *u = clip + 3 * k, *v = clip + 3 * ((k + 1) % nc);
        double n[3];
        cross3(u, v, n);
        if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) continue; /* zero-length edge */
        for (int i = 0; i < m; ++i) d[i] = dot3(n, cur + 3 * i);
        int mo = 0;
        for (int i = 0; i < m; ++i) {
            int j = (i + 1 == m) ? 0 : i + 1;
            const double *p = cur + 3 * i, *q = cur + 3 * j;
            double dp = d[i], dq = d[j];
            int in_p = dp >= 0.0, in_q = dq >= 0.0;
            if (in_p != in_q) {
                double t = dp / (dp - dq);
                double r[3] = {p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1]),
                               p[2] + t * (q[2] - p[2])};
                double inv = 1.0 / sqrt(dot3(r, r));
                if (mo < ORC_MAXV) {
                    nxt[3 * mo] = r[0] * inv; nxt[3 * mo + 1] = r[1] * inv; nxt[3 * mo + 2] = r[2] * inv;
                    ++mo;
                }
            }
            if (in_q && mo < ORC_MAXV) {
                nxt[3 * mo] = q[0]; nxt[3 * mo + 1] = q[1]; nxt[3 * mo + 2] = q[2];
                ++mo;
            }
        }
        double *t2 = cur; cur = nxt; nxt = t2;
        m = mo;
    }
    if (m < 3) return 0;
    memcpy(out, cur, sizeof(double) * 3 * m);
    return m;
}

int orc_planar_clip(const double *subj, int ns, const double *clip, int nc, double *out) {
    double bufA[ORC_MAXV * 2], bufB[ORC_MAXV * 2], d[ORC_MAXV];
    double *cur = bufA, *nxt = bufB;
    int m = ns;
    memcpy(cur, subj, sizeof(double) * 2 * ns);
    for (int k = 0; k < nc && m > 0; ++k) {
        const double *u = clip + 2 * k, *v = clip + 2 * ((k + 1) % nc);
        double ex = v[0] - u[0], ey = v[1] - u[1];
        if (ex == 0.0 && ey == 0.0) continue;
        for (int i = 0; i < m; ++i)
            d[i] = ex * (cur[2 * i + 1] - u[1]) - ey * (cur[2 * i] - u[0]);
        int mo = 0;
        for (int i = 0; i < m; ++i) {
            int j = (i + 1 == m) ? 0 : i + 1;
            const double *p = cur + 2 * i, *q = cur + 2 * j;
            double dp = d[i], dq = d[j];
            int in_p = dp >= 0.0, in_q = dq >= 0.0;
            if (in_p != in_q) {
                double t = dp / (dp - dq);
                if (mo < ORC_MAXV) {
                    nxt[2 * mo] = p[0] + t * (q[0] - p[0]);
                    nxt[2 * mo + 1] = p[1] + t * (q[1] - p[1]);
                    ++mo;
                }
            }
            if (in_q && mo < ORC_MAXV) {
                nxt[2 * mo] = q[0]; nxt[2 * mo + 1] = q[1];
                ++mo;
            }
        }
        double *t2 = cur; cur = nxt; nxt = t2;
        m = mo;
    }
    if (m < 3) return 0;
    memcpy(out, cur, sizeof(double) * 2 * m);
    return m;
}

static void reverse_ring(double *p, int n, int dim) {
    for (int i = 0, j = n - 1; i < j; ++i, --j)
        for (int c = 0; c < dim; ++c) {
            double t = p[dim * i + c]; p[dim * i + c] = p[dim * j + c]; p[dim * j + c] = t;
        }
}

/* DefaultIntersectionOperator (regridder.jl:87-103): area of the intersection of two
 * convex polygons; unit sphere / plane; always >= 0 up to round-off sign noise.
 * Orientation-robust: both rings are normalised to CCW first. */
double orc_intersection_area(int manifold, const double *p1, int n1, const double *p2, int n2) {
    int dim = manifold ? 3 : 2;
    double a[ORC_MAXV * 3], b[ORC_MAXV * 3], out[ORC_MAXV * 3];
    if (n1 > ORC_MAXV / 2 || n2 > ORC_MAXV / 2) return NAN;
    memcpy(a, p1, sizeof(double) * dim * n1);
    memcpy(b, p2, sizeof(double) * dim * n2);
    double sa = manifold ? orc_sph_polygon_area(a, n1) : orc_planar_polygon_area(a, n1);
    double sb = manifold ? orc_sph_polygon_area(b, n2) : orc_planar_polygon_area(b, n2);
    if (sa < 0) reverse_ring(a, n1, dim);
    if (sb < 0) reverse_ring(b, n2, dim);
    int m = manifold ? orc_sph_clip(a, n1, b, n2, out) : orc_planar_clip(a, n1, b, n2, out);
    if (m < 3) return 0.0;
    return manifold ? orc_sph_polygon_area(out, m) : orc_planar_polygon_area(out, m);
}

/* ------------------------------------------------------------------------- */
/* cells                                                                      */
/* ------------------------------------------------------------------------- */
typedef struct {
    const double *verts; /* fixed: [ncells][nv][dim]; ragged: [total][dim] */
    const int32_t *off;  /* NULL => fixed nv */
    int64_t ncells;
    int nv, dim;
} orc_grid;

static inline int grid_cell(const orc_grid *g, int64_t i, const double **p) {
    if (g->off) { *p = g->verts + (int64_t)g->dim * g->off[i]; return g->off[i + 1] - g->off[i]; }
    *p = g->verts + (int64_t)g->dim * g->nv * i;
    return g->nv;
}

/* areas(manifold, tree) (regridder.jl:165-178): |GO.area(cell)| on the unit sphere. */
void orc_cell_areas(int manifold, const double *verts, const int32_t *off, int64_t ncells, int nv,
                    double *areas) {
    orc_grid g = {verts, off, ncells, nv, manifold ? 3 : 2};
    for (int64_t i = 0; i < ncells; ++i) {
        const double *p; int n = grid_cell(&g, i, &p);
        areas[i] = fabs(manifold ? orc_sph_polygon_area(p, n) : orc_planar_polygon_area(p, n));
    }
}

/* compute_intersection_areas (intersection_areas.jl:4-32): loop the candidate pairs
 * (src i1, dst i2), keep area > 0.  0-based indices.  Output arrays must hold npairs.
 * nthreads > 1 mirrors the reference's chunked StableTasks.@spawn (:86-113): chunk
 * results are concatenated in chunk order, so the output order equals the serial one. */
int64_t orc_compute_intersection_areas(int manifold,
                                       const double *dverts, const int32_t *doff, int64_t ndst, int dnv,
                                       const double *sverts, const int32_t *soff, int64_t nsrc, int snv,
                                       const int64_t *pair_src, const int64_t *pair_dst, int64_t npairs,
                                       int nthreads,
                                       int64_t *out_src, int64_t *out_dst, double *out_area) {
    orc_grid gd = {dverts, doff, ndst, dnv, manifold ? 3 : 2};
    orc_grid gs = {sverts, soff, nsrc, snv, manifold ? 3 : 2};
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(npairs > 0 ? npairs : 1));
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
#endif
    for (int64_t k = 0; k < npairs; ++k) {
        const double *p1, *p2;
        int n1 = grid_cell(&gs, pair_src[k], &p1);
        int n2 = grid_cell(&gd, pair_dst[k], &p2);
        tmp[k] = orc_intersection_area(manifold, p1, n1, p2, n2);
    }
    int64_t m = 0;
    for (int64_t k = 0; k < npairs; ++k)
        if (tmp[k] > 0.0) { out_src[m] = pair_src[k]; out_dst[m] = pair_dst[k]; out_area[m] = tmp[k]; ++m; }
    free(tmp);
    return m;
}

/* ------------------------------------------------------------------------- */
/* SparseArrays.sparse(I, J, V, m, n)  (intersection_areas.jl:115-121)        */
/* COO -> CSC, duplicates summed, row indices sorted within each column.      */
/* 0-based.  colptr has n+1 entries; rowval/nzval capacity nnz.               */
/* ------------------------------------------------------------------------- */
typedef struct { int64_t r; double v; int64_t seq; } orc_ent;
static int ent_cmp(const void *a, const void *b) {
    const orc_ent *x = (const orc_ent *)a, *y = (const orc_ent *)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
int64_t orc_coo_to_csc(int64_t nrows, int64_t ncols, int64_t nnz, const int64_t *rows,
                       const int64_t *cols, const double *vals, int64_t *colptr, int64_t *rowval,
                       double *nzval) {
    (void)nrows;
    int64_t *cnt = (int64_t *)calloc((size_t)ncols + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nnz; ++k) cnt[cols[k] + 1]++;
    for (int64_t c = 0; c < ncols; ++c) cnt[c + 1] += cnt[c];
    orc_ent *e = (orc_ent *)malloc(sizeof(orc_ent) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncols + 1));
    memcpy(cur, cnt, sizeof(int64_t) * (size_t)(ncols + 1));
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t pos = cur[cols[k]]++;
        e[pos].r = rows[k]; e[pos].v = vals[k]; e[pos].seq = k;
    }
    int64_t out = 0;
    colptr[0] = 0;
    for (int64_t c = 0; c < ncols; ++c) {
        int64_t lo = cnt[c], hi = cnt[c + 1];
        qsort(e + lo, (size_t)(hi - lo), sizeof(orc_ent), ent_cmp);
        for (int64_t k = lo; k < hi; ++k) {
            if (k > lo && e[k].r == e[k - 1].r) nzval[out - 1] += e[k].v;
            else { rowval[out] = e[k].r; nzval[out] = e[k].v; ++out; }
        }
        colptr[c + 1] = out;
    }
    free(e); free(cur); free(cnt);
    return out;
}

/* normalize!(R) (regridder.jl:54-62). */
void orc_normalize(int64_t nnz, double *nzval, int64_t ndst, double *dst_areas, int64_t nsrc,
                   double *src_areas) {
    if (nnz <= 0) return;
    double m = nzval[0];
    for (int64_t k = 1; k < nnz; ++k) if (nzval[k] > m) m = nzval[k];
    for (int64_t k = 0; k < nnz; ++k) nzval[k] /= m;
    for (int64_t i = 0; i < ndst; ++i) dst_areas[i] /= m;
    for (int64_t i = 0; i < nsrc; ++i) src_areas[i] /= m;
}

/* perform_regridding! (regrid.jl:95-98): y = A x with A in CSC -- the stdlib's serial
 * column-scatter kernel; and y = A' x (row gather) for transpose(R).  Then
 * finalize_regridding! (regrid.jl:104-118): y ./= areas when `divide`. */
void orc_csc_mul(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                 const double *nzval, const double *x, double *y, const double *areas, int divide) {
    for (int64_t i = 0; i < nrows; ++i) y[i] = 0.0;
    for (int64_t c = 0; c < ncols; ++c) {
        double xc = x[c];
        for (int64_t k = colptr[c]; k < colptr[c + 1]; ++k) y[rowval[k]] += nzval[k] * xc;
    }
    if (divide) for (int64_t i = 0; i < nrows; ++i) y[i] /= areas[i];
}
void orc_csc_tmul(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                  const double *nzval, const double *x, double *y, const double *areas, int divide) {
    (void)nrows;
    for (int64_t c = 0; c < ncols; ++c) {
        double s = 0.0;
        for (int64_t k = colptr[c]; k < colptr[c + 1]; ++k) s += nzval[k] * x[rowval[k]];
        y[c] = divide ? s / areas[c] : s;
    }
}

/* ------------------------------------------------------------------------- */
/* Dual-tree candidate search over bounding caps (the reference's broad phase) */
/* ------------------------------------------------------------------------- */
/* Explicit tree: node k has children [child_lo[k], child_hi[k]) (node ids) or, for a
 * leaf, cells leaf_cells[leaf_lo[k] .. leaf_hi[k]) with their own extents.
 * Extent = spherical cap (cx,cy,cz,radius) or planar box (xmin,xmax,ymin,ymax). */
typedef struct {
    int64_t nnodes;
    const int64_t *child_lo, *child_hi; /* child_hi == child_lo => leaf */
    const int64_t *leaf_lo, *leaf_hi;
    const int64_t *leaf_cells;          /* field-linear cell indices */
    const double *node_ext;             /* [nnodes][4] */
    const double *cell_ext;             /* [n leaf_cells entries][4], aligned with leaf_cells */
    int manifold;
} orc_tree;

static inline int ext_intersects(int manifold, const double *a, const double *b) {
    if (manifold) { /* GO.UnitSpherical._intersects(cap, cap): centre distance <= r1 + r2 */
        double cr[3];
        cross3(a, b, cr);
        double dist = atan2(sqrt(dot3(cr, cr)), dot3(a, b));
        return dist <= a[3] + b[3];
    }
    /* Extents.intersects */
    return a[0] <= b[1] && b[0] <= a[1] && a[2] <= b[3] && b[2] <= a[3];
}

typedef struct { int64_t *src, *dst; int64_t n, cap; } orc_pairvec;
static void pv_push(orc_pairvec *v, int64_t s, int64_t d) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->src = (int64_t *)realloc(v->src, sizeof(int64_t) * (size_t)v->cap);
        v->dst = (int64_t *)realloc(v->dst, sizeof(int64_t) * (size_t)v->cap);
    }
    v->src[v->n] = s; v->dst[v->n] = d; v->n++;
}

/* STI.dual_depth_first_search(f, pred, n1, n2): 4-way case split on leaf/non-leaf. */
static void dual_dfs(const orc_tree *t1, int64_t n1, const orc_tree *t2, int64_t n2, orc_pairvec *out) {
    int leaf1 = t1->child_hi[n1] == t1->child_lo[n1];
    int leaf2 = t2->child_hi[n2] == t2->child_lo[n2];
    int mf = t1->manifold;
    if (leaf1 && leaf2) {
        for (int64_t a = t1->leaf_lo[n1]; a < t1->leaf_hi[n1]; ++a)
            for (int64_t b = t2->leaf_lo[n2]; b < t2->leaf_hi[n2]; ++b)
                if (ext_intersects(mf, t1->cell_ext + 4 * a, t2->cell_ext + 4 * b))
                    pv_push(out, t1->leaf_cells[a], t2->leaf_cells[b]);
    } else if (leaf1) {
        for (int64_t c2 = t2->child_lo[n2]; c2 < t2->child_hi[n2]; ++c2)
            if (ext_intersects(mf, t1->node_ext + 4 * n1, t2->node_ext + 4 * c2)) dual_dfs(t1, n1, t2, c2, out);
    } else if (leaf2) {
        for (int64_t c1 = t1->child_lo[n1]; c1 < t1->child_hi[n1]; ++c1)
            if (ext_intersects(mf, t1->node_ext + 4 * c1, t2->node_ext + 4 * n2)) dual_dfs(t1, c1, t2, n2, out);
    } else {
        for (int64_t c1 = t1->child_lo[n1]; c1 < t1->child_hi[n1]; ++c1)
            for (int64_t c2 = t2->child_lo[n2]; c2 < t2->child_hi[n2]; ++c2)
                if (ext_intersects(mf, t1->node_ext + 4 * c1, t2->node_ext + 4 * c2))
                    dual_dfs(t1, c1, t2, c2, out);
    }
}

/* multithreaded_dual_query (MultithreadedDualDepthFirstSearch.jl:12-65): descend both
 * trees together until `spawn_depth` levels down (the reference uses a should_parallelize
 * size policy), then run one serial dual DFS per surviving (subtree, subtree) task.
 * Returns malloc'ed arrays via out_src/out_dst (free with orc_free). */
typedef struct { int64_t a, b; } orc_task;
static void collect_tasks(const orc_tree *t1, int64_t n1, const orc_tree *t2, int64_t n2, int depth,
                          orc_task **tasks, int64_t *nt, int64_t *cap) {
    int leaf1 = t1->child_hi[n1] == t1->child_lo[n1];
    int leaf2 = t2->child_hi[n2] == t2->child_lo[n2];
    if (depth == 0 || leaf1 || leaf2) {
        if (*nt == *cap) { *cap = *cap ? *cap * 2 : 256; *tasks = (orc_task *)realloc(*tasks, sizeof(orc_task) * (size_t)*cap); }
        (*tasks)[*nt].a = n1; (*tasks)[*nt].b = n2; (*nt)++;
        return;
    }
    for (int64_t c1 = t1->child_lo[n1]; c1 < t1->child_hi[n1]; ++c1)
        for (int64_t c2 = t2->child_lo[n2]; c2 < t2->child_hi[n2]; ++c2)
            if (ext_intersects(t1->manifold, t1->node_ext + 4 * c1, t2->node_ext + 4 * c2))
                collect_tasks(t1, c1, t2, c2, depth - 1, tasks, nt, cap);
}

int64_t orc_dual_query(int manifold, int nthreads, int spawn_depth,
                       int64_t nn1, const int64_t *clo1, const int64_t *chi1, const int64_t *llo1,
                       const int64_t *lhi1, const int64_t *lc1, const double *next1, const double *cext1,
                       int64_t nn2, const int64_t *clo2, const int64_t *chi2, const int64_t *llo2,
                       const int64_t *lhi2, const int64_t *lc2, const double *next2, const double *cext2,
                       int64_t **out_src, int64_t **out_dst) {
    orc_tree t1 = {nn1, clo1, chi1, llo1, lhi1, lc1, next1, cext1, manifold};
    orc_tree t2 = {nn2, clo2, chi2, llo2, lhi2, lc2, next2, cext2, manifold};
    orc_task *tasks = NULL; int64_t nt = 0, cap = 0;
    if (ext_intersects(manifold, next1, next2))
        collect_tasks(&t1, 0, &t2, 0, spawn_depth, &tasks, &nt, &cap);
    orc_pairvec *res = (orc_pairvec *)calloc((size_t)(nt > 0 ? nt : 1), sizeof(orc_pairvec));
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int64_t k = 0; k < nt; ++k) dual_dfs(&t1, tasks[k].a, &t2, tasks[k].b, &res[k]);
    int64_t total = 0;
    for (int64_t k = 0; k < nt; ++k) total += res[k].n;
    int64_t *s = (int64_t *)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    int64_t *d = (int64_t *)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    int64_t o = 0;
    for (int64_t k = 0; k < nt; ++k) { /* reduce(vcat, map(fetch, tasks)) */
        if (res[k].n) {
            memcpy(s + o, res[k].src, sizeof(int64_t) * (size_t)res[k].n);
            memcpy(d + o, res[k].dst, sizeof(int64_t) * (size_t)res[k].n);
            o += res[k].n;
        }
        free(res[k].src); free(res[k].dst);
    }
    free(res); free(tasks);
    *out_src = s; *out_dst = d;
    return total;
}

void orc_free(void *p) { free(p); }

/* Bounding cap of one cell, restating _spherical_cap (src/trees/grids.jl:256-273) for a
 * single cell: centre = normalised vertex mean, radius = 1.0001 * max distance over the
 * vertices and the slerp mid-points of the edges. */
static double sph_dist(const double *a, const double *b) {
    double cr[3];
    cross3(a, b, cr);
    return atan2(sqrt(dot3(cr, cr)), dot3(a, b));
}
void orc_cell_caps(const double *verts, const int32_t *off, int64_t ncells, int nv, double *caps) {
    orc_grid g = {verts, off, ncells, nv, 3};
    for (int64_t i = 0; i < ncells; ++i) {
        const double *p; int n = grid_cell(&g, i, &p);
        double c[3] = {0, 0, 0};
        for (int k = 0; k < n; ++k) { c[0] += p[3 * k]; c[1] += p[3 * k + 1]; c[2] += p[3 * k + 2]; }
        double inv = 1.0 / sqrt(dot3(c, c));
        c[0] *= inv; c[1] *= inv; c[2] *= inv;
        double r = 0.0;
        for (int k = 0; k < n; ++k) {
            const double *a = p + 3 * k, *b = p + 3 * ((k + 1) % n);
            double d = sph_dist(c, a);
            if (d > r) r = d;
            double m[3] = {a[0] + b[0], a[1] + b[1], a[2] + b[2]};
            double mm = dot3(m, m);
            if (mm > 0) { /* slerp(a, b, 0.5) */
                double im = 1.0 / sqrt(mm);
                m[0] *= im; m[1] *= im; m[2] *= im;
                d = sph_dist(c, m);
                if (d > r) r = d;
            }
        }
        caps[4 * i] = c[0]; caps[4 * i + 1] = c[1]; caps[4 * i + 2] = c[2]; caps[4 * i + 3] = r * 1.0001;
    }
}

/* cell_range_extent for a spherical CellBasedGrid (src/trees/grids.jl:245-287): cap centred
 * on the normalised mean of the 4 corners of the index rectangle, radius = 1.0001 * max
 * distance over the corners, the slerp mid-points of the 4 sides and every perimeter
 * vertex.  P is the (nx+1) x (ny+1) vertex matrix, i-major; ranges = [i0, i1, j0, j1)
 * half-open cell ranges.  (The reference recomputes this on every node visit; the port
 * computes each node once.) */
void orc_structured_node_caps(const double *P, int64_t nx, int64_t ny, const int64_t *ranges,
                              int64_t nn, double *caps, int nthreads) {
    (void)nx;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
#endif
    for (int64_t k = 0; k < nn; ++k) {
        int64_t imin = ranges[4 * k], imax = ranges[4 * k + 1], jmin = ranges[4 * k + 2], jmax = ranges[4 * k + 3];
#define PV(i, j) (P + 3 * ((i) * (ny + 1) + (j)))
        const double *c4[4] = {PV(imin, jmin), PV(imax, jmin), PV(imax, jmax), PV(imin, jmax)};
        double c[3] = {0, 0, 0};
        for (int q = 0; q < 4; ++q) { c[0] += c4[q][0]; c[1] += c4[q][1]; c[2] += c4[q][2]; }
        double nrm = sqrt(dot3(c, c));
        double *o = caps + 4 * k;
        if (!(nrm > 0.0)) { o[0] = 0; o[1] = 0; o[2] = 1; o[3] = 3.2; continue; }
        c[0] /= nrm; c[1] /= nrm; c[2] /= nrm;
        double r = 0.0;
        for (int q = 0; q < 4; ++q) {
            const double *a = c4[q], *b = c4[(q + 1) & 3];
            double d = sph_dist(c, a);
            if (d > r) r = d;
            double m[3] = {a[0] + b[0], a[1] + b[1], a[2] + b[2]};
            double mm = dot3(m, m);
            if (mm > 0) {
                double im = 1.0 / sqrt(mm);
                m[0] *= im; m[1] *= im; m[2] *= im;
                d = sph_dist(c, m);
                if (d > r) r = d;
            }
        }
        for (int64_t j = jmin; j <= jmax; ++j) {
            double d = sph_dist(c, PV(imin, j)); if (d > r) r = d;
            d = sph_dist(c, PV(imax, j)); if (d > r) r = d;
        }
        for (int64_t i = imin + 1; i < imax; ++i) {
            double d = sph_dist(c, PV(i, jmin)); if (d > r) r = d;
            d = sph_dist(c, PV(i, jmax)); if (d > r) r = d;
        }
#undef PV
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = r * 1.0001;
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline asks for the cores it may use. */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

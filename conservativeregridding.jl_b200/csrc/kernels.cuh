// kernels.cuh -- clip/area kernels, COO -> CSR/CSC assembly helpers, normalize, and the CSR SpMM
// apply kernels (area division fused in).  The SpMV lives in sell.cuh, the cell areas in
// broadphase.cuh (fused into the bounds pass).
#pragma once
#include "common.cuh"
#include "geom.cuh"
#include "clipfast.cuh"
#include "broadphase.cuh"

namespace crg {

// =======================================================================================
// K3: clip + area of every candidate pair.  Replaces compute_intersection_areas
// (/root/reference/src/regridder/intersection_areas.jl:4-32).  FP64-pipe / issue bound.
//
// Neither kernel has CTA-wide synchronisation: every pair's area goes to a dense array (0 when the
// pair does not survive `area > threshold`) and the survivor counts are added to the counter of the
// pair's 1024-pair tile.  A streaming pass (compact_pairs_kernel, after a scan of the tile counters)
// then compacts the survivors IN ORDER.  Because the candidate list is grouped by destination cell
// in increasing order, the compacted COO triples are grouped by row, which is what the assembly
// builds on (crg_b200.cu: assemble).  (A single-pass chained-scan compaction inside the clip kernel
// was measured 40 % slower: tiles with long clips hold back the retirement of their successors.)
// =======================================================================================
constexpr int CLIP_TILE = 1024;      // pairs per compaction tile

// General cells (ragged rings, triangles, up to CRG_MAX_VERTS vertices): one thread per pair,
// Sutherland-Hodgman ping-ponged through two shared-memory polygons (geom.cuh: clip_pair_area).
template <int DIM, int NT, int MAXW>
__global__ void __launch_bounds__(NT) clip_kernel(CellsView gd, CellsView gs, const int2 *__restrict__ pairs,
                                                  int64_t npairs, double thresh, double *__restrict__ area_out,
                                                  uint32_t *__restrict__ tile_count) {
    extern __shared__ double clip_smem[];
    const int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x;
    double area = 0.0;
    if (idx < npairs) {
        const int2 pr = pairs[idx];
        area = clip_pair_area<DIM, NT, MAXW>(gs, pr.x, gd, pr.y, clip_smem);
        if (!(area > thresh) || !(area > 0.0)) area = 0.0;     // `area > 0` (intersection_areas.jl:24); NaN drops too
        area_out[idx] = area;
    }
    const unsigned mask = __ballot_sync(CRG_FULL, area != 0.0);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(&tile_count[idx / CLIP_TILE], (uint32_t)__popc(mask));
}

// Quadrilateral x quadrilateral (every structured grid of BASELINE.json), geom.cuh's symbolic clip.
// A warp owns CLIP_CHUNK consecutive candidate pairs and alternates between two stages, without
// block barriers:
//   1. 32 pairs at a time, every lane runs the static pre-pass of one pair: empty (area 0, written at
//      once), untouched (the source cell's own area, when K0's unit-sphere areas are at hand) or the
//      set of cutting edges; the surviving pairs are appended to one of the warp's two queues in
//      shared memory, by number of cutting edges (<= 1, >= 2) -- 16-bit entries, so that the queues
//      cost no more shared memory than one (a third queue cost a CTA per SM);
//   2. whenever a queue holds 32 jobs, every lane takes one and runs the cuts + the area -- full warps
//      (a third of the candidates are false, a quarter of the rest is not cut at all) whose lanes loop
//      over the same number of cuts.  What is left at the end of the chunk is run in mixed batches.
// Every pair's area is written exactly once; the non-zero count goes to the tile counter at the end.
constexpr int CLIP_CHUNK = 512;     // divides CLIP_TILE
constexpr int CLIP_QUEUES = 2;
// clip_nrm: the destination grid's edge-plane normals from the bounds pass (sphere; null: computed per pair).  WIDE: 256-bit
// vertex loads (every cell record 32-byte aligned).
template <int DIM, int NT, bool WIDE = false>
__global__ void __launch_bounds__(NT) clip_quad_kernel(CellsView gd, CellsView gs, const int2 *__restrict__ pairs,
                                                       int64_t npairs, double thresh,
                                                       const double *__restrict__ unit_src_areas,
                                                       double *__restrict__ area_out, uint32_t *__restrict__ tile_count,
                                                       const double *__restrict__ clip_nrm = nullptr) {
    extern __shared__ double clip_smem[];
    __shared__ uint16_t s_queue[NT / 32][CLIP_QUEUES][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t base = ((int64_t)blockIdx.x * (NT / 32) + wid) * CLIP_CHUNK;
    if (base >= npairs) return;
    uint16_t (*queue)[64] = s_queue[wid];
    const unsigned lt_mask = (1u << lane) - 1u;
    int q0 = 0, q1 = 0, nz = 0;                           // queue lengths (warp-uniform)
#pragma unroll 1
    for (int round = 0; round <= CLIP_CHUNK / 32; ++round) {        // the extra round only drains the queues
        const bool last = round == CLIP_CHUNK / 32;
        const int64_t idx = base + round * 32 + lane;
        int state = -1;
        if (!last && idx < npairs) {
            const int2 pr = pairs[idx];
            state = quad_prepass<DIM, WIDE>(gs, pr.x, gd, pr.y, clip_nrm);
            double area = 0.0;
            if (state == 0 && unit_src_areas) { area = unit_src_areas[pr.x]; state = -1; }
            if (state < 0) {
                if (!(area > thresh) || !(area > 0.0)) area = 0.0;
                area_out[idx] = area;
                nz += area != 0.0;
            }
        }
        const int ncut = __popc((unsigned)max(state, 0));
        const int cls = state < 0 ? -1 : (ncut <= 1 ? 0 : 1);
        const uint16_t job = (uint16_t)((round * 32 + lane) | (max(state, 0) << 9));      // 9 bits pair, 4 bits edges
        const unsigned b0 = __ballot_sync(CRG_FULL, cls == 0), b1 = __ballot_sync(CRG_FULL, cls == 1);
        if (cls == 0) queue[0][q0 + __popc(b0 & lt_mask)] = job;
        if (cls == 1) queue[1][q1 + __popc(b1 & lt_mask)] = job;
        q0 += __popc(b0); q1 += __popc(b1);
        __syncwarp();
#pragma unroll 1
        for (;;) {
            uint32_t j = 0;
            bool have = false;
            if (q0 >= 32) { q0 -= 32; j = queue[0][q0 + lane]; have = true; }
            else if (q1 >= 32) { q1 -= 32; j = queue[1][q1 + lane]; have = true; }
            else if (last && q0 + q1 > 0) {                // leftovers: one mixed batch, queue 0 first
                const int t0 = min(q0, 32), t1 = min(q1, 32 - t0);
                if (lane < t0) { j = queue[0][q0 - t0 + lane]; have = true; }
                else if (lane < t0 + t1) { j = queue[1][q1 - t1 + (lane - t0)]; have = true; }
                q0 -= t0; q1 -= t1;
            } else {
                break;
            }
            if (have) {
                const int64_t jdx = base + (j & 511u);
                const int2 pr = pairs[jdx];
                double area = quad_cut_area<DIM, NT, WIDE>(gs, pr.x, gd, pr.y, j >> 9, clip_smem, clip_nrm);
                if (!(area > thresh) || !(area > 0.0)) area = 0.0;     // `area > 0` (intersection_areas.jl:24); NaN drops too
                area_out[jdx] = area;
                nz += area != 0.0;
            }
            __syncwarp();
        }
    }
    nz = warp_sum(nz);
    if (lane == 0 && nz) atomicAdd(&tile_count[base / CLIP_TILE], (uint32_t)nz);
}

// ---- spherical quadrilaterals: wedge sums (clipfast.cuh) for most pairs, symbolic Sutherland-Hodgman for the rest ------
// Per cell of the CLIP grid, once: the edge-plane normals (exact products, orientation folded in) and the four corners
// AS THE COMPUTED PLANES DEFINE THEM, corner j = normalise(n_{j-1} x n_j).  The cross product of two nearly parallel
// unit vectors carries an absolute error of ~1e-16, i.e. the computed great circle misses its own end points by
// ~1e-16 / (edge length); a polygon clipped against the computed planes (Sutherland-Hodgman) has its corner where the
// planes meet, and the wedge sum of clipfast.cuh must use the same point P -- with the stored vertex instead, the two
// differ by ~1e-16 * (polygon size / edge length), measured 2e-17 on config 5 (4e-12 of the largest entry).
__global__ void __launch_bounds__(256) quad_normals_kernel(CellsView g, double *__restrict__ nrm, double *__restrict__ corners) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    double v[4][3], n[4][3], p[4][3];
    load_quad<3>(g.verts + c * 12, false, v);
    cf_quad_normals(v, g.flip && g.flip[c], n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double *a = n[(j + 3) & 3], *b = n[j];
        const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
        const double cc = c0 * c0 + c1 * c1 + c2 * c2;
        const double na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], nb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        if (cc > 1e-24 * na * nb && cc > 0.0) {              // the two planes meet at a usable angle
            double inv = rsqrt(cc);
            if (c0 * v[j][0] + c1 * v[j][1] + c2 * v[j][2] < 0.0) inv = -inv;
            p[j][0] = c0 * inv; p[j][1] = c1 * inv; p[j][2] = c2 * inv;
        } else {                                              // a zero-length edge (pole) or a straight angle: the vertex,
            const double *m = na >= nb ? a : b;               // moved onto the plane of the longer edge
            const double mm = na >= nb ? na : nb;
            const double t = mm > 0.0 ? (m[0] * v[j][0] + m[1] * v[j][1] + m[2] * v[j][2]) / mm : 0.0;
            p[j][0] = v[j][0] - t * m[0]; p[j][1] = v[j][1] - t * m[1]; p[j][2] = v[j][2] - t * m[2];
        }
    }
    double2 *o = reinterpret_cast<double2 *>(nrm + c * 12), *q = reinterpret_cast<double2 *>(corners + c * 12);
    const double *f = &n[0][0], *h = &p[0][0];
#pragma unroll
    for (int i = 0; i < 6; ++i) { o[i] = make_double2(f[2 * i], f[2 * i + 1]); q[i] = make_double2(h[2 * i], h[2 * i + 1]); }
}

// A warp owns CF_CHUNK consecutive candidate pairs and alternates between three stages, without block barriers:
//   1. CLASSIFY, 32 pairs at a time, one lane per pair (cf_classify): empty -> area 0; inside -> the subject's own area;
//      fast (one edge or two adjacent edges cut) -> queue F; general (three or four cutting edges, two opposite ones)
//      -> queue G.  Queue entries carry the two cell ids, so the later stages do not go back to the pair list.
//   2. WEDGE, whenever queue F holds 32 jobs (and for what is left at the end of the chunk): every lane takes one job
//      and runs cf_wedge_area -- full warps of straight-line code, no shared-memory polygon, no divergence.
//   3. GENERAL, whenever queue G holds 32 jobs: geom.cuh's symbolic Sutherland-Hodgman (quad_cut_area) on the warp's
//      point table, which shares its shared memory with the record stage.
// The 96-byte records with SCATTERED cell ids (source-side cells in stage 1, subject cells in stage 2) are fetched by
// the whole warp -- six consecutive lanes read the six 16-byte pieces of one record, so a load instruction touches ~6
// lines instead of 32 -- and handed to their lanes through a padded shared-memory stage (112-byte stride:
// conflict-free 16-byte reads).  The destination side's records are the same for runs of consecutive pairs (the list
// is grouped by destination cell) and are read directly.
// The source cell is always the subject and the destination cell the clip cell, like the reference's
// intersection_operator(src_polygon, dst_polygon) (intersection_areas.jl:21-23): with the roles swapped the result
// differs by the round-off of the other clip order (measured up to 1.6e-16 on config 4, beyond the parity bar for
// slivers), however much cheaper the swap would make grids whose source cells are the larger ones.
#ifndef CF_MINB
#define CF_MINB 5
#endif
constexpr int CF_NT = 128;
constexpr int CF_REC = 112;
constexpr int CF_CHUNK = 1024;      // = CLIP_TILE: one tile counter per warp
static_assert(CF_CHUNK == CLIP_TILE, "one warp per compaction tile");
constexpr int CF_WARP_SMEM = QUAD_SLOTS * 3 * 32 * 8;      // the warp's point table (stage 3); the record stage lives in it
static_assert(CF_WARP_SMEM >= 32 * CF_REC, "the record stage fits into the point table");

// the warp fetches the 96-byte records rec[id * 12 ..] of its 32 lanes' ids into the stage; every lane then reads its own
__device__ __forceinline__ void cf_stage_records(const double *__restrict__ rec, int id, int lane, unsigned char *stage,
                                                 double (&out)[4][3]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int c = k * 32 + lane, j = c / 6, part = c - 6 * j;
        const int sj = __shfl_sync(CRG_FULL, id, j);
        const double2 v = __ldg(reinterpret_cast<const double2 *>(rec + (int64_t)sj * 12) + part);
        *reinterpret_cast<double2 *>(stage + j * CF_REC + part * 16) = v;
    }
    __syncwarp();
    double *f = &out[0][0];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double2 v = *reinterpret_cast<const double2 *>(stage + lane * CF_REC + k * 16);
        f[2 * k] = v.x; f[2 * k + 1] = v.y;
    }
    __syncwarp();
}

struct CfJob { int subj, clp; uint32_t bits; };      // bits: 10 pair-in-chunk | 4 cutting edges / (2 first edge, 1 two, 1 none)

__global__ void __launch_bounds__(CF_NT, CF_MINB)
clip_quad_fast_kernel(CellsView gd, CellsView gs, const double *__restrict__ clip_nrm, const double *__restrict__ clip_corners,
                      const int2 *__restrict__ pairs, int64_t npairs, double thresh,
                      const double *__restrict__ unit_subj_areas, double *__restrict__ area_out,
                      uint32_t *__restrict__ tile_count) {
    extern __shared__ __align__(16) unsigned char cf_smem[];
    __shared__ CfJob s_queue[CF_NT / 32][2][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t base = ((int64_t)blockIdx.x * (CF_NT / 32) + wid) * CF_CHUNK;
    if (base >= npairs) return;
    unsigned char *stage = cf_smem + wid * CF_WARP_SMEM;
    CfJob (*queue)[64] = s_queue[wid];
    const unsigned lt_mask = (1u << lane) - 1u;
    const double *__restrict__ subj_verts = gs.verts;
    const uint8_t *__restrict__ subj_flip = gs.flip;
    int qf = 0, qg = 0, nz = 0;                           // queue lengths (warp-uniform), surviving pairs of this lane
#pragma unroll 1
    for (int round = 0;; ++round) {
        const bool more = round < CF_CHUNK / 32 && base + round * 32 < npairs;      // (warp-uniform)
        if (more) {
            // ---- stage 1: classify 32 pairs ---------------------------------------------------------------------
            const int64_t idx = base + round * 32 + lane;
            const bool live = idx < npairs;
            const int2 pr = live ? pairs[idx] : make_int2(0, 0);
            double s[4][3], n[4][3];
            {
                const double2 *yp = reinterpret_cast<const double2 *>(clip_nrm + (int64_t)pr.y * 12);      // read directly
                double *yf = &n[0][0];
#pragma unroll
                for (int k = 0; k < 6; ++k) { const double2 v = __ldg(yp + k); yf[2 * k] = v.x; yf[2 * k + 1] = v.y; }
                cf_stage_records(subj_verts, pr.x, lane, stage, s);
            }
            int e1;
            bool two;
            uint32_t cut;
            int kind = cf_classify(s, n, e1, two, cut);
            if (!live) kind = CF_EMPTY;
            const int subj = pr.x, clp = pr.y;
            const bool inside_wedge = kind == CF_INSIDE && !unit_subj_areas;      // (radius != 1: no unit areas at hand)
            const bool pushf = kind == CF_FAST || inside_wedge, pushg = kind == CF_GENERAL;
            if (live && !pushf && !pushg) {
                double area = kind == CF_INSIDE ? unit_subj_areas[subj] : 0.0;
                if (!(area > thresh) || !(area > 0.0)) area = 0.0;     // `area > 0` (intersection_areas.jl:24); NaN drops too
                area_out[idx] = area;
                nz += area != 0.0;
            }
            const unsigned fm = __ballot_sync(CRG_FULL, pushf), gm = __ballot_sync(CRG_FULL, pushg);
            const uint32_t pic = (uint32_t)(round * 32 + lane);
            if (pushf) queue[0][qf + __popc(fm & lt_mask)] = CfJob{subj, clp, pic | ((uint32_t)e1 << 10) | ((two ? 1u : 0u) << 12) | ((inside_wedge ? 1u : 0u) << 13)};
            if (pushg) queue[1][qg + __popc(gm & lt_mask)] = CfJob{subj, clp, pic | (cut << 10)};
            qf += __popc(fm); qg += __popc(gm);
            __syncwarp();
        }
        // ---- stage 2: wedge sums, full warps -------------------------------------------------------------------------
#pragma unroll 1
        while (qf >= 32 || (!more && qf > 0)) {
            const int take = min(qf, 32);
            qf -= take;
            const bool have = lane < take;
            const CfJob job = queue[0][qf + (have ? lane : 0)];                  // (idle lanes shadow lane 0's job)
            const int e1 = (job.bits >> 10) & 3, e2 = (e1 + 1) & 3;
            const bool two = (job.bits >> 12) & 1u, none = (job.bits >> 13) & 1u;
            const double *np1 = clip_nrm + (int64_t)job.clp * 12 + 3 * e1, *np2 = clip_nrm + (int64_t)job.clp * 12 + 3 * e2;
            const double *pv = clip_corners + (int64_t)job.clp * 12 + 3 * e2;
            const double n1[3] = {__ldg(np1), __ldg(np1 + 1), __ldg(np1 + 2)};
            const double n2[3] = {__ldg(np2), __ldg(np2 + 1), __ldg(np2 + 2)};
            const double P[3] = {__ldg(pv), __ldg(pv + 1), __ldg(pv + 2)};
            const bool sflip = subj_flip && subj_flip[job.subj];
            double s[4][3];
            cf_stage_records(subj_verts, job.subj, lane, stage, s);
            double d1[4], d2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {      // the expressions of cf_classify: the same bits
                const double a = fma(n1[0], s[i][0], fma(n1[1], s[i][1], n1[2] * s[i][2]));
                const double b = fma(n2[0], s[i][0], fma(n2[1], s[i][1], n2[2] * s[i][2]));
                d1[i] = none ? 1.0 : a;
                d2[i] = two ? b : 1.0;
            }
            double area = cf_wedge_area(s, d1, d2, P);
            if (sflip) area = -area;
            if (!(area > thresh) || !(area > 0.0)) area = 0.0;
            if (have) { area_out[base + (job.bits & 1023u)] = area; nz += area != 0.0; }
        }
        // ---- stage 3: the general pairs, symbolic Sutherland-Hodgman on the warp's point table -----------------------------
#pragma unroll 1
        while (qg >= 32 || (!more && qg > 0)) {
            const int take = min(qg, 32);
            qg -= take;
            const bool have = lane < take;
            const CfJob job = queue[1][qg + (have ? lane : 0)];
            __syncwarp();
            double area = 0.0;
            if (have) {
                // (quad_cut_area addresses its table as smem[(id * 3 + c) * 32 + threadIdx.x]: hand it the warp's block)
                double *table = reinterpret_cast<double *>(stage) - (threadIdx.x - lane);
                area = quad_cut_area<3, 32>(gs, job.subj, gd, job.clp, (job.bits >> 10) & 15u, table);
                if (!(area > thresh) || !(area > 0.0)) area = 0.0;
                area_out[base + (job.bits & 1023u)] = area;
                nz += area != 0.0;
            }
            __syncwarp();
        }
        if (!more) break;
    }
    nz = warp_sum(nz);
    if (lane == 0 && nz) atomicAdd(&tile_count[base / CLIP_TILE], (uint32_t)nz);
}

// Stable compaction of the surviving pairs of one tile: tile_off = exclusive scan of tile_count.
__global__ void __launch_bounds__(256) compact_pairs_kernel(const int2 *__restrict__ pairs, const double *__restrict__ area,
                                                            int64_t npairs, const uint32_t *__restrict__ tile_off,
                                                            double scale, uint64_t *__restrict__ coo_key,
                                                            double *__restrict__ coo_val) {
    __shared__ uint32_t sm[33];
    const int64_t base = (int64_t)blockIdx.x * CLIP_TILE + (int64_t)threadIdx.x * (CLIP_TILE / 256);
    double a[CLIP_TILE / 256];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < CLIP_TILE / 256; ++k) {
        a[k] = base + k < npairs ? area[base + k] : 0.0;
        cnt += a[k] != 0.0;
    }
    uint32_t total;
    uint32_t pos = tile_off[blockIdx.x] + block_exclusive_scan(cnt, sm, &total);
#pragma unroll
    for (int k = 0; k < CLIP_TILE / 256; ++k)
        if (a[k] != 0.0) {
            const int2 pr = pairs[base + k];
            coo_key[pos] = ((uint64_t)(uint32_t)pr.y << 32) | (uint32_t)pr.x;
            coo_val[pos] = a[k] * scale;
            ++pos;
        }
}

// (K4, per-cell geometric area + orientation flags, is fused into bp_bounds_kernel -- broadphase.cuh)

// =======================================================================================
// K5 helpers: duplicate summation, key swap, CSR split.
// =======================================================================================
__global__ void __launch_bounds__(256) mark_heads_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                         uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// heads_pos = exclusive scan of the head flags.  Every head sums its run (segmented reduce).
__global__ void __launch_bounds__(256) dedupe_kernel(const uint64_t *__restrict__ keys, const double *__restrict__ vals,
                                                     const uint32_t *__restrict__ flags,
                                                     const uint32_t *__restrict__ heads_pos, int64_t n,
                                                     uint64_t *__restrict__ okeys, double *__restrict__ ovals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    double s = vals[i];
    for (int64_t j = i + 1; j < n && !flags[j]; ++j) s += vals[j];
    okeys[heads_pos[i]] = keys[i];
    ovals[heads_pos[i]] = s;
}

__global__ void __launch_bounds__(256) swap_key_kernel(const uint64_t *__restrict__ in, int64_t n,
                                                       uint64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const uint64_t k = in[i]; out[i] = (k << 32) | (k >> 32); }
}

// Sorted triples -> CSR.  keys = row << 32 | col; the triples are sorted by (row, col) [LOW = false: CSR of
// A] or by (col, row) [LOW = true: CSR of A^T, whose rows are the columns -- no swapped key copy is made].
//   csr_split_kernel: one thread per ENTRY writes the CSR column index and value; the first entry of a
//     row also writes the row pointer of its row and of up to ROWPTR_GAP empty rows just before it
//     (coalesced key reads, no search) into a rowptr array preset to -1;
//   rowptr_kernel: one thread per ROW -- rows still at -1 (inside longer runs of empty rows: half the
//     rows of a destination-sharded transpose block, or past the last entry) are found by binary search.
constexpr int ROWPTR_GAP = 4;
template <bool LOW>
__device__ __forceinline__ int64_t key_row(uint64_t k) { return LOW ? (int64_t)(uint32_t)k : (int64_t)(k >> 32); }

template <bool LOW, bool SPLIT>
__global__ void __launch_bounds__(256) csr_split_kernel(const uint64_t *__restrict__ keys, const double *__restrict__ vals,
                                                        int64_t nnz, int32_t *__restrict__ rowptr,
                                                        int32_t *__restrict__ colidx, double *__restrict__ ovals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint64_t k = keys[i];
    if (SPLIT) {
        colidx[i] = LOW ? (int32_t)(k >> 32) : (int32_t)(uint32_t)k;
        ovals[i] = vals[i];
    }
    const int64_t r = key_row<LOW>(k);
    const int64_t rp = i == 0 ? -1 : key_row<LOW>(keys[i - 1]);
    if (r == rp) return;
    const int64_t lo = max(rp + 1, r - ROWPTR_GAP);
    for (int64_t q = lo; q <= r; ++q) rowptr[q] = (int32_t)i;
}
template <bool LOW>
__global__ void __launch_bounds__(256) rowptr_kernel(const uint64_t *__restrict__ keys, int64_t nnz, int64_t n_rows,
                                                     int32_t *__restrict__ rowptr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (rowptr[r] >= 0) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (key_row<LOW>(keys[mid]) < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int32_t)lo;
}

// Row-grouped COO with short rows -> CSR: a block stages the entries of its ROWSORT_ROWS rows in
// shared memory (coalesced), one thread per row sorts its entries by column (insertion sort; the
// candidate lists are a few entries long), and the block writes the CSR column/value arrays, coalesced.
// A block whose rows hold more than ROWSORT_CAP entries sorts inside the CSR arrays (global memory).
constexpr int ROWSORT_ROWS = 128;
constexpr int ROWSORT_CAP = 2560;
__global__ void __launch_bounds__(ROWSORT_ROWS) row_sort_split_kernel(const uint64_t *__restrict__ keys,
                                                                      const double *__restrict__ vals,
                                                                      const int32_t *__restrict__ rowptr, int64_t n_rows,
                                                                      int32_t *__restrict__ colidx,
                                                                      double *__restrict__ out_vals) {
    // The triples are only READ (the column passes of the transpose run concurrently on them).
    __shared__ int32_t sc[ROWSORT_CAP];
    __shared__ double sv[ROWSORT_CAP];
    const int64_t r0 = (int64_t)blockIdx.x * ROWSORT_ROWS;
    const int64_t r = r0 + threadIdx.x;
    const int64_t r1 = min(r0 + (int64_t)ROWSORT_ROWS, n_rows);
    const int A = rowptr[r0], B = rowptr[r1];
    const bool staged = B - A <= ROWSORT_CAP;                                // block-uniform
    // the block's entries, unsorted, into shared memory (or, when they do not fit, into the CSR arrays)
    int32_t *c_ = staged ? sc - A : colidx;                                  // indexed by the global entry number
    double *v_ = staged ? sv - A : out_vals;
    for (int i = A + threadIdx.x; i < B; i += ROWSORT_ROWS) { c_[i] = (int32_t)(uint32_t)keys[i]; v_[i] = vals[i]; }
    __syncthreads();
    if (r < n_rows) {
        const int a = rowptr[r], b = rowptr[r + 1];
        for (int i = a + 1; i < b; ++i) {
            const int32_t k = c_[i];
            const double v = v_[i];
            int j = i - 1;
            while (j >= a) {
                const int32_t kj = c_[j];
                if (kj <= k) break;
                c_[j + 1] = kj;
                v_[j + 1] = v_[j];
                --j;
            }
            if (j + 1 != i) { c_[j + 1] = k; v_[j + 1] = v; }
        }
    }
    if (staged) {
        __syncthreads();
        for (int i = A + threadIdx.x; i < B; i += ROWSORT_ROWS) { colidx[i] = sc[i - A]; out_vals[i] = sv[i - A]; }
    }
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t *p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// (dst, src, area) int64/double triples -> packed keys
__global__ void __launch_bounds__(256) pack_coo_kernel(const int64_t *__restrict__ rows, const int64_t *__restrict__ cols,
                                                       int64_t n, uint64_t *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)(uint32_t)rows[i] << 32) | (uint32_t)cols[i];
}

// =======================================================================================
// K6: normalize!  (regridder.jl:54-62)
// =======================================================================================
__global__ void __launch_bounds__(256) max_kernel(const double *__restrict__ v, int64_t n, double *out) {
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, v[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_pos_double(out, m);
}
__global__ void __launch_bounds__(256) div_by_kernel(double *__restrict__ v, int64_t n, const double *__restrict__ m) {
    const double d = *m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = v[i] / d;
}

// =======================================================================================
// K8a: CSR SpMM, level-fastest layout x[cell * ldx + k]: one warp per row, lanes across the
// K levels (coalesced 8*K-byte reads of each gathered source row, L2-resident reuse).
// Replaces the NDSliceLoop of K sequential SpMVs (regrid.jl:303-318).
// =======================================================================================
template <int KT, bool DIVIDE>
__global__ void __launch_bounds__(256) spmm_lf_kernel(const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      double *__restrict__ y, const double *__restrict__ areas,
                                                      int64_t n_rows, int64_t K, int64_t ldx, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int a = rowptr[r], b = rowptr[r + 1];
    const double inv_area_num = DIVIDE ? areas[r] : 1.0;
    for (int64_t kb = 0; kb < K; kb += 32 * KT) {
        double acc[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) acc[t] = 0.0;
#pragma unroll 4
        for (int j = a; j < b; ++j) {
            const double v = vals[j];
            const double *xp = x + (int64_t)colidx[j] * ldx + kb;
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                const int64_t k = t * 32 + lane;
                if (kb + k < K) acc[t] += v * __ldg(&xp[k]);
            }
        }
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            const int64_t k = kb + t * 32 + lane;
            if (k < K) y[r * ldy + k] = DIVIDE ? acc[t] / inv_area_num : acc[t];
        }
    }
}

// =======================================================================================
// K8b: CSR SpMM, cell-fastest layout x[k * ldx + cell] (Julia dims = 1): one thread per row,
// KC levels per thread so the row's (col, val) are read once per KC levels; lanes hold
// neighbouring rows, so a gather instruction touches few 128-byte lines, and output writes
// are coalesced across rows.  (Measured alternatives on cfg3, K = 100: the same loop on the SELL
// copy 249 us, row entries held in registers for a whole level group 212 us, one entry per
// iteration 137 us, U = 2 entries per iteration 112 us, U = 4 109-121 us at 122 registers.  ncu: L1
// wavefronts bound it -- the j-th entries of 32 neighbouring rows lie on ~13 different lines.)
// =======================================================================================
template <int KC, bool DIVIDE, int U = 2>
__global__ void __launch_bounds__(128) spmm_cf_kernel(const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      double *__restrict__ y, const double *__restrict__ areas,
                                                      int64_t n_rows, int64_t K, int64_t ldx, int64_t ldy) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k0 = (int64_t)blockIdx.y * KC;
    if (r >= n_rows) return;
    const int a = rowptr[r], b = rowptr[r + 1];
    double acc[KC];
#pragma unroll
    for (int t = 0; t < KC; ++t) acc[t] = 0.0;
    const double ar = DIVIDE ? areas[r] : 1.0;
    // U entries per iteration, all U*KC gathers issued before the first FMA (memory-level
    // parallelism); entries past the row end are clamped to the last one with weight 0.
    for (int j = a; j < b; j += U) {
        double v[U];
        const double *xp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int jj = min(j + u, b - 1);
            v[u] = j + u < b ? vals[jj] : 0.0;
            xp[u] = x + colidx[jj] + k0 * ldx;
        }
        double g[U][KC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int t = 0; t < KC; ++t) g[u][t] = k0 + t < K ? __ldg(&xp[u][t * ldx]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int t = 0; t < KC; ++t) acc[t] += v[u] * g[u][t];
    }
#pragma unroll
    for (int t = 0; t < KC; ++t)
        if (k0 + t < K) y[(k0 + t) * ldy + r] = DIVIDE ? acc[t] / ar : acc[t];
}

// =======================================================================================
// FP64 FMA throughput micro-benchmark (roofline denominator of the clip kernel)
// =======================================================================================
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace crg

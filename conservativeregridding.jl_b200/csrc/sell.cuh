// sell.cuh -- K7: SpMV y = (A x) ./ areas on a SELL-32-sigma copy of the matrix.
//
// Replaces LinearAlgebra.mul! + the separate `dst ./= dst_areas` pass
// (/root/reference/src/regridder/regrid.jl:95-118).  HBM-bound: 12 B/nnz of matrix data dominate.
//
// Regridding matrices have short rows (3..15 entries) of locally similar length, which is the
// worst case for CSR on a GPU (every load either uncoalesced or behind a dependent row-pointer
// chain).  At assembly the rows are therefore re-laid out once in sliced ELLPACK:
//   * rows are sorted by length inside windows of SELL_SIGMA rows (stable, so equally long
//     neighbours stay neighbours) -> perm[], rlen[];
//   * 32 consecutive sorted rows form a slice, padded to its longest row; entry j of lane l is
//     stored at (slice_off + j) * 32 + l, so one warp instruction reads 32 consecutive values
//     (256 B) / column indices (128 B) -- perfectly coalesced, no shared memory, no barriers;
//   * slices taller than SELL_HP steps (polar rows) are cut into pieces of SELL_HP steps.
// The kernel runs one warp per piece, one lane per row: all loads of a step are independent of
// the previous step, lanes gather x for neighbouring rows at the same position (neighbouring
// source cells -> few 128 B lines per gather instruction).  Pieces of a cut slice leave their
// partial sums in scratch; the last piece to arrive (ticket) adds them in piece order, so the
// result is deterministic.  Padding measured on BASELINE cfg2: 0.8 % (forward), 1.9 % (transpose).
#pragma once
#include "common.cuh"

namespace crg {

constexpr int SELL_SIGMA = 1024;   // sorting window (rows)
constexpr int SELL_HP = 32;        // max steps per piece
constexpr int SELL_NB = 64;        // window sort: counting-sort bins (row lengths below this)
constexpr int SELL_UNR = 4;
#ifndef SELL_LD
#define SELL_LD __ldcs
#endif

struct SellView {
    const double *vals;        // padded, slice-major
    const int32_t *cols;
    const int32_t *perm;       // sorted position -> original row, -1 for padding rows
    const int32_t *rlen;       // sorted position -> row length
    const int32_t *slice_off;  // [nslices + 1], in steps
    const int4 *pieces;        // extra pieces (q >= 1) of cut slices: {slice, step_begin, partial_base, q | npieces << 16}
    const uint32_t *cut_base;  // [nslices] partial-slot base of a cut slice (read only when steps > SELL_HP)
    double *partial;           // [n_partial_slots][32]
    unsigned int *ticket;      // [n_partial_slots]
    int nslices;
    int npieces;               // number of extra pieces
};

// ---- build step 1: sort the rows of each window by decreasing length ----------------------------------
__global__ void __launch_bounds__(SELL_SIGMA) sell_sort_kernel(const int32_t *__restrict__ rowptr, int64_t n_rows,
                                                               int32_t *__restrict__ perm, int32_t *__restrict__ rlen,
                                                               int32_t *__restrict__ slice_steps) {
    __shared__ uint64_t key[SELL_SIGMA];
    const int t = threadIdx.x;
    const int64_t r = (int64_t)blockIdx.x * SELL_SIGMA + t;
    uint32_t len = 0;
    if (r < n_rows) len = (uint32_t)(rowptr[r + 1] - rowptr[r]);
    // Fast path (every row of the window shorter than SELL_NB): stable counting sort.  Per-warp
    // digit counts by __match_any_sync, one scan over (bin-major, warp-minor) counts -- five
    // barriers instead of the 55 of the bitonic network below.
    if (!__syncthreads_or(len >= (uint32_t)SELL_NB)) {
        uint32_t *h = reinterpret_cast<uint32_t *>(key);          // SELL_NB * 32 counters <= 8 KB
        __shared__ uint32_t scan_tmp[33];
        constexpr int PER = SELL_NB * 32 / SELL_SIGMA;
#pragma unroll
        for (int i = 0; i < PER; ++i) h[t * PER + i] = 0;
        __syncthreads();
        const int lane = t & 31, wid = t >> 5;
        const int bin = SELL_NB - 1 - (int)len;                   // descending length
        const unsigned same = __match_any_sync(CRG_FULL, bin);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (rank == 0) h[bin * 32 + wid] = (uint32_t)__popc(same);
        __syncthreads();
        uint32_t loc[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { loc[i] = sum; sum += h[t * PER + i]; }
        uint32_t total;
        const uint32_t base = block_exclusive_scan<uint32_t>(sum, scan_tmp, &total);
#pragma unroll
        for (int i = 0; i < PER; ++i) h[t * PER + i] = base + loc[i];
        __syncthreads();
        const int64_t pos = (int64_t)blockIdx.x * SELL_SIGMA + h[bin * 32 + wid] + rank;
        perm[pos] = r < n_rows ? (int32_t)r : -1;
        rlen[pos] = (int32_t)len;
        if ((pos & 31) == 0) slice_steps[pos >> 5] = (int32_t)len;
        return;
    }
    // descending by length, ascending by original position (stable); padding rows (len 0, beyond
    // n_rows) sort last because their position is largest
    key[t] = ((uint64_t)len << 32) | (uint32_t)(SELL_SIGMA - 1 - t);
    __syncthreads();
    for (int k = 2; k <= SELL_SIGMA; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int ixj = t ^ j;
            if (ixj > t) {
                const uint64_t a = key[t], b = key[ixj];
                const bool desc = (t & k) == 0;       // final order: descending
                if (desc ? (a < b) : (a > b)) { key[t] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
    const uint64_t kk = key[t];
    const int src = SELL_SIGMA - 1 - (int)(uint32_t)kk;
    const int64_t rr = (int64_t)blockIdx.x * SELL_SIGMA + src;
    const int64_t pos = (int64_t)blockIdx.x * SELL_SIGMA + t;
    perm[pos] = rr < n_rows ? (int32_t)rr : -1;
    rlen[pos] = (int32_t)(kk >> 32);
    if ((t & 31) == 0) slice_steps[pos >> 5] = (int32_t)(kk >> 32);   // first lane of a slice is its longest row
}

// ---- build step 2: pieces per slice (after exclusive scans of piece / partial-slot counts) --------------
__global__ void __launch_bounds__(256) sell_count_kernel(const int32_t *__restrict__ slice_steps, int nslices,
                                                         uint32_t *__restrict__ npieces, uint32_t *__restrict__ nslots) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    const int steps = slice_steps[s];
    const int np = steps <= SELL_HP ? 1 : (steps + SELL_HP - 1) / SELL_HP;
    npieces[s] = np - 1;                 // piece 0 is implicit (one warp per slice)
    nslots[s] = np > 1 ? np : 0;
}
__global__ void __launch_bounds__(256) sell_pieces_kernel(const int32_t *__restrict__ slice_steps, int nslices,
                                                          const uint32_t *__restrict__ piece_off,
                                                          const uint32_t *__restrict__ slot_off, int4 *__restrict__ pieces) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    const int steps = slice_steps[s];
    const int np = steps <= SELL_HP ? 1 : (steps + SELL_HP - 1) / SELL_HP;
    for (int q = 1; q < np; ++q)
        pieces[piece_off[s] + q - 1] = make_int4(s, q * SELL_HP, (int)slot_off[s], q | (np << 16));
}

// ---- build step 3: scatter the CSR entries into the slices ----------------------------------------------
__global__ void __launch_bounds__(256) sell_fill_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                        const double *__restrict__ vals, const int32_t *__restrict__ perm,
                                                        const int32_t *__restrict__ slice_off, int64_t npos,
                                                        double *__restrict__ svals, int32_t *__restrict__ scols) {
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int lane = (int)(pos & 31);
    const int64_t s = pos >> 5;
    const int off = slice_off[s], steps = slice_off[s + 1] - off;
    const int r = perm[pos];
    int a = 0, len = 0;
    if (r >= 0) { a = rowptr[r]; len = rowptr[r + 1] - a; }
    for (int j = 0; j < steps; ++j) {
        const size_t dst = ((size_t)off + j) * 32 + lane;
        const bool ok = j < len;
        svals[dst] = ok ? vals[a + j] : 0.0;
        scols[dst] = ok ? colidx[a + j] : 0;
    }
}

// ---- the SpMV ----------------------------------------------------------------------------------------------
// Pieces of a cut slice leave their partial sums in scratch; the last one to arrive adds them in order.
template <bool DIVIDE>
__device__ __forceinline__ void sell_finish_cut(const SellView &S, int base, int q, int np, int lane, int r, double acc,
                                                double area, double *__restrict__ y) {
    S.partial[((size_t)base + q) * 32 + lane] = acc;
    __syncwarp();
    unsigned prev = 0;
    if (lane == 0) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(S.ticket + base) : "memory");
    prev = __shfl_sync(CRG_FULL, prev, 0);
    if (prev != (unsigned)(np - 1)) return;
    double sum = 0.0;
    for (int t = 0; t < np; ++t) sum += __ldcg(&S.partial[((size_t)base + t) * 32 + lane]);
    if (r >= 0) y[r] = DIVIDE ? sum / area : sum;
    if (lane == 0) S.ticket[base] = 0;                  // ready for the next launch
}

// One warp per slice (lane = row), SELL_UNR steps in flight at once.  Warps beyond the slice range run
// the extra pieces of cut slices.
template <bool DIVIDE, int UNR>
__global__ void __launch_bounds__(256) spmv_sell_kernel(SellView S, const double *__restrict__ x, double *__restrict__ y,
                                                        const double *__restrict__ areas) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int s, j0 = 0, base = -1, q = 0;
    if (w < S.nslices) {
        s = w;
    } else {                                          // extra piece of a cut slice (rare)
        if (w - S.nslices >= S.npieces) return;
        const int4 d = __ldg(&S.pieces[w - S.nslices]);
        s = d.x; j0 = d.y; base = d.z; q = d.w & 0xffff;
    }
    const int off = __ldg(&S.slice_off[s]);
    const int steps = __ldg(&S.slice_off[s + 1]) - off;
    const int64_t pos = (int64_t)s * 32 + lane;
    const int len = __ldg(&S.rlen[pos]);
    const int r = __ldg(&S.perm[pos]);
    const int jend = min(len, j0 + SELL_HP);            // this lane's last step (exclusive) in this piece
    const int wend = min(steps, j0 + SELL_HP);          // the slice's (lane 0 holds its longest row)
    double area = 1.0;
    if (DIVIDE && r >= 0) area = __ldg(&areas[r]);
    const double *vp = S.vals + ((size_t)off + j0) * 32 + lane;
    const int32_t *cp = S.cols + ((size_t)off + j0) * 32 + lane;
    double acc = 0.0;
    for (int j = j0; j < wend; j += UNR) {
        int c[UNR];
        double v[UNR], xv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const bool ok = j + u < jend;
            c[u] = ok ? SELL_LD(cp + (size_t)u * 32) : 0;      // matrix stream: read once, evict first
            v[u] = ok ? SELL_LD(vp + (size_t)u * 32) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) xv[u] = (j + u < jend) ? __ldg(&x[c[u]]) : 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) if (j + u < jend) acc += v[u] * xv[u];
        vp += UNR * 32; cp += UNR * 32;
    }
    if (steps <= SELL_HP) {                             // the whole slice is this piece
        if (r >= 0) y[r] = DIVIDE ? acc / area : acc;
        return;
    }
    if (base < 0) base = (int)__ldg(&S.cut_base[s]);    // piece 0 of a cut slice
    sell_finish_cut<DIVIDE>(S, base, q, (steps + SELL_HP - 1) / SELL_HP, lane, r, acc, area, y);
}


// Persistent, software-pipelined variant for matrices whose slices are short (the transpose direction: ~3 entries per
// row): a warp walks slices w, w + W, w + 2W, ... three deep -- while slice s is summed, the gathers of x for slice
// s + W, the first SELL_UNR steps of (col, val) + the area of slice s + 2W and the header (offset, row length,
// permutation) of slice s + 3W are in flight, and everything a slice consumes was requested at least one iteration
// earlier: no load of an iteration depends on another load of the same iteration.  cfg5 transpose on B200: one warp per
// slice 62 us -> two-deep pipeline (next header + first steps in flight; still ~3 serialised latencies per slice)
// 56 us -> three deep 45 us (62 % of measured HBM).  Measured and dropped: a contiguous run of slices per warp
// instead of the stride W (80-200 us: the warps of a wave then hit far fewer DRAM pages at a time), 128-thread
// blocks (45-61 us, no better), 64 registers for 4 blocks per SM (spills: 78-98 us).  The grid size matters more than
// expected (8 blocks per SM in the grid, 3 resident: 45 us; 6: 51; 16: 49; 48: 63).
template <bool DIVIDE>
__global__ void __launch_bounds__(256, 3) spmv_sell_pipelined_kernel(SellView S, const double *__restrict__ x,
                                                                      double *__restrict__ y,
                                                                      const double *__restrict__ areas) {
    const int lane = threadIdx.x & 31;
    const int W = gridDim.x * (blockDim.x >> 5);
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int step = W;
    const int n = S.nslices;
    struct Hdr { int off, steps, len, r; };
    auto load_hdr = [&](int s) {
        Hdr h{0, 0, 0, -1};
        if (s < n) {
            h.off = __ldg(&S.slice_off[s]);
            h.steps = __ldg(&S.slice_off[s + 1]) - h.off;
            h.len = __ldg(&S.rlen[(int64_t)s * 32 + lane]);
            h.r = __ldg(&S.perm[(int64_t)s * 32 + lane]);
        }
        return h;
    };
    auto load_round = [&](const Hdr &h, int (&cc)[SELL_UNR], double (&vv)[SELL_UNR], double &ar) {
        const size_t e = (size_t)h.off * 32 + lane;
        const int jend = min(h.len, SELL_HP);
#pragma unroll
        for (int u = 0; u < SELL_UNR; ++u) {
            const bool ok = u < jend;
            cc[u] = ok ? SELL_LD(S.cols + e + (size_t)u * 32) : 0;
            vv[u] = ok ? SELL_LD(S.vals + e + (size_t)u * 32) : 0.0;
        }
        ar = (DIVIDE && h.r >= 0) ? __ldg(&areas[h.r]) : 1.0;
    };
    auto gather = [&](const Hdr &h, const int (&cc)[SELL_UNR], double (&xx)[SELL_UNR]) {
        const int jend = min(h.len, SELL_HP);
#pragma unroll
        for (int u = 0; u < SELL_UNR; ++u) xx[u] = (u < jend) ? __ldg(&x[cc[u]]) : 0.0;
    };
    int s = w;
    Hdr h0 = load_hdr(s), h1 = load_hdr(s + step), h2 = load_hdr(s + 2 * step);
    int c1[SELL_UNR], c2[SELL_UNR];
    double v0[SELL_UNR], v1[SELL_UNR], v2[SELL_UNR], x0[SELL_UNR], x1[SELL_UNR];
    double a0 = 1.0, a1 = 1.0, a2 = 1.0;
    {
        int c0[SELL_UNR];
        load_round(h0, c0, v0, a0);
        load_round(h1, c1, v1, a1);
        gather(h0, c0, x0);
    }
    while (s < n) {
        const Hdr h3 = load_hdr(s + 3 * step);                    // three ahead: header
        load_round(h2, c2, v2, a2);                            // two ahead: first steps + area (its header arrived an iteration ago)
        gather(h1, c1, x1);                                    // one ahead: gathers (its column indices arrived an iteration ago)
        // ---- current slice --------------------------------------------------------------------------------------
        const int jend = min(h0.len, SELL_HP), wend = min(h0.steps, SELL_HP);
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < SELL_UNR; ++u) if (u < jend) acc += v0[u] * x0[u];
        for (int j = SELL_UNR; j < wend; j += SELL_UNR) {      // taller slices: remaining steps, loaded here
            const size_t e = ((size_t)h0.off + j) * 32 + lane;
#pragma unroll
            for (int u = 0; u < SELL_UNR; ++u)
                if (j + u < jend) acc += SELL_LD(S.vals + e + (size_t)u * 32) * __ldg(&x[SELL_LD(S.cols + e + (size_t)u * 32)]);
        }
        if (h0.steps <= SELL_HP) {
            if (h0.r >= 0) y[h0.r] = DIVIDE ? acc / a0 : acc;
        } else {                                               // piece 0 of a cut slice
            sell_finish_cut<DIVIDE>(S, (int)__ldg(&S.cut_base[s]), 0, (h0.steps + SELL_HP - 1) / SELL_HP, lane, h0.r, acc, a0, y);
        }
        s += step;
        h0 = h1; h1 = h2; h2 = h3;
        a0 = a1; a1 = a2;
#pragma unroll
        for (int u = 0; u < SELL_UNR; ++u) { v0[u] = v1[u]; x0[u] = x1[u]; c1[u] = c2[u]; v1[u] = v2[u]; }
    }
    // ---- extra pieces of cut slices (rare) -----------------------------------------------------------------
    for (int p = w; p < S.npieces; p += W) {
        const int4 d = __ldg(&S.pieces[p]);
        const Hdr hp = load_hdr(d.x);
        const int jend = min(hp.len, d.y + SELL_HP), wend = min(hp.steps, d.y + SELL_HP);
        const double *vp = S.vals + ((size_t)hp.off + d.y) * 32 + lane;
        const int32_t *cp = S.cols + ((size_t)hp.off + d.y) * 32 + lane;
        double acc = 0.0;
        for (int j = d.y; j < wend; ++j, vp += 32, cp += 32)
            if (j < jend) acc += __ldg(vp) * __ldg(&x[__ldg(cp)]);
        double ar = 1.0;
        if (DIVIDE && hp.r >= 0) ar = __ldg(&areas[hp.r]);
        sell_finish_cut<DIVIDE>(S, d.z, d.w & 0xffff, d.w >> 16, lane, hp.r, acc, ar, y);
    }
}

}  // namespace crg

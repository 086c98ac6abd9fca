// crg_b200.cu -- C ABI (include/crg_b200.h) and host orchestration of the B200 engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

#include "broadphase.cuh"
#include "common.cuh"
#include "geom.cuh"
#include "gridgen.cuh"
#include "kernels.cuh"
#include "scan.cuh"
#include "sort.cuh"
#include "sell.cuh"

namespace crg {
thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
thread_local Arena *t_arena = nullptr;

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DeviceGuard {
    int prev = -1;
    int set(int dev) {
        CRG_CUDA(cudaGetDevice(&prev));
        if (dev >= 0 && dev != prev) CRG_CUDA(cudaSetDevice(dev));
        return CRG_OK;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// One CSR matrix on the device.
struct Csr {
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    DevBuf<int32_t> rowptr, colidx;
    DevBuf<double> vals;
    // SELL-32-sigma copy used by the SpMV (sell.cuh)
    DevBuf<double> sell_vals, sell_partial;
    DevBuf<int32_t> sell_cols, sell_perm, sell_rlen, sell_slice_off;
    DevBuf<int4> sell_pieces;
    DevBuf<unsigned int> sell_ticket;
    DevBuf<uint32_t> sell_cut_base;
    int sell_npieces = 0, sell_nslices = 0;
    int64_t sell_padded = 0;
    SellView sell_view() const {
        return SellView{sell_vals.p, sell_cols.p, sell_perm.p, sell_rlen.p, sell_slice_off.p, sell_pieces.p,
                        sell_cut_base.p, sell_partial.p, sell_ticket.p, sell_nslices, sell_npieces};
    }
};
}  // namespace crg

using namespace crg;

struct crg_regridder {
    crg_options opts;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int64_t n_dst = 0, n_src = 0, nnz = 0;
    Csr A;    // rows = dst, cols = src
    Csr At;   // rows = src, cols = dst  (== the reference's CSC of A)
    bool has_At = false;
    DevBuf<double> dst_areas, src_areas;
    DevBuf<double> scratch_max;
    DevBuf<int2> cand_pairs;
    int64_t n_cand_kept = 0;
    crg_build_stats stats;
    // staging for host-pointer applies
    DevBuf<double> stage_src, stage_dst;
};

namespace crg {

static int check_device_available() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_error(CRG_ERR_NO_DEVICE, "no CUDA device available (%s); libcrgb200 has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return CRG_OK;
}

static int validate_cells(const crg_cells *c, const char *name) {
    if (!c || !c->verts) return set_error(CRG_ERR_INVALID, "%s: null cells/verts", name);
    if (c->ncells < 0 || c->ncells >= ((int64_t)1 << 31))
        return set_error(CRG_ERR_INVALID, "%s: ncells=%lld out of range", name, (long long)c->ncells);
    if (!c->offsets && (c->nv < 3 || c->nv > CRG_MAX_VERTS))
        return set_error(CRG_ERR_UNSUPPORTED, "%s: nv=%d outside [3, %d]", name, c->nv, CRG_MAX_VERTS);
    return CRG_OK;
}

// A grid staged on the device.
struct DevCells {
    DevBuf<double> verts_own;
    DevBuf<int32_t> off_own;
    DevBuf<uint8_t> flip;
    DevBuf<float> diam;
    CellsView view{};
    int64_t total_verts = 0;
};

__global__ void __launch_bounds__(256) check_offsets_kernel(const int32_t *off, int64_t n, int maxv, int *bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int k = off[i + 1] - off[i];
        if (k < 3 || k > maxv) atomicExch(bad, 1);
    }
}

__global__ void __launch_bounds__(256) check_coo_kernel(const int64_t *rows, const int64_t *cols, int64_t n,
                                                        int64_t n_rows, int64_t n_cols, int *bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (rows[i] < 0 || rows[i] >= n_rows || cols[i] < 0 || cols[i] >= n_cols)) atomicExch(bad, 1);
}

// (src, dst) index lists -> int2 pairs, range-checked
__global__ void __launch_bounds__(256) pack_pairs_kernel(const int64_t *src_idx, const int64_t *dst_idx, int64_t n,
                                                         int64_t n_src, int64_t n_dst, int2 *pairs, int *bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t a = src_idx[i], b = dst_idx[i];
    const bool ok = a >= 0 && a < n_src && b >= 0 && b < n_dst;
    if (!ok) atomicExch(bad, 1);
    pairs[i] = ok ? make_int2((int)a, (int)b) : make_int2(0, 0);
}
__global__ void __launch_bounds__(256) scale_kernel(double *v, int64_t n, double f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= f;
}

// mirror_fold_partners! of the reference's Oceananigans extension (a KernelAbstractions kernel over 2 Nq slots there)
__global__ void __launch_bounds__(128) mirror_fold_kernel(double *__restrict__ f, int64_t nx, int64_t ny, int64_t K,
                                                          int64_t ld, int level_fastest) {
    const int64_t nq = nx / 4, nh = nx / 2, base = (ny - 1) * nx;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * nq * K) return;
    const int64_t k = level_fastest ? t % K : t / (2 * nq), i = level_fastest ? t / K : t % (2 * nq);
    const int64_t r = i < nq ? i : nh + (i - nq), q = nx - 1 - r;
    if (level_fastest) f[(base + q) * ld + k] = f[(base + r) * ld + k];
    else f[k * ld + base + q] = f[k * ld + base + r];
}

static int stage_cells(const crg_cells *c, int dim, cudaStream_t st, DevCells *out, const char *name) {
    const int64_t n = c->ncells;
    out->view.ncells = n;
    out->view.nv = c->nv;
    out->view.flip = nullptr;
    if (c->offsets) {
        int32_t last = 0;
        const int32_t *doff;
        if (is_device_ptr(c->offsets)) {
            doff = c->offsets;
            CRG_CUDA(cudaMemcpyAsync(&last, c->offsets + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            CRG_CUDA(cudaStreamSynchronize(st));
        } else {
            last = c->offsets[n];
            CRG_TRY(out->off_own.alloc_tmp((size_t)n + 1, st));
            CRG_CUDA(cudaMemcpyAsync(out->off_own.p, c->offsets, sizeof(int32_t) * (size_t)(n + 1),
                                     cudaMemcpyHostToDevice, st));
            doff = out->off_own.p;
        }
        out->view.off = doff;
        out->total_verts = last;
        if (n > 0) {
            DevBuf<int> bad;
            CRG_TRY(bad.alloc_tmp(1, st));
            CRG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
            check_offsets_kernel<<<ceil_div(n, 256), 256, 0, st>>>(doff, n, CRG_MAX_VERTS, bad.p);
            CRG_LAUNCH_CHECK();
            int hb = 0;
            CRG_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CRG_CUDA(cudaStreamSynchronize(st));
            if (hb) return set_error(CRG_ERR_UNSUPPORTED, "%s: a cell has fewer than 3 or more than %d vertices", name, CRG_MAX_VERTS);
        }
    } else {
        out->view.off = nullptr;
        out->total_verts = n * c->nv;
    }
    if (is_device_ptr(c->verts)) {
        out->view.verts = c->verts;
    } else {
        CRG_TRY(out->verts_own.alloc_tmp((size_t)out->total_verts * dim, st));
        CRG_CUDA(cudaMemcpyAsync(out->verts_own.p, c->verts, sizeof(double) * (size_t)out->total_verts * dim,
                                 cudaMemcpyHostToDevice, st));
        out->view.verts = out->verts_own.p;
    }
    return CRG_OK;
}

struct Timer {
    std::vector<cudaEvent_t> ev;
    cudaStream_t st;
    explicit Timer(cudaStream_t s) : st(s) {}
    ~Timer() { for (auto e : ev) cudaEventDestroy(e); }
    int mark() {
        cudaEvent_t e;
        CRG_CUDA(cudaEventCreate(&e));
        CRG_CUDA(cudaEventRecord(e, st));
        ev.push_back(e);
        return CRG_OK;
    }
    double ms(int a, int b) {
        float f = 0.f;
        if (cudaEventElapsedTime(&f, ev[a], ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.0; }
        return f;
    }
};

static int alloc_csr(Csr &M, int64_t n_rows, int64_t n_cols, int64_t nnz, cudaStream_t st) {
    M.n_rows = n_rows; M.n_cols = n_cols; M.nnz = nnz;
    CRG_TRY(M.rowptr.alloc((size_t)n_rows + 1, st));
    CRG_TRY(M.colidx.alloc((size_t)nnz, st));
    CRG_TRY(M.vals.alloc((size_t)nnz, st));
    CRG_CUDA(cudaMemsetAsync(M.rowptr.p, 0, sizeof(int32_t) * (size_t)(n_rows + 1), st));
    return CRG_OK;
}

// Sorted triples -> rowptr (+ colidx, vals when `split`) of M; low = the triples are in (col, row) order and
// M is the transpose (kernels.cuh: csr_split_kernel, rowptr_kernel).
static int fill_csr(Csr &M, const uint64_t *keys, const double *vals, int64_t nnz, bool low, bool split, cudaStream_t st) {
    CRG_CUDA(cudaMemsetAsync(M.rowptr.p, 0xFF, sizeof(int32_t) * (size_t)(M.n_rows + 1), st));
    const int g = ceil_div(nnz, 256), gr = ceil_div(M.n_rows + 1, 256);
    if (low) {
        if (split) csr_split_kernel<true, true><<<g, 256, 0, st>>>(keys, vals, nnz, M.rowptr.p, M.colidx.p, M.vals.p);
        else csr_split_kernel<true, false><<<g, 256, 0, st>>>(keys, vals, nnz, M.rowptr.p, M.colidx.p, M.vals.p);
        CRG_LAUNCH_CHECK();
        rowptr_kernel<true><<<gr, 256, 0, st>>>(keys, nnz, M.n_rows, M.rowptr.p);
    } else {
        if (split) csr_split_kernel<false, true><<<g, 256, 0, st>>>(keys, vals, nnz, M.rowptr.p, M.colidx.p, M.vals.p);
        else csr_split_kernel<false, false><<<g, 256, 0, st>>>(keys, vals, nnz, M.rowptr.p, M.colidx.p, M.vals.p);
        CRG_LAUNCH_CHECK();
        rowptr_kernel<false><<<gr, 256, 0, st>>>(keys, nnz, M.n_rows, M.rowptr.p);
    }
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

// CSR -> SELL-32-sigma (sell.cuh): window sort, slice offsets, piece list, scatter -- in two halves
// around the one host round trip (the padded size), so that two matrices on two streams can both be
// enqueued before the host waits for either.
static uint32_t *g_host_scratch[64];       // small pinned staging area per device (8 slots of 4 words)
static int host_scratch(int dev, int slot, uint32_t **out) {
    if (dev < 0 || dev >= 64) return set_error(CRG_ERR_INVALID, "device ordinal %d not supported", dev);
    if (!g_host_scratch[dev]) {
        uint32_t *p = nullptr;
        CRG_CUDA(cudaHostAlloc((void **)&p, 8 * 4 * sizeof(uint32_t), cudaHostAllocDefault));
        uint32_t *expected = nullptr;
        if (!__atomic_compare_exchange_n(&g_host_scratch[dev], &expected, p, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            cudaFreeHost(p);
    }
    *out = g_host_scratch[dev] + 4 * slot;
    return CRG_OK;
}

struct SellCtx {
    DevBuf<int32_t> steps;
    DevBuf<uint32_t> cnt, off;     // [0, nslices]: extra pieces, (nslices, 2 nslices]: partial slots
    uint32_t *h = nullptr;         // pinned: total steps, extra pieces, partial slots
    int nslices = 0;
    int64_t npos = 0;
};

static int sell_pre(Csr &M, cudaStream_t st, SellCtx &c, int device, int slot) {
    M.sell_npieces = 0;
    M.sell_padded = 0;
    if (M.n_rows == 0) return CRG_OK;
    const int nwin = (int)((M.n_rows + SELL_SIGMA - 1) / SELL_SIGMA);
    c.npos = (int64_t)nwin * SELL_SIGMA;
    c.nslices = (int)(c.npos / 32);
    const int nslices = c.nslices;
    M.sell_nslices = nslices;
    CRG_TRY(host_scratch(device, slot, &c.h));
    CRG_TRY(M.sell_perm.alloc((size_t)c.npos, st));
    CRG_TRY(M.sell_rlen.alloc((size_t)c.npos, st));
    CRG_TRY(M.sell_slice_off.alloc((size_t)nslices + 1, st));
    CRG_TRY(c.steps.alloc_tmp((size_t)nslices, st));
    CRG_TRY(c.cnt.alloc_tmp((size_t)2 * nslices, st));
    CRG_TRY(c.off.alloc_tmp((size_t)2 * nslices + 2, st));
    sell_sort_kernel<<<nwin, SELL_SIGMA, 0, st>>>(M.rowptr.p, M.n_rows, M.sell_perm.p, M.sell_rlen.p, c.steps.p);
    CRG_LAUNCH_CHECK();
    CRG_TRY((exclusive_scan<int32_t, int32_t>(c.steps.p, nslices, M.sell_slice_off.p, st)));
    sell_count_kernel<<<ceil_div(nslices, 256), 256, 0, st>>>(c.steps.p, nslices, c.cnt.p, c.cnt.p + nslices);
    CRG_LAUNCH_CHECK();
    CRG_TRY((exclusive_scan<uint32_t, uint32_t>(c.cnt.p, nslices, c.off.p, st)));
    CRG_TRY(M.sell_cut_base.alloc((size_t)nslices + 1, st));
    CRG_TRY((exclusive_scan<uint32_t, uint32_t>(c.cnt.p + nslices, nslices, M.sell_cut_base.p, st)));
    CRG_CUDA(cudaMemcpyAsync(c.h + 0, M.sell_slice_off.p + nslices, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(c.h + 1, c.off.p + nslices, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(c.h + 2, M.sell_cut_base.p + nslices, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    return CRG_OK;
}

// (after the stream of sell_pre has been synchronised)
static int sell_post(Csr &M, cudaStream_t st, SellCtx &c) {
    if (M.n_rows == 0) return CRG_OK;
    const int32_t total_steps = (int32_t)c.h[0];
    const uint32_t np = c.h[1], nslots = c.h[2];
    M.sell_npieces = (int)np;
    M.sell_padded = (int64_t)total_steps * 32;
    CRG_TRY(M.sell_vals.alloc((size_t)M.sell_padded, st));
    CRG_TRY(M.sell_cols.alloc((size_t)M.sell_padded, st));
    CRG_TRY(M.sell_pieces.alloc((size_t)np, st));
    CRG_TRY(M.sell_partial.alloc((size_t)nslots * 32, st));
    CRG_TRY(M.sell_ticket.alloc((size_t)nslots, st));
    CRG_CUDA(cudaMemsetAsync(M.sell_ticket.p, 0, sizeof(unsigned int) * (size_t)(nslots > 0 ? nslots : 1), st));
    sell_pieces_kernel<<<ceil_div(c.nslices, 256), 256, 0, st>>>(c.steps.p, c.nslices, c.off.p, M.sell_cut_base.p, M.sell_pieces.p);
    CRG_LAUNCH_CHECK();
    sell_fill_kernel<<<ceil_div(c.npos, 256), 256, 0, st>>>(M.rowptr.p, M.colidx.p, M.vals.p, M.sell_perm.p, M.sell_slice_off.p,
                                                            c.npos, M.sell_vals.p, M.sell_cols.p);
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

static int build_sell(Csr &M, cudaStream_t st, int device) {
    SellCtx c;
    CRG_TRY(sell_pre(M, st, c, device, 0));
    CRG_CUDA(cudaStreamSynchronize(st));
    return sell_post(M, st, c);
}

static int finish_csr(Csr &M, cudaStream_t st, int device) { return build_sell(M, st, device); }

// Sort COO (keys = row<<32|col, f64 values) -> CSR; optionally also the transposed CSR.
// keys/vals buffers have capacity `cap` (>= n) and are consumed.
// COO (keys = row<<32|col, f64 values) -> CSR(A) and, optionally, CSR(A^T).
// keys/vals have at least n elements and are consumed.
//   row_sorted_unique = true  (build path: the stable compaction of K3 leaves the triples grouped by
//     increasing row, every pair once): one stable pass over the column bits gives the (col, row)
//     order = CSC; a further stable pass over the row bits gives (row, col) = CSR.
//   row_sorted_unique = false (crg_build_from_coo: arbitrary order, duplicates): full-key sort,
//     segmented duplicate sum, then the column-bit passes for the transpose.
static int assemble(crg_regridder *R, DevBuf<uint64_t> &keyA, DevBuf<double> &valA, int64_t n, bool row_sorted_unique,
                    bool short_rows, cudaStream_t s2, Timer &tm, int *t_sort_csr0, int *t_sort_csr1, int *t_sort_csc1) {
    cudaStream_t st = R->stream;
    const int bits_src = ilog2_ceil((uint64_t)(R->n_src > 1 ? R->n_src : 2));
    const int bits_dst = ilog2_ceil((uint64_t)(R->n_dst > 1 ? R->n_dst : 2));
    DevBuf<uint64_t> keyB;
    DevBuf<double> valB;
    CRG_TRY(keyB.alloc_tmp((size_t)(n > 0 ? n : 1), st));
    CRG_TRY(valB.alloc_tmp((size_t)(n > 0 ? n : 1), st));
    *t_sort_csr0 = (int)tm.ev.size();
    CRG_TRY(tm.mark());
    uint64_t *ka = keyA.p, *kb = keyB.p;
    uint64_t *va = (uint64_t *)valA.p, *vb = (uint64_t *)valB.p;
    bool inb = false;
    int p1 = 0, p2 = 0, p3 = 0;
    int64_t nnz = n;
    Csr &A = R->A;
    Csr &T = R->At;
    R->has_At = false;
    if (n >= ((int64_t)1 << 31)) return set_error(CRG_ERR_NOMEM, "nnz=%lld exceeds int32 indexing", (long long)n);

    if (row_sorted_unique && short_rows) {
        // Every row is a few entries long: CSR(A) comes straight from the row-grouped triples (one thread
        // sorts a row by column); the stable column-bit passes over the same triples give the (col, row)
        // order for A^T.  The two only READ the triples, so the A side runs on the side stream while the
        // radix passes of the A^T side run on the main stream (the passes ping-pong between two scratch
        // pairs and never write the input).
        R->nnz = nnz;
        const bool fork = R->opts.build_transpose && nnz > 0 && s2 != st;
        cudaStream_t sa = fork ? s2 : st;                 // stream of the A side
        cudaEvent_t ev_in = nullptr, ev_a0 = nullptr, ev_a1 = nullptr;
        CRG_CUDA(cudaEventCreate(&ev_a0));
        CRG_CUDA(cudaEventCreate(&ev_a1));
        struct EvGuard { cudaEvent_t *e[3]; ~EvGuard() { for (auto p : e) if (*p) cudaEventDestroy(*p); } } evg{{&ev_in, &ev_a0, &ev_a1}};
        uint64_t *kt = ka, *vt = va;                      // where the (col, row)-ordered triples end up
        if (fork) {
            CRG_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
            CRG_CUDA(cudaEventRecord(ev_in, st));
            CRG_CUDA(cudaStreamWaitEvent(sa, ev_in, 0));
        }
        DevBuf<uint64_t> keyC;
        DevBuf<double> valC;
        if (R->opts.build_transpose && nnz > 0) {
            // first pass a -> b, the others between b and c
            const int first_hi = bits_src < 8 ? bits_src : 8;
            int q1 = 0, q2 = 0;
            bool in_second = false;
            CRG_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, first_hi, &inb, &q1, st));
            kt = inb ? kb : ka; vt = inb ? vb : va;
            if (bits_src > 8) {
                CRG_TRY(keyC.alloc_tmp((size_t)n, st));
                CRG_TRY(valC.alloc_tmp((size_t)n, st));
                // (n <= 1: no pass runs and the triples stay where they are)
                CRG_TRY(radix_sort_pairs(kt, vt, keyC.p, (uint64_t *)valC.p, n, 8, bits_src, &in_second, &q2, st));
                if (in_second) { kt = keyC.p; vt = (uint64_t *)valC.p; }
            }
            p1 = q1 + q2;
        }
        // ---- A side --------------------------------------------------------------------------------
        CRG_CUDA(cudaEventRecord(ev_a0, sa));
        CRG_TRY(alloc_csr(A, R->n_dst, R->n_src, nnz, sa));
        if (nnz > 0) {
            CRG_TRY(fill_csr(A, ka, nullptr, nnz, false, false, sa));
            row_sort_split_kernel<<<ceil_div(A.n_rows, ROWSORT_ROWS), ROWSORT_ROWS, 0, sa>>>(ka, (const double *)va, A.rowptr.p, A.n_rows, A.colidx.p, A.vals.p);
            CRG_LAUNCH_CHECK();
        }
        SellCtx ca, ct;
        CRG_TRY(sell_pre(A, sa, ca, R->device, 0));
        R->stats.sort_passes_csr = 0;
        // ---- A^T side, after the passes ----------------------------------------------------------------
        if (R->opts.build_transpose) {
            CRG_TRY(alloc_csr(T, R->n_src, R->n_dst, nnz, st));
            if (nnz > 0) CRG_TRY(fill_csr(T, kt, (const double *)vt, nnz, true, true, st));      // (col, row) order: rows of A^T = low word
            CRG_TRY(sell_pre(T, st, ct, R->device, 1));
        }
        // both matrices are enqueued up to their one host round trip: wait, then the second halves
        CRG_CUDA(cudaStreamSynchronize(sa));
        CRG_TRY(sell_post(A, sa, ca));
        CRG_CUDA(cudaEventRecord(ev_a1, sa));
        if (R->opts.build_transpose) {
            CRG_CUDA(cudaStreamSynchronize(st));
            CRG_TRY(sell_post(T, st, ct));
            R->has_At = true;
        }
        R->stats.sort_passes_csc = p1;
        if (fork) CRG_CUDA(cudaStreamWaitEvent(st, ev_a1, 0));          // join
        *t_sort_csr1 = -1;                                // the A side is timed by its own events (it overlaps)
        *t_sort_csc1 = (int)tm.ev.size();
        CRG_TRY(tm.mark());
        CRG_CUDA(cudaEventSynchronize(ev_a1));
        float fa = 0.f;
        if (cudaEventElapsedTime(&fa, ev_a0, ev_a1) != cudaSuccess) cudaGetLastError();
        R->stats.ms_sort_csr = fa;
        return CRG_OK;
    }

    if (row_sorted_unique) {
        R->nnz = nnz;
        // (col, row) order
        CRG_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, bits_src, &inb, &p1, st));
        if (inb) { std::swap(ka, kb); std::swap(va, vb); }
        if (R->opts.build_transpose) {
            CRG_TRY(alloc_csr(T, R->n_src, R->n_dst, nnz, st));
            if (nnz > 0) {
                CRG_TRY(fill_csr(T, ka, (const double *)va, nnz, true, true, st));      // (col, row) order: rows of A^T = low word
            }
            CRG_TRY(finish_csr(T, st, R->device));
            R->has_At = true;
        }
        R->stats.sort_passes_csc = p1;
        *t_sort_csr1 = (int)tm.ev.size();     // (long-row path only: here phase "sort_csr" = column passes + A^T, "sort_csc" = row passes + A)
        CRG_TRY(tm.mark());
        // (row, col) order
        CRG_TRY(radix_sort_pairs(ka, va, kb, vb, n, 32, 32 + bits_dst, &inb, &p2, st));
        if (inb) { std::swap(ka, kb); std::swap(va, vb); }
        R->stats.sort_passes_csr = p2;
        CRG_TRY(alloc_csr(A, R->n_dst, R->n_src, nnz, st));
        if (nnz > 0) {
            CRG_TRY(fill_csr(A, ka, (const double *)va, nnz, false, true, st));
        }
        CRG_TRY(finish_csr(A, st, R->device));
        *t_sort_csc1 = (int)tm.ev.size();
        CRG_TRY(tm.mark());
        return CRG_OK;
    }

    // ---- general path -------------------------------------------------------------------------------
    CRG_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, bits_src, &inb, &p1, st));
    if (inb) { std::swap(ka, kb); std::swap(va, vb); }
    CRG_TRY(radix_sort_pairs(ka, va, kb, vb, n, 32, 32 + bits_dst, &inb, &p2, st));
    if (inb) { std::swap(ka, kb); std::swap(va, vb); }
    R->stats.sort_passes_csr = p1 + p2;
    if (n > 0) {   // duplicate summation (segmented reduce over equal keys)
        DevBuf<uint32_t> flags, pos;
        CRG_TRY(flags.alloc_tmp((size_t)n, st));
        CRG_TRY(pos.alloc_tmp((size_t)n + 1, st));
        mark_heads_kernel<<<ceil_div(n, 256), 256, 0, st>>>(ka, n, flags.p);
        CRG_LAUNCH_CHECK();
        CRG_TRY((exclusive_scan<uint32_t, uint32_t>(flags.p, n, pos.p, st)));
        uint32_t nu = 0;
        CRG_CUDA(cudaMemcpyAsync(&nu, pos.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CRG_CUDA(cudaStreamSynchronize(st));
        if ((int64_t)nu != n) {
            dedupe_kernel<<<ceil_div(n, 256), 256, 0, st>>>(ka, (const double *)va, flags.p, pos.p, n, kb, (double *)vb);
            CRG_LAUNCH_CHECK();
            std::swap(ka, kb); std::swap(va, vb);
            nnz = nu;
        }
    }
    R->nnz = nnz;
    CRG_TRY(alloc_csr(A, R->n_dst, R->n_src, nnz, st));
    if (nnz > 0) {
        CRG_TRY(fill_csr(A, ka, (const double *)va, nnz, false, true, st));
    }
    CRG_TRY(finish_csr(A, st, R->device));
    *t_sort_csr1 = (int)tm.ev.size();
    CRG_TRY(tm.mark());
    if (R->opts.build_transpose) {   // swap the key halves and (stably) sort by the new high word only
        CRG_TRY(alloc_csr(T, R->n_src, R->n_dst, nnz, st));
        if (nnz > 0) {
            swap_key_kernel<<<ceil_div(nnz, 256), 256, 0, st>>>(ka, nnz, ka);
            CRG_LAUNCH_CHECK();
            CRG_TRY(radix_sort_pairs(ka, va, kb, vb, nnz, 32, 32 + bits_src, &inb, &p3, st));
            if (inb) { std::swap(ka, kb); std::swap(va, vb); }
            CRG_TRY(fill_csr(T, ka, (const double *)va, nnz, false, true, st));
        }
        CRG_TRY(finish_csr(T, st, R->device));
        R->stats.sort_passes_csc = p3;
        R->has_At = true;
    }
    *t_sort_csc1 = (int)tm.ev.size();
    CRG_TRY(tm.mark());
    return CRG_OK;
}

// maximum(A) into R->scratch_max (device); 0 for an empty matrix
static int device_maximum(crg_regridder *R) {
    cudaStream_t st = R->stream;
    CRG_TRY(R->scratch_max.alloc(1, st));
    CRG_CUDA(cudaMemsetAsync(R->scratch_max.p, 0, sizeof(double), st));
    if (R->nnz > 0) {
        max_kernel<<<296, 256, 0, st>>>(R->A.vals.p, R->nnz, R->scratch_max.p);
        CRG_LAUNCH_CHECK();
    }
    return CRG_OK;
}

// A, A^T (CSR and SELL copies) and both area vectors divided by the value in R->scratch_max
static int divide_by_scratch(crg_regridder *R) {
    cudaStream_t st = R->stream;
    if (R->nnz > 0) {
        div_by_kernel<<<296, 256, 0, st>>>(R->A.vals.p, R->nnz, R->scratch_max.p);
        CRG_LAUNCH_CHECK();
        if (R->has_At) { div_by_kernel<<<296, 256, 0, st>>>(R->At.vals.p, R->nnz, R->scratch_max.p); CRG_LAUNCH_CHECK(); }
        for (Csr *M : {&R->A, &R->At})
            if (M->sell_padded > 0) { div_by_kernel<<<296, 256, 0, st>>>(M->sell_vals.p, M->sell_padded, R->scratch_max.p); CRG_LAUNCH_CHECK(); }
    }
    if (R->n_dst > 0) { div_by_kernel<<<148, 256, 0, st>>>(R->dst_areas.p, R->n_dst, R->scratch_max.p); CRG_LAUNCH_CHECK(); }
    if (R->n_src > 0) { div_by_kernel<<<148, 256, 0, st>>>(R->src_areas.p, R->n_src, R->scratch_max.p); CRG_LAUNCH_CHECK(); }
    return CRG_OK;
}

static int do_normalize(crg_regridder *R) {
    if (R->nnz == 0) return CRG_OK;   // maximum() of an empty matrix: nothing to scale
    CRG_TRY(device_maximum(R));
    return divide_by_scratch(R);
}

static int device_side_stream(int dev, cudaStream_t *out);

// K3 launcher: dense areas of `n_cand` (src, dst) pairs + survivor counts per CLIP_TILE pairs (kernels.cuh).
// Quadrilateral grids take the symbolic-polygon kernel with in-kernel queues, everything else the general one.
template <int DIM>
static int launch_clip(const CellsView &gdv, const CellsView &gsv, bool fixed, int nv_dst, int nv_src, const int2 *pairs,
                       int64_t n_cand, double thresh, const double *unit_src_areas, double *pair_area,
                       uint32_t *tile_count, cudaStream_t st, const double *clip_nrm = nullptr) {
    static const bool allow_quad = !(getenv("CRG_CLIP_QUAD") && atoi(getenv("CRG_CLIP_QUAD")) == 0);
    const bool fixed4 = fixed && nv_dst <= 4 && nv_src <= 4;
    const bool quad = allow_quad && fixed4 && nv_dst == 4 && nv_src == 4 && ((uintptr_t)gdv.verts % 16 == 0) &&
                      ((uintptr_t)gsv.verts % 16 == 0);
#define CRG_CLIP(NT_, MW_)                                                                                            \
    do {                                                                                                              \
        const size_t smem = sizeof(double) * 2 * MW_ * DIM * NT_;                                                     \
        auto kern = clip_kernel<DIM, NT_, MW_>;                                                                       \
        CRG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        kern<<<ceil_div(n_cand, NT_), NT_, smem, st>>>(gdv, gsv, pairs, n_cand, thresh, pair_area, tile_count);       \
    } while (0)
    static const bool allow_fast = getenv("CRG_CLIP_FAST") && atoi(getenv("CRG_CLIP_FAST")) != 0;   // opt-in: measured slower than clip_quad_kernel overall (profiles/README.md)
    if (quad && DIM == 3 && allow_fast) {
        // Spherical quadrilaterals: wedge sums for the pairs cut by one edge or two adjacent ones, symbolic
        // Sutherland-Hodgman for the rest, one kernel (kernels.cuh).
        DevBuf<double> nrm, corners;
        CRG_TRY(nrm.alloc_tmp((size_t)gdv.ncells * 12, st));
        CRG_TRY(corners.alloc_tmp((size_t)gdv.ncells * 12, st));
        quad_normals_kernel<<<ceil_div(gdv.ncells, 256), 256, 0, st>>>(gdv, nrm.p, corners.p);
        CRG_LAUNCH_CHECK();
        const int grid = ceil_div(ceil_div(n_cand, CF_CHUNK), CF_NT / 32);
        const size_t smem = (size_t)(CF_NT / 32) * CF_WARP_SMEM;
        clip_quad_fast_kernel<<<grid, CF_NT, smem, st>>>(gdv, gsv, nrm.p, corners.p, pairs, n_cand, thresh, unit_src_areas, pair_area,
                                                         tile_count);
    } else if (quad) {
        constexpr int NT = 128;
        const size_t smem = sizeof(double) * QUAD_SLOTS * DIM * NT;
        // 256-bit vertex loads when every cell record is 32-byte aligned (spherical quadrilaterals are 96 bytes)
        static const bool allow_wide = !(getenv("CRG_CLIP_WIDE") && atoi(getenv("CRG_CLIP_WIDE")) == 0);
        const double *nrm = ((uintptr_t)clip_nrm % 32 == 0) ? clip_nrm : nullptr;
        const bool wide = allow_wide && DIM == 3 && (uintptr_t)gdv.verts % 32 == 0 && (uintptr_t)gsv.verts % 32 == 0;
        auto kern = wide ? clip_quad_kernel<DIM, NT, true> : clip_quad_kernel<DIM, NT, false>;
        CRG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // an untouched source cell contributes its own area: reuse K4's (only when they are unit-sphere areas)
        kern<<<ceil_div(ceil_div(n_cand, CLIP_CHUNK), NT / 32), NT, smem, st>>>(gdv, gsv, pairs, n_cand, thresh,
                                                                                unit_src_areas, pair_area, tile_count, nrm);
    } else if (fixed4) CRG_CLIP(128, 8);
    else CRG_CLIP(64, 2 * CRG_MAX_VERTS);
#undef CRG_CLIP
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

template <int DIM>
static int build_impl(crg_regridder *R, const crg_cells *dst, const crg_cells *src) {
    cudaStream_t st = R->stream;
    crg_build_stats &S = R->stats;
    const double t_begin = now_ms();
    Timer tm(st);

    // ---- stage inputs -------------------------------------------------------------------
    DevCells gd, gs;
    CRG_TRY(stage_cells(dst, DIM, st, &gd, "dst"));
    CRG_TRY(stage_cells(src, DIM, st, &gs, "src"));
    if (gd.verts_own.p || gs.verts_own.p || gd.off_own.p || gs.off_own.p) CRG_CUDA(cudaStreamSynchronize(st));   // host inputs: time the upload
    S.ms_h2d = now_ms() - t_begin;
    const int64_t nd = R->n_dst, ns = R->n_src;
    const double r2 = DIM == 3 ? R->opts.radius * R->opts.radius : 1.0;

    CRG_TRY(tm.mark());   // 0
    // ---- K4: areas + orientation ----------------------------------------------------------
    CRG_TRY(R->dst_areas.alloc((size_t)nd, st));
    CRG_TRY(R->src_areas.alloc((size_t)ns, st));
    CRG_TRY(gd.flip.alloc_tmp((size_t)nd, st));
    CRG_TRY(gs.flip.alloc_tmp((size_t)ns, st));
    DevBuf<unsigned int> nflip;      // [0..1] clockwise cells (dst, src), [2..3] non-convex cells (dst, src)
    CRG_TRY(nflip.alloc_tmp(4, st));
    CRG_CUDA(cudaMemsetAsync(nflip.p, 0, 4 * sizeof(unsigned int), st));
    CRG_TRY(tm.mark());   // 1

    // ---- K1: bounds -------------------------------------------------------------------------
    CRG_TRY(gd.diam.alloc_tmp((size_t)nd, st));
    CRG_TRY(gs.diam.alloc_tmp((size_t)ns, st));
    DevBuf<BPStats> dstats;
    CRG_TRY(dstats.alloc_tmp(2, st));
    BPStats hst[2];
    memset(hst, 0, sizeof(hst));
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 3; ++j) { hst[k].lo[j] = ~0ull; hst[k].hi[j] = 0ull; }
    CRG_CUDA(cudaMemcpyAsync(dstats.p, hst, sizeof(hst), cudaMemcpyHostToDevice, st));
    const double big_chord = 2.0 * std::sin(BP_BIG_ANGLE / 2.0);
    // (the views' flip pointers are still null here: the kernel sees the cells as stored)
    // spherical quadrilaterals: the destination (clip) grid's edge-plane normals come out of the same pass
    DevBuf<double> dst_nrm;
    static const bool allow_nrm = !(getenv("CRG_CLIP_NORMALS") && atoi(getenv("CRG_CLIP_NORMALS")) == 0);
    if (allow_nrm && DIM == 3 && nd && !dst->offsets && !src->offsets && dst->nv == 4 && src->nv == 4)
        CRG_TRY(dst_nrm.alloc_tmp((size_t)nd * 12, st));
    if (nd) { bp_bounds_kernel<DIM><<<ceil_div(nd, 256), 256, 0, st>>>(gd.view, gd.diam.p, dstats.p, (float)big_chord, r2, R->dst_areas.p, gd.flip.p, nflip.p, dst_nrm.p); CRG_LAUNCH_CHECK(); }
    if (ns) { bp_bounds_kernel<DIM><<<ceil_div(ns, 256), 256, 0, st>>>(gs.view, gs.diam.p, dstats.p + 1, (float)big_chord, r2, R->src_areas.p, gs.flip.p, nflip.p + 1); CRG_LAUNCH_CHECK(); }
    gd.view.flip = gd.flip.p;
    gs.view.flip = gs.flip.p;
    unsigned int h_nflip[4] = {0, 0, 0, 0};
    CRG_CUDA(cudaMemcpyAsync(hst, dstats.p, sizeof(hst), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(h_nflip, nflip.p, sizeof(h_nflip), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    if (h_nflip[2] || h_nflip[3])
        return set_error(CRG_ERR_UNSUPPORTED,
                         "%u destination and %u source cells are not convex: the device clip is convex-convex "
                         "Sutherland-Hodgman (split such cells into convex parts; the Python front end does it for planar grids)",
                         h_nflip[2], h_nflip[3]);
    CRG_TRY(tm.mark());   // 2

    // ---- choose the bin grid ----------------------------------------------------------------
    BPParams P;
    memset(&P, 0, sizeof(P));
    P.dim = DIM;
    P.big_chord = big_chord;
    const int NB_CAP = 2048;
    const double mean_src = hst[1].count ? hst[1].sum_diam / (double)hst[1].count : 0.0;
    if (DIM == 3) {
        P.nfaces = 6;
        double margin = hst[0].count ? 1.4 * 2.0 * std::asin(std::fmin(1.0, 0.5 * (double)hst[0].max_diam)) + 1e-6 : 0.0;
        if (margin > 0.3) margin = 0.3;
        const double A = 0.25 * M_PI + margin;      // half-width of a face's domain as an angle
        const double F = std::sin(A);               // ... and in face coordinates f = sin(angle)
        static const double bin_scale = getenv("CRG_BIN_SCALE") ? atof(getenv("CRG_BIN_SCALE")) : 1.5;   // measured on cfg5: 0.5 -> 5.49 ms, 0.75 -> 4.92, 1.0 -> 4.74, 1.5 -> 4.68, 2.0 -> 4.67
        double h = bin_scale * mean_src;            // target bin size (radians)
        if (!(h > 2.0 * A / NB_CAP)) h = 2.0 * A / NB_CAP;
        if (hst[1].count == 0) h = 2.0 * A;
        int nb = (int)std::ceil(2.0 * A / h);
        if (nb < 1) nb = 1;
        if (nb > NB_CAP) nb = NB_CAP;
        h = 2.0 * A / nb;
        P.nbx = P.nby = nb;
        P.ox = P.oy = -F;
        P.hx = P.hy = F;
        P.inv_hq = BP_SUB / (2.0 * F / nb);
        P.eps = 1e-6;
        {   // Source cells that cannot meet any destination cell are not binned.  Every point of a non-big
            // destination cell lies within the box of the destination vertices inflated by the bulge of the
            // sphere over the cell, at most 1 - cos(diam) <= diam^2/2; a source cell meeting it has its first vertex within one source
            // diameter (chord) of that point.  Big destination cells pair with everything: no culling.
            const bool all_small = hst[0].count == (unsigned long long)nd && nd > 0;
            const double dd = (double)hst[0].max_diam, ds = (double)hst[1].max_diam;
            const double infl = ds + 0.5 * dd * dd * 1.05 + 1e-6;
            for (int j = 0; j < 3; ++j) {
                P.cull_lo[j] = all_small ? (float)(ordered_bits_to_double(hst[0].lo[j]) - infl) : -2.f;
                P.cull_hi[j] = all_small ? (float)(ordered_bits_to_double(hst[0].hi[j]) + infl) : 2.f;
            }
        }
        {   // quick-reject bound of cell_face_qbox: tan(A + 4 * largest non-big cell diameter as an angle)
            const double dmax = std::fmax((double)hst[0].max_diam, (double)hst[1].max_diam);
            const double ang = A + 4.0 * 2.0 * std::asin(std::fmin(1.0, 0.5 * dmax));
            P.u_reject = (float)(std::tan(std::fmin(ang, 1.5)) * (1.0 + 1e-5));
        }
        S.bin_size = h;
    } else {
        P.nfaces = 1;
        double lo[2], hi[2];
        for (int k = 0; k < 2; ++k) {
            const unsigned long long l = hst[0].lo[k] < hst[1].lo[k] ? hst[0].lo[k] : hst[1].lo[k];
            const unsigned long long u = hst[0].hi[k] > hst[1].hi[k] ? hst[0].hi[k] : hst[1].hi[k];
            lo[k] = (nd + ns) ? ordered_bits_to_double(l) : 0.0;
            hi[k] = (nd + ns) ? ordered_bits_to_double(u) : 1.0;
        }
        const double ext = std::fmax(std::fmax(hi[0] - lo[0], hi[1] - lo[1]),
                                     std::fmax(std::fmax(std::fabs(lo[0]), std::fabs(hi[0])),
                                               std::fmax(std::fabs(lo[1]), std::fabs(hi[1]))));
        P.eps = 1e-9 * (ext > 0 ? ext : 1.0);
        P.ox = lo[0] - 4 * P.eps; P.oy = lo[1] - 4 * P.eps;
        P.hx = hi[0] + 4 * P.eps; P.hy = hi[1] + 4 * P.eps;
        const double Lx = P.hx - P.ox, Ly = P.hy - P.oy;
        double h = 1.5 * mean_src;
        const double hmin = std::fmax(Lx, Ly) / NB_CAP;
        if (!(h > hmin)) h = hmin;
        if (!(h > 0)) h = 1.0;
        P.nbx = (int)std::fmin((double)NB_CAP, std::fmax(1.0, std::ceil(Lx / h)));
        P.nby = (int)std::fmin((double)NB_CAP, std::fmax(1.0, std::ceil(Ly / h)));
        P.inv_hq = BP_SUB / h;
        P.big_chord = 1e300;
        P.u_reject = 3.0e38f;
        S.bin_size = h;
    }
    const size_t nbins = (size_t)P.nfaces * P.nbx * P.nby;
    S.n_bins = (int64_t)nbins;

    // ---- K2a: bin the source cells (count / scan / fill) ------------------------------------
    DevBuf<uint32_t> bin_count, bin_start, counters;
    DevBuf<int32_t> big_src, big_dst;
    CRG_TRY(bin_count.alloc_tmp(nbins + 1, st));
    CRG_TRY(bin_start.alloc_tmp(nbins + 1, st));
    CRG_TRY(counters.alloc_tmp(4, st));
    CRG_TRY(big_src.alloc_tmp((size_t)ns, st));
    CRG_TRY(big_dst.alloc_tmp((size_t)nd, st));
    CRG_CUDA(cudaMemsetAsync(bin_count.p, 0, sizeof(uint32_t) * (nbins + 1), st));
    CRG_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(uint32_t) * 4, st));
    DevBuf<int4> bin_rec;                     // the count pass' box records, consumed by the fill pass
    static const bool allow_rec = !(getenv("CRG_BIN_RECORDS") && atoi(getenv("CRG_BIN_RECORDS")) == 0);
    if (allow_rec && ns) CRG_TRY(bin_rec.alloc_tmp((size_t)ns, st));
    if (ns) bp_bin_kernel<DIM, false><<<ceil_div(ns, 256), 256, 0, st>>>(gs.view, gs.diam.p, P, bin_count.p, nullptr,
                                                                          nullptr, big_src.p, counters.p, bin_rec.p);
    if (ns) CRG_LAUNCH_CHECK();
    CRG_TRY((exclusive_scan<uint32_t, uint32_t>(bin_count.p, (int64_t)nbins, bin_start.p, st)));
    uint32_t h_entries = 0, h_counters[4] = {0, 0, 0, 0};
    CRG_CUDA(cudaMemcpyAsync(&h_entries, bin_start.p + nbins, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    const int n_big_src = (int)h_counters[0];
    S.n_bin_entries = h_entries;
    S.n_big_src = n_big_src;
    DevBuf<int4> entries;
    CRG_TRY(entries.alloc_tmp((size_t)h_entries, st));
    CRG_CUDA(cudaMemsetAsync(bin_count.p, 0, sizeof(uint32_t) * (nbins + 1), st));
    if (ns && bin_rec.p) bp_bin_fill_kernel<DIM><<<ceil_div(ns, 256), 256, 0, st>>>(gs.view, P, bin_rec.p, bin_count.p, bin_start.p, entries.p);
    else if (ns) bp_bin_kernel<DIM, true><<<ceil_div(ns, 256), 256, 0, st>>>(gs.view, gs.diam.p, P, bin_count.p, bin_start.p,
                                                                              entries.p, nullptr, nullptr);
    if (ns) CRG_LAUNCH_CHECK();
    CRG_TRY(tm.mark());   // 3

    // ---- K2b: destination queries (count / scan / fill) ---------------------------------------
    DevBuf<uint32_t> cand_count;
    DevBuf<int64_t> cand_off;
    CRG_TRY(cand_count.alloc_tmp((size_t)nd + 1, st));
    CRG_TRY(cand_off.alloc_tmp((size_t)nd + 1, st));
    DevBuf<int32_t> slab;
    static const bool allow_slab = !(getenv("CRG_QUERY_SLAB") && atoi(getenv("CRG_QUERY_SLAB")) == 0);
    if (allow_slab && nd) CRG_TRY(slab.alloc_tmp((size_t)BP_SLAB * nd, st));
    DevBuf<int32_t> heavy_list;
    DevBuf<uint32_t> heavy_counter;
    static const bool allow_heavy = !(getenv("CRG_QUERY_HEAVY") && atoi(getenv("CRG_QUERY_HEAVY")) == 0);
    CRG_TRY(heavy_counter.alloc_tmp(1, st));
    CRG_CUDA(cudaMemsetAsync(heavy_counter.p, 0, sizeof(uint32_t), st));
    if (allow_heavy && nd) CRG_TRY(heavy_list.alloc_tmp((size_t)nd, st));
    constexpr int HEAVY_GRID = 148 * 4;
    if (nd) bp_query_kernel<DIM, false><<<ceil_div(nd, 128), 128, 0, st>>>(
        gd.view, gd.diam.p, P, bin_start.p, entries.p, big_src.p, n_big_src, ns, cand_count.p, nullptr, nullptr,
        big_dst.p, counters.p + 1, slab.p, heavy_list.p, heavy_counter.p);
    if (nd) CRG_LAUNCH_CHECK();
    if (nd && heavy_list.p) {
        bp_query_heavy_kernel<DIM, false><<<HEAVY_GRID, 128, 0, st>>>(gd.view, P, bin_start.p, entries.p, big_src.p, n_big_src,
                                                                     cand_count.p, nullptr, nullptr, counters.p + 1, slab.p,
                                                                     heavy_list.p, heavy_counter.p);
        CRG_LAUNCH_CHECK();
    }
    CRG_TRY((exclusive_scan<uint32_t, int64_t>(cand_count.p, nd, cand_off.p, st)));
    int64_t n_cand = 0;
    CRG_CUDA(cudaMemcpyAsync(&n_cand, cand_off.p + nd, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    const int n_big_dst = (int)h_counters[1];
    static const bool allow_rowsort = !(getenv("CRG_ROW_SORT") && atoi(getenv("CRG_ROW_SORT")) == 0);
    const bool short_rows = allow_rowsort && h_counters[2] == 0;   // no destination cell has more than BP_SHORT_ROW candidates
    S.n_candidates = n_cand;
    S.n_big_dst = n_big_dst;
    if (n_cand >= ((int64_t)1 << 32))
        return set_error(CRG_ERR_NOMEM, "candidate pair count %lld exceeds 2^32", (long long)n_cand);
    DevBuf<int2> pairs;
    if (R->opts.keep_candidates) CRG_TRY(pairs.alloc((size_t)n_cand, st));   // outlives the build
    else CRG_TRY(pairs.alloc_tmp((size_t)n_cand, st));
    if (nd && slab.p && h_counters[3] == 0) {            // every candidate is in the slab: copy, no second traversal
        bp_fill_slab_kernel<<<ceil_div(nd, 256), 256, 0, st>>>(slab.p, nd, cand_count.p, cand_off.p, pairs.p);
    } else if (nd) {
        bp_query_kernel<DIM, true><<<ceil_div(nd, 128), 128, 0, st>>>(
            gd.view, gd.diam.p, P, bin_start.p, entries.p, big_src.p, n_big_src, ns, nullptr, cand_off.p, pairs.p, nullptr,
            nullptr, nullptr, heavy_list.p, heavy_counter.p);
        if (heavy_list.p) {
            CRG_LAUNCH_CHECK();
            bp_query_heavy_kernel<DIM, true><<<HEAVY_GRID, 128, 0, st>>>(gd.view, P, bin_start.p, entries.p, big_src.p, n_big_src,
                                                                        nullptr, cand_off.p, pairs.p, nullptr, nullptr,
                                                                        heavy_list.p, heavy_counter.p);
        }
    }
    if (nd) CRG_LAUNCH_CHECK();
    if (n_big_dst && ns) {
        dim3 grid((unsigned)std::min<int64_t>(ceil_div(ns, 256), 1024), (unsigned)n_big_dst);
        bp_fill_big_dst_kernel<<<grid, 256, 0, st>>>(big_dst.p, cand_off.p, ns, pairs.p);
        CRG_LAUNCH_CHECK();
    }
    entries.release(); bin_count.release(); bin_start.release();
    CRG_TRY(tm.mark());   // 4

    // ---- K3: clip + area, compaction ------------------------------------------------------------
    DevBuf<uint64_t> coo_key;
    DevBuf<double> coo_val;
    DevBuf<double> pair_area;
    DevBuf<uint32_t> tile_count;
    const int64_t ntiles = (n_cand + CLIP_TILE - 1) / CLIP_TILE;
    uint32_t h_keep = 0;
    if (n_cand > 0) {
        CRG_TRY(pair_area.alloc_tmp((size_t)n_cand, st));
        CRG_TRY(tile_count.alloc_tmp((size_t)ntiles + 1, st));
        CRG_CUDA(cudaMemsetAsync(tile_count.p, 0, sizeof(uint32_t) * ((size_t)ntiles + 1), st));
        CRG_TRY((launch_clip<DIM>(gd.view, gs.view, !dst->offsets && !src->offsets, dst->nv, src->nv, pairs.p, n_cand,
                                  R->opts.area_threshold, r2 == 1.0 ? R->src_areas.p : nullptr, pair_area.p,
                                  tile_count.p, st, dst_nrm.p)));
        CRG_TRY((exclusive_scan<uint32_t, uint32_t>(tile_count.p, ntiles, tile_count.p, st)));
        CRG_CUDA(cudaMemcpyAsync(&h_keep, tile_count.p + ntiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CRG_CUDA(cudaStreamSynchronize(st));
        CRG_TRY(coo_key.alloc_tmp((size_t)h_keep, st));
        CRG_TRY(coo_val.alloc_tmp((size_t)h_keep, st));
        compact_pairs_kernel<<<(unsigned)ntiles, 256, 0, st>>>(pairs.p, pair_area.p, n_cand, tile_count.p, r2, coo_key.p,
                                                              coo_val.p);
        CRG_LAUNCH_CHECK();
        pair_area.release();
    } else {
        CRG_TRY(coo_key.alloc_tmp(1, st));
        CRG_TRY(coo_val.alloc_tmp(1, st));
    }
    if (R->opts.keep_candidates) {
        R->cand_pairs = std::move(pairs);
        R->n_cand_kept = n_cand;
    } else {
        pairs.release();
    }
    CRG_TRY(tm.mark());   // 5

    // ---- K5: assembly ------------------------------------------------------------------------------
    int t0 = 0, t1 = 0, t2 = 0;
    cudaStream_t side_stream = st;
    static const bool two_streams = !(getenv("CRG_TWO_STREAMS") && atoi(getenv("CRG_TWO_STREAMS")) == 0);
    if (two_streams) CRG_TRY(device_side_stream(R->device, &side_stream));
    CRG_TRY(assemble(R, coo_key, coo_val, (int64_t)h_keep, true, short_rows, side_stream, tm, &t0, &t1, &t2));
    coo_key.release(); coo_val.release();

    // ---- K6: normalize ------------------------------------------------------------------------------
    if (R->opts.normalize) CRG_TRY(do_normalize(R));
    CRG_TRY(tm.mark());   // last
    CRG_CUDA(cudaStreamSynchronize(st));
    const int last = (int)tm.ev.size() - 1;
    S.n_dst = nd; S.n_src = ns; S.nnz = R->nnz;
    S.ms_areas = tm.ms(0, 1);
    S.ms_bounds = tm.ms(1, 2);
    S.ms_bin = tm.ms(2, 3);
    S.ms_query = tm.ms(3, 4);
    S.ms_clip = tm.ms(4, 5);
    if (t1 >= 0) { S.ms_sort_csr = tm.ms(t0, t1); S.ms_sort_csc = tm.ms(t1, t2); }
    else S.ms_sort_csc = tm.ms(t0, t2);      // short-row path: the A side (ms_sort_csr, own events) runs inside this span
    S.ms_finish = tm.ms(t2, last);
    S.ms_device = tm.ms(0, last);
    S.ms_total = now_ms() - t_begin;
    return CRG_OK;
}

// compute_intersection_areas (intersection_areas.jl:4-32) for an explicit pair list, with the kernels of the build.
template <int DIM>
static int clip_pairs_impl(const crg_options *opts, const crg_cells *dst, const crg_cells *src, int64_t n_pairs,
                           const int64_t *src_idx, const int64_t *dst_idx, double *area_out, int dev, cudaStream_t st) {
    DevCells gd, gs;
    CRG_TRY(stage_cells(dst, DIM, st, &gd, "dst"));
    CRG_TRY(stage_cells(src, DIM, st, &gs, "src"));
    const int64_t nd = dst->ncells, ns = src->ncells;
    DevBuf<double> a_dst, a_src;
    DevBuf<unsigned int> nflip;
    DevBuf<BPStats> dstats;
    CRG_TRY(a_dst.alloc_tmp((size_t)nd, st));
    CRG_TRY(a_src.alloc_tmp((size_t)ns, st));
    CRG_TRY(gd.flip.alloc_tmp((size_t)nd, st));
    CRG_TRY(gs.flip.alloc_tmp((size_t)ns, st));
    CRG_TRY(gd.diam.alloc_tmp((size_t)nd, st));
    CRG_TRY(gs.diam.alloc_tmp((size_t)ns, st));
    CRG_TRY(nflip.alloc_tmp(4, st));
    CRG_TRY(dstats.alloc_tmp(2, st));
    CRG_CUDA(cudaMemsetAsync(nflip.p, 0, 4 * sizeof(unsigned int), st));
    CRG_CUDA(cudaMemsetAsync(dstats.p, 0, 2 * sizeof(BPStats), st));
    const float big_chord = 1e30f;
    if (nd) { bp_bounds_kernel<DIM><<<ceil_div(nd, 256), 256, 0, st>>>(gd.view, gd.diam.p, dstats.p, big_chord, 1.0, a_dst.p, gd.flip.p, nflip.p); CRG_LAUNCH_CHECK(); }
    if (ns) { bp_bounds_kernel<DIM><<<ceil_div(ns, 256), 256, 0, st>>>(gs.view, gs.diam.p, dstats.p + 1, big_chord, 1.0, a_src.p, gs.flip.p, nflip.p + 1); CRG_LAUNCH_CHECK(); }
    gd.view.flip = gd.flip.p;
    gs.view.flip = gs.flip.p;
    DevBuf<int64_t> di, si;
    DevBuf<int2> pairs;
    DevBuf<int> bad;
    DevBuf<double> pair_area;
    DevBuf<uint32_t> tile_count;
    const int64_t ntiles = (n_pairs + CLIP_TILE - 1) / CLIP_TILE;
    CRG_TRY(di.alloc_tmp((size_t)n_pairs, st));
    CRG_TRY(si.alloc_tmp((size_t)n_pairs, st));
    CRG_TRY(pairs.alloc_tmp((size_t)n_pairs, st));
    CRG_TRY(bad.alloc_tmp(1, st));
    CRG_TRY(pair_area.alloc_tmp((size_t)n_pairs, st));
    CRG_TRY(tile_count.alloc_tmp((size_t)ntiles + 1, st));
    CRG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    CRG_CUDA(cudaMemsetAsync(tile_count.p, 0, sizeof(uint32_t) * ((size_t)ntiles + 1), st));
    CRG_CUDA(cudaMemcpyAsync(si.p, src_idx, sizeof(int64_t) * (size_t)n_pairs, cudaMemcpyDefault, st));
    CRG_CUDA(cudaMemcpyAsync(di.p, dst_idx, sizeof(int64_t) * (size_t)n_pairs, cudaMemcpyDefault, st));
    pack_pairs_kernel<<<ceil_div(n_pairs, 256), 256, 0, st>>>(si.p, di.p, n_pairs, ns, nd, pairs.p, bad.p);
    CRG_LAUNCH_CHECK();
    unsigned int h_nflip[4] = {0, 0, 0, 0};
    int hb = 0;
    CRG_CUDA(cudaMemcpyAsync(h_nflip, nflip.p, sizeof(h_nflip), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    if (hb) return set_error(CRG_ERR_INVALID, "crg_clip_pairs: a cell index is out of range");
    if (h_nflip[2] || h_nflip[3])
        return set_error(CRG_ERR_UNSUPPORTED, "%u destination and %u source cells are not convex", h_nflip[2], h_nflip[3]);
    CRG_TRY((launch_clip<DIM>(gd.view, gs.view, !dst->offsets && !src->offsets, dst->nv, src->nv, pairs.p, n_pairs,
                              opts->area_threshold, a_src.p, pair_area.p, tile_count.p, st)));
    const double r2 = DIM == 3 ? opts->radius * opts->radius : 1.0;
    if (r2 != 1.0) { scale_kernel<<<ceil_div(n_pairs, 256), 256, 0, st>>>(pair_area.p, n_pairs, r2); CRG_LAUNCH_CHECK(); }
    CRG_CUDA(cudaMemcpyAsync(area_out, pair_area.p, sizeof(double) * (size_t)n_pairs, cudaMemcpyDefault, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    return CRG_OK;
}

// One library-owned stream per device, created on first use and shared by every handle that is
// not given a caller stream: creating/destroying a stream per handle defeats the stream-ordered
// allocator's block reuse (measured: ~30 ms per build on cfg5).
static cudaStream_t g_dev_stream[64];
static int device_stream(int dev, cudaStream_t *out) {
    if (dev < 0 || dev >= 64) return set_error(CRG_ERR_INVALID, "device ordinal %d not supported", dev);
    if (!g_dev_stream[dev]) {
        cudaStream_t s;
        CRG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaStream_t expected = nullptr;
        if (!__atomic_compare_exchange_n(&g_dev_stream[dev], &expected, s, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            cudaStreamDestroy(s);
    }
    *out = g_dev_stream[dev];
    return CRG_OK;
}

// Per-device build arena + the mutex that serialises builds on one device.
// ... and one side stream per device: the assembly builds CSR(A) on it while the radix passes of A^T run
static cudaStream_t g_dev_stream2[64];
static int device_side_stream(int dev, cudaStream_t *out) {
    if (dev < 0 || dev >= 64) return set_error(CRG_ERR_INVALID, "device ordinal %d not supported", dev);
    if (!g_dev_stream2[dev]) {
        cudaStream_t s;
        CRG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaStream_t expected = nullptr;
        if (!__atomic_compare_exchange_n(&g_dev_stream2[dev], &expected, s, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            cudaStreamDestroy(s);
    }
    *out = g_dev_stream2[dev];
    return CRG_OK;
}

static Arena g_arena[64];
static std::mutex g_build_mutex[64];

struct ArenaScope {
    std::unique_lock<std::mutex> lock;
    Arena *a = nullptr;
    int begin(int dev, cudaStream_t st) {
        if (dev < 0 || dev >= 64) return CRG_OK;        // no arena: everything falls back to the pool
        lock = std::unique_lock<std::mutex>(g_build_mutex[dev]);
        a = &g_arena[dev];
        if (a->want > a->cap) {                          // the previous build overflowed: grow (rare)
            CRG_CUDA(cudaStreamSynchronize(st));
            if (a->base) CRG_CUDA(cudaFree(a->base));
            a->base = nullptr;
            const size_t cap = a->want + a->want / 4 + (64u << 20);
            if (cudaMalloc((void **)&a->base, cap) == cudaSuccess) a->cap = cap;
            else { cudaGetLastError(); a->base = nullptr; a->cap = 0; }
        }
        a->off = 0;
        a->want = 0;
        t_arena = a;
        return CRG_OK;
    }
    ~ArenaScope() { t_arena = nullptr; }
};

static int new_handle(const crg_options *opts, crg_regridder **out, DeviceGuard &guard) {
    CRG_TRY(check_device_available());
    int ndev = 0;
    CRG_CUDA(cudaGetDeviceCount(&ndev));
    if (opts->device >= ndev) return set_error(CRG_ERR_INVALID, "device %d out of range (%d devices)", opts->device, ndev);
    CRG_TRY(guard.set(opts->device));
    crg_regridder *R = new (std::nothrow) crg_regridder();
    if (!R) return set_error(CRG_ERR_NOMEM, "out of host memory");
    R->opts = *opts;
    memset(&R->stats, 0, sizeof(R->stats));
    cudaError_t e = cudaGetDevice(&R->device);
    if (e != cudaSuccess) { delete R; return fail_cuda(e, "cudaGetDevice", __FILE__, __LINE__); }
    int rc = device_stream(R->device, &R->own_stream);
    if (rc != CRG_OK) { delete R; return rc; }
    R->stream = opts->stream ? (cudaStream_t)opts->stream : R->own_stream;
    // keep freed blocks in the pool: repeated builds/applies do not hit the driver allocator
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, R->device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    *out = R;
    return CRG_OK;
}

// A failed build may have forked work to the device's side stream (assemble): both streams are drained before
// the temporaries (arena) and the handle's buffers are released.
static void drain_after_failure(crg_regridder *R) {
    cudaStreamSynchronize(R->stream);
    if (R->device >= 0 && R->device < 64 && g_dev_stream2[R->device]) cudaStreamSynchronize(g_dev_stream2[R->device]);
    cudaGetLastError();
}

static void destroy_handle(crg_regridder *R) {
    if (!R) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(R->device);
    // Everything is released stream-ordered on the stream the handle WORKS on (crg_options.stream / crg_set_stream, else
    // the library stream): work still in flight there (crg_apply_async) is ordered before the release, and the next
    // build on the same stream gets the blocks back at once.  (Releasing on the library stream behind an event of the
    // caller's stream was measured: the pool cannot hand such blocks to an allocation that the host enqueues before the
    // event has completed, and a steady loop of builds paid ~1 ms per build for fresh driver allocations.)
    // The caller's stream must therefore outlive the handle (include/crg_b200.h: crg_free).
    cudaStream_t st = R->stream ? R->stream : R->own_stream;
    auto rebind = [&](auto &buf) { buf.s = st; buf.release(); };
    rebind(R->A.rowptr); rebind(R->A.colidx); rebind(R->A.vals);
    rebind(R->At.rowptr); rebind(R->At.colidx); rebind(R->At.vals);
    for (Csr *M : {&R->A, &R->At}) {
        rebind(M->sell_vals); rebind(M->sell_partial); rebind(M->sell_cols); rebind(M->sell_perm); rebind(M->sell_rlen);
        rebind(M->sell_slice_off); rebind(M->sell_pieces); rebind(M->sell_ticket); rebind(M->sell_cut_base);
    }
    rebind(R->dst_areas); rebind(R->src_areas); rebind(R->scratch_max); rebind(R->cand_pairs);
    rebind(R->stage_src); rebind(R->stage_dst);
    delete R;   // buffers were released stream-ordered; no sync needed
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

static int apply_impl(crg_regridder *R, int transpose, int divide, double *dst, const double *src, int64_t K,
                      int64_t ld_dst, int64_t ld_src, int level_fastest, bool allow_host, bool sync) {
    if (!R || !dst || !src) return set_error(CRG_ERR_INVALID, "crg_apply: null argument");
    if (K < 1) return set_error(CRG_ERR_INVALID, "crg_apply: K=%lld must be >= 1", (long long)K);
    if (transpose && !R->has_At) return set_error(CRG_ERR_INVALID, "crg_apply: transpose requested but the regridder was built with build_transpose=0");
    const Csr &M = transpose ? R->At : R->A;
    const double *areas = transpose ? R->src_areas.p : R->dst_areas.p;
    const int64_t n_out = M.n_rows, n_in = M.n_cols;
    if (K == 1) { if (ld_dst <= 0) ld_dst = level_fastest ? 1 : n_out; if (ld_src <= 0) ld_src = level_fastest ? 1 : n_in; }
    if (level_fastest) {
        if (ld_dst < K || ld_src < K) return set_error(CRG_ERR_INVALID, "crg_apply: level-fastest leading dimensions (%lld, %lld) smaller than K=%lld", (long long)ld_dst, (long long)ld_src, (long long)K);
    } else {
        if (ld_dst < n_out || ld_src < n_in) return set_error(CRG_ERR_INVALID, "crg_apply: cell-fastest leading dimensions (%lld, %lld) smaller than (%lld, %lld)", (long long)ld_dst, (long long)ld_src, (long long)n_out, (long long)n_in);
    }
    DeviceGuard guard;
    CRG_TRY(guard.set(R->device));
    cudaStream_t st = R->stream;
    const bool src_dev = is_device_ptr(src), dst_dev = is_device_ptr(dst);
    if ((!src_dev || !dst_dev) && !allow_host)
        return set_error(CRG_ERR_INVALID, "crg_apply_async: host pointers are not allowed");
    const size_t src_elems = level_fastest ? (size_t)(n_in - 1) * ld_src + K : (size_t)(K - 1) * ld_src + n_in;
    const size_t dst_elems = level_fastest ? (size_t)(n_out - 1) * ld_dst + K : (size_t)(K - 1) * ld_dst + n_out;
    const double *xs = src;
    double *yd = dst;
    if (!src_dev) {
        if (R->stage_src.n < src_elems) CRG_TRY(R->stage_src.alloc(src_elems, st));
        CRG_CUDA(cudaMemcpyAsync(R->stage_src.p, src, sizeof(double) * src_elems, cudaMemcpyHostToDevice, st));
        xs = R->stage_src.p;
    }
    if (!dst_dev) {
        if (R->stage_dst.n < dst_elems) CRG_TRY(R->stage_dst.alloc(dst_elems, st));
        yd = R->stage_dst.p;
        if ((level_fastest ? ld_dst != K : ld_dst != n_out))   // padded layout: keep the caller's padding bytes
            CRG_CUDA(cudaMemcpyAsync(yd, dst, sizeof(double) * dst_elems, cudaMemcpyHostToDevice, st));
    }
    if (n_out > 0) {
        if (K == 1 && (level_fastest ? (ld_src == 1 && ld_dst == 1) : true)) {
            if (M.nnz == 0) {
                CRG_CUDA(cudaMemsetAsync(yd, 0, sizeof(double) * (size_t)n_out, st));
            } else {
                const SellView V = M.sell_view();
                static const int mode = getenv("CRG_SELL_MODE") ? atoi(getenv("CRG_SELL_MODE")) : 0;
                static const int bps = getenv("CRG_SELL_BPS") ? atoi(getenv("CRG_SELL_BPS")) : 8;
                const double mean_steps = (double)M.sell_padded / 32.0 / (double)std::max(1, M.sell_nslices);
                const bool pipelined = mode == 1 || (mode == 0 && mean_steps <= 6.0);
                if (pipelined) {
                    static int n_sm = 0;
                    if (!n_sm) CRG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, R->device));
                    const int grid = std::min(ceil_div(M.sell_nslices, 8), bps * n_sm);
                    if (divide) spmv_sell_pipelined_kernel<true><<<grid, 256, 0, st>>>(V, xs, yd, areas);
                    else spmv_sell_pipelined_kernel<false><<<grid, 256, 0, st>>>(V, xs, yd, areas);
                } else {
                    const int nw = M.sell_nslices + M.sell_npieces;  // one warp per slice + extra pieces of cut slices
                    constexpr int BS = 128;                          // measured: 128 > 256 threads, 4 steps in flight > 2, 6, 8
                    if (divide) spmv_sell_kernel<true, SELL_UNR><<<ceil_div(nw, BS / 32), BS, 0, st>>>(V, xs, yd, areas);
                    else spmv_sell_kernel<false, SELL_UNR><<<ceil_div(nw, BS / 32), BS, 0, st>>>(V, xs, yd, areas);
                }
            }
        } else if (level_fastest) {
            const int nblk = ceil_div(n_out, 8);
#define CRG_LF(KT)                                                                                                   \
    do {                                                                                                             \
        if (divide) spmm_lf_kernel<KT, true><<<nblk, 256, 0, st>>>(M.rowptr.p, M.colidx.p, M.vals.p, xs, yd, areas, n_out, K, ld_src, ld_dst); \
        else spmm_lf_kernel<KT, false><<<nblk, 256, 0, st>>>(M.rowptr.p, M.colidx.p, M.vals.p, xs, yd, areas, n_out, K, ld_src, ld_dst);      \
    } while (0)
            // (measured alternatives on cfg3, K = 100, 65.5 us here: double2 lanes + shuffled entries 78 us,
            //  a warp owning 4-8 consecutive rows with one coalesced entry load 82 us)
            if (K <= 32) CRG_LF(1);
            else if (K <= 64) CRG_LF(2);
            else CRG_LF(4);
#undef CRG_LF
        } else {
            constexpr int KC = 8;
            dim3 grid((unsigned)ceil_div(n_out, 128), (unsigned)ceil_div(K, KC));
            if (divide) spmm_cf_kernel<KC, true><<<grid, 128, 0, st>>>(M.rowptr.p, M.colidx.p, M.vals.p, xs, yd, areas, n_out, K, ld_src, ld_dst);
            else spmm_cf_kernel<KC, false><<<grid, 128, 0, st>>>(M.rowptr.p, M.colidx.p, M.vals.p, xs, yd, areas, n_out, K, ld_src, ld_dst);
        }
        CRG_LAUNCH_CHECK();
    }
    if (!dst_dev) CRG_CUDA(cudaMemcpyAsync(dst, yd, sizeof(double) * dst_elems, cudaMemcpyDeviceToHost, st));
    if (sync || !dst_dev || !src_dev) CRG_CUDA(cudaStreamSynchronize(st));
    return CRG_OK;
}

template <typename T>
static int copy_out(T *dst, const T *dev_src, size_t n, cudaStream_t st) {
    if (!dst || n == 0) return CRG_OK;
    CRG_CUDA(cudaMemcpyAsync(dst, dev_src, sizeof(T) * n, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return CRG_OK;
}

static int export_csr_matrix(const crg_regridder *R, const Csr &M, int32_t base, int64_t *ptr, int64_t *idx, double *val) {
    DeviceGuard guard;
    CRG_TRY(guard.set(R->device));
    cudaStream_t st = R->stream;
    std::vector<int32_t> tmp;
    if (ptr) {
        tmp.resize((size_t)M.n_rows + 1);
        CRG_CUDA(cudaMemcpyAsync(tmp.data(), M.rowptr.p, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost, st));
        CRG_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < tmp.size(); ++i) ptr[i] = (int64_t)tmp[i] + base;
    }
    if (idx && M.nnz) {
        tmp.resize((size_t)M.nnz);
        CRG_CUDA(cudaMemcpyAsync(tmp.data(), M.colidx.p, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost, st));
        CRG_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < (size_t)M.nnz; ++i) idx[i] = (int64_t)tmp[i] + base;
    }
    if (val && M.nnz) {
        CRG_CUDA(cudaMemcpyAsync(val, M.vals.p, sizeof(double) * (size_t)M.nnz, cudaMemcpyDeviceToHost, st));
        CRG_CUDA(cudaStreamSynchronize(st));
    }
    return CRG_OK;
}

// ---- described grids ---------------------------------------------------------------------------
static int grid_ncells_full(const crg_grid *g, int64_t *n);
// cells the descriptor stands for: the whole grid, or its slice [cell_lo, cell_hi)
static int grid_ncells(const crg_grid *g, int64_t *n) {
    CRG_TRY(grid_ncells_full(g, n));
    if (g->kind != CRG_GRID_CELLS && (g->cell_lo != 0 || g->cell_hi != 0)) {
        if (g->cell_lo < 0 || g->cell_hi < g->cell_lo || g->cell_hi > *n)
            return set_error(CRG_ERR_INVALID, "grid slice [%lld, %lld) outside [0, %lld)", (long long)g->cell_lo, (long long)g->cell_hi, (long long)*n);
        *n = g->cell_hi - g->cell_lo;
    }
    return CRG_OK;
}
static int grid_ncells_full(const crg_grid *g, int64_t *n) {
    if (!g) return set_error(CRG_ERR_INVALID, "null grid");
    switch (g->kind) {
        case CRG_GRID_CELLS: *n = g->cells.ncells; return CRG_OK;
        case CRG_GRID_LONLAT:
        case CRG_GRID_FULL_RING:
            if (g->n1 < 1 || g->n2 < 1) return set_error(CRG_ERR_INVALID, "grid: n1, n2 must be positive");
            if (g->kind == CRG_GRID_FULL_RING && !g->lat_deg) return set_error(CRG_ERR_INVALID, "full-ring grid: null lat_deg");
            *n = g->n1 * g->n2; return CRG_OK;
        case CRG_GRID_HEALPIX:
            if (g->n1 < 1 || (g->n1 & (g->n1 - 1)) || g->n1 > (1 << 13)) return set_error(CRG_ERR_INVALID, "healpix: nside must be a power of two <= 8192");
            *n = 12 * g->n1 * g->n1; return CRG_OK;
        case CRG_GRID_CUBED_SPHERE:
            if (g->n1 < 1) return set_error(CRG_ERR_INVALID, "cubed sphere: n must be positive");
            *n = 6 * g->n1 * g->n1; return CRG_OK;
        case CRG_GRID_REDUCED_RING: {
            const int64_t a = (int64_t)g->p[1], b = (int64_t)g->p[2], nh = g->n2 / 2;
            if (g->n2 < 2 || (g->n2 & 1) || a < 1 || b < 0 || (double)a != g->p[1] || (double)b != g->p[2])
                return set_error(CRG_ERR_INVALID, "reduced ring grid: n2 (rings) must be even and >= 2, p[1] >= 1 and p[2] >= 0 integers");
            if (!g->lat_deg) return set_error(CRG_ERR_INVALID, "reduced ring grid: null lat_deg");
            *n = 2 * (a * nh + b * (nh * (nh + 1) / 2)); return CRG_OK;
        }
        default: return set_error(CRG_ERR_INVALID, "unknown grid kind %d", g->kind);
    }
}

// Generate the [ncells][4][3] vertices of a described grid into `verts` (device memory).
static int generate_grid(const crg_grid *g, int64_t n, double *verts, cudaStream_t st) {
    if (n == 0) return CRG_OK;
    const int nblk = ceil_div(n, 256);
    const int64_t c0 = g->cell_lo, c1 = g->cell_lo + n;       // (whole grid: cell_lo = 0, n = all)
    switch (g->kind) {
        case CRG_GRID_LONLAT:
            gen_lonlat_kernel<<<nblk, 256, 0, st>>>(g->n1, g->n2, g->p[0], g->p[1], g->p[2], g->p[3], c0, c1, verts);
            break;
        case CRG_GRID_HEALPIX:
            gen_healpix_kernel<<<nblk, 256, 0, st>>>(g->n1, g->flags & 1, c0, c1, verts);
            break;
        case CRG_GRID_CUBED_SPHERE:
            gen_cubed_sphere_kernel<<<nblk, 256, 0, st>>>(g->n1, c0, c1, verts);
            break;
        case CRG_GRID_REDUCED_RING:
        case CRG_GRID_FULL_RING: {
            const double *lat = g->lat_deg;
            DevBuf<double> dlat;
            if (!is_device_ptr(lat)) {
                CRG_TRY(dlat.alloc_tmp((size_t)g->n2, st));
                CRG_CUDA(cudaMemcpyAsync(dlat.p, lat, sizeof(double) * (size_t)g->n2, cudaMemcpyHostToDevice, st));
                lat = dlat.p;
            }
            if (g->kind == CRG_GRID_FULL_RING) gen_full_ring_kernel<<<nblk, 256, 0, st>>>(g->n1, g->n2, g->p[0], lat, c0, c1, verts);
            else gen_reduced_ring_kernel<<<nblk, 256, 0, st>>>(g->n2, (int64_t)g->p[1], (int64_t)g->p[2], g->p[0], lat, c0, c1, verts);
            break;
        }
        default: return set_error(CRG_ERR_INVALID, "grid kind %d cannot be generated", g->kind);
    }
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

template <int DIM>
static int grid_areas_impl(const crg_options *opts, const crg_cells *c, double *areas, cudaStream_t st) {
    DevCells g;
    CRG_TRY(stage_cells(c, DIM, st, &g, "grid"));
    const int64_t n = c->ncells;
    DevBuf<double> a;
    DevBuf<unsigned int> nflip;
    DevBuf<BPStats> dstats;
    CRG_TRY(a.alloc_tmp((size_t)n, st));
    CRG_TRY(g.flip.alloc_tmp((size_t)n, st));
    CRG_TRY(g.diam.alloc_tmp((size_t)n, st));
    CRG_TRY(nflip.alloc_tmp(4, st));
    CRG_TRY(dstats.alloc_tmp(1, st));
    CRG_CUDA(cudaMemsetAsync(nflip.p, 0, 4 * sizeof(unsigned int), st));
    CRG_CUDA(cudaMemsetAsync(dstats.p, 0, sizeof(BPStats), st));
    const double r2 = DIM == 3 ? opts->radius * opts->radius : 1.0;
    bp_bounds_kernel<DIM><<<ceil_div(n, 256), 256, 0, st>>>(g.view, g.diam.p, dstats.p, 1e30f, r2, a.p, g.flip.p, nflip.p);
    CRG_LAUNCH_CHECK();
    CRG_CUDA(cudaMemcpyAsync(areas, a.p, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
    CRG_CUDA(cudaStreamSynchronize(st));
    return CRG_OK;
}

}  // namespace crg

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char *crg_last_error(void) { return g_err; }
const char *crg_version(void) { return "crg_b200 0.1.0 (sm_100a)"; }

int crg_launch_count(uint64_t *count) {
    if (!count) return set_error(CRG_ERR_INVALID, "null count");
    *count = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    return CRG_OK;
}

int crg_device_count(int32_t *count) {
    if (!count) return set_error(CRG_ERR_INVALID, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return CRG_OK;
}

int crg_options_init(crg_options *o) {
    if (!o) return set_error(CRG_ERR_INVALID, "null options");
    memset(o, 0, sizeof(*o));
    o->manifold = CRG_SPHERICAL;
    o->normalize = 0;
    o->radius = 1.0;
    o->area_threshold = 0.0;
    o->device = -1;
    o->build_transpose = 1;
    o->keep_candidates = 0;
    o->stream = nullptr;
    return CRG_OK;
}

int crg_build(const crg_options *opts, const crg_cells *dst, const crg_cells *src, crg_regridder **out) {
    if (!opts || !out) return set_error(CRG_ERR_INVALID, "crg_build: null argument");
    *out = nullptr;
    if (opts->manifold != CRG_PLANAR && opts->manifold != CRG_SPHERICAL)
        return set_error(CRG_ERR_INVALID, "crg_build: unknown manifold %d", opts->manifold);
    if (!(opts->radius > 0.0)) return set_error(CRG_ERR_INVALID, "crg_build: radius must be positive");
    CRG_TRY(validate_cells(dst, "dst"));
    CRG_TRY(validate_cells(src, "src"));
    DeviceGuard guard;
    crg_regridder *R = nullptr;
    CRG_TRY(new_handle(opts, &R, guard));
    R->n_dst = dst->ncells;
    R->n_src = src->ncells;
    int rc;
    {
        ArenaScope arena;
        rc = arena.begin(R->device, R->stream);
        if (rc == CRG_OK) rc = opts->manifold == CRG_SPHERICAL ? build_impl<3>(R, dst, src) : build_impl<2>(R, dst, src);
        if (rc != CRG_OK) drain_after_failure(R);
    }
    if (rc != CRG_OK) { destroy_handle(R); return rc; }
    *out = R;
    return CRG_OK;
}

int crg_grid_ncells(const crg_grid *g, int64_t *ncells) {
    if (!ncells) return set_error(CRG_ERR_INVALID, "null argument");
    return grid_ncells(g, ncells);
}

int crg_grid_cells(const crg_grid *g, int32_t device, double *verts) {
    int64_t n = 0;
    CRG_TRY(grid_ncells(g, &n));
    if (!verts) return set_error(CRG_ERR_INVALID, "null verts");
    if (g->kind == CRG_GRID_CELLS) return set_error(CRG_ERR_INVALID, "grid is already explicit");
    CRG_TRY(check_device_available());
    DeviceGuard guard;
    CRG_TRY(guard.set(device));
    int dev = 0;
    CRG_CUDA(cudaGetDevice(&dev));
    cudaStream_t st;
    CRG_TRY(device_stream(dev, &st));
    if (is_device_ptr(verts)) {
        CRG_TRY(generate_grid(g, n, verts, st));
    } else {
        DevBuf<double> tmp;
        CRG_TRY(tmp.alloc((size_t)n * 12, st));
        CRG_TRY(generate_grid(g, n, tmp.p, st));
        CRG_CUDA(cudaMemcpyAsync(verts, tmp.p, sizeof(double) * (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    }
    CRG_CUDA(cudaStreamSynchronize(st));
    return CRG_OK;
}

int crg_grid_areas(const crg_options *opts, const crg_grid *g, double *areas) {
    if (!opts || !g) return set_error(CRG_ERR_INVALID, "crg_grid_areas: null argument");
    int64_t n = 0;
    CRG_TRY(grid_ncells(g, &n));
    if (n > 0 && !areas) return set_error(CRG_ERR_INVALID, "crg_grid_areas: null areas");
    if (g->kind == CRG_GRID_CELLS) CRG_TRY(validate_cells(&g->cells, "grid"));
    else if (opts->manifold != CRG_SPHERICAL) return set_error(CRG_ERR_INVALID, "crg_grid_areas: described grids live on the sphere");
    if (!(opts->radius > 0.0)) return set_error(CRG_ERR_INVALID, "crg_grid_areas: radius must be positive");
    CRG_TRY(check_device_available());
    if (n == 0) return CRG_OK;
    DeviceGuard guard;
    CRG_TRY(guard.set(opts->device));
    int dev = 0;
    CRG_CUDA(cudaGetDevice(&dev));
    cudaStream_t st = (cudaStream_t)opts->stream;
    if (!st) CRG_TRY(device_stream(dev, &st));
    ArenaScope arena;
    int rc = arena.begin(dev, st);
    DevBuf<double> verts;
    crg_cells c = g->cells;
    if (rc == CRG_OK && g->kind != CRG_GRID_CELLS) {
        rc = verts.alloc_tmp((size_t)n * 12, st);
        if (rc == CRG_OK) rc = generate_grid(g, n, verts.p, st);
        c.verts = verts.p; c.offsets = nullptr; c.ncells = n; c.nv = 4; c.reserved = 0;
    }
    if (rc == CRG_OK) rc = (g->kind != CRG_GRID_CELLS || opts->manifold == CRG_SPHERICAL) ? grid_areas_impl<3>(opts, &c, areas, st)
                                                                                         : grid_areas_impl<2>(opts, &c, areas, st);
    if (rc != CRG_OK) { cudaStreamSynchronize(st); cudaGetLastError(); }
    return rc;
}

int crg_build_grids(const crg_options *opts, const crg_grid *dst, const crg_grid *src, crg_regridder **out) {
    if (!opts || !out || !dst || !src) return set_error(CRG_ERR_INVALID, "crg_build_grids: null argument");
    *out = nullptr;
    if (dst->kind == CRG_GRID_CELLS && src->kind == CRG_GRID_CELLS) return crg_build(opts, &dst->cells, &src->cells, out);
    if (opts->manifold != CRG_SPHERICAL) return set_error(CRG_ERR_INVALID, "crg_build_grids: described grids live on the sphere");
    if (!(opts->radius > 0.0)) return set_error(CRG_ERR_INVALID, "crg_build_grids: radius must be positive");
    int64_t nd = 0, ns = 0;
    CRG_TRY(grid_ncells(dst, &nd));
    CRG_TRY(grid_ncells(src, &ns));
    if (dst->kind == CRG_GRID_CELLS) CRG_TRY(validate_cells(&dst->cells, "dst"));
    if (src->kind == CRG_GRID_CELLS) CRG_TRY(validate_cells(&src->cells, "src"));
    if (nd >= ((int64_t)1 << 31) || ns >= ((int64_t)1 << 31)) return set_error(CRG_ERR_INVALID, "grid too large");
    DeviceGuard guard;
    crg_regridder *R = nullptr;
    CRG_TRY(new_handle(opts, &R, guard));
    R->n_dst = nd;
    R->n_src = ns;
    int rc;
    {
        ArenaScope arena;
        rc = arena.begin(R->device, R->stream);
        DevBuf<double> vd, vs;
        crg_cells cd = dst->cells, cs = src->cells;
        auto materialise = [&](const crg_grid *g, int64_t n, DevBuf<double> &buf, crg_cells *c) -> int {
            if (g->kind == CRG_GRID_CELLS) return CRG_OK;
            CRG_TRY(buf.alloc_tmp((size_t)(n > 0 ? n : 1) * 12, R->stream));
            CRG_TRY(generate_grid(g, n, buf.p, R->stream));
            c->verts = buf.p; c->offsets = nullptr; c->ncells = n; c->nv = 4; c->reserved = 0;
            return CRG_OK;
        };
        if (rc == CRG_OK) rc = materialise(dst, nd, vd, &cd);
        if (rc == CRG_OK) rc = materialise(src, ns, vs, &cs);
        if (rc == CRG_OK) rc = build_impl<3>(R, &cd, &cs);
        if (rc != CRG_OK) drain_after_failure(R);
    }
    if (rc != CRG_OK) { destroy_handle(R); return rc; }
    *out = R;
    return CRG_OK;
}

int crg_build_from_coo(const crg_options *opts, int64_t n_dst, int64_t n_src, int64_t nnz, const int64_t *dst_idx,
                       const int64_t *src_idx, const double *area, const double *dst_areas, const double *src_areas,
                       crg_regridder **out) {
    if (!opts || !out) return set_error(CRG_ERR_INVALID, "crg_build_from_coo: null argument");
    *out = nullptr;
    if (n_dst < 0 || n_src < 0 || nnz < 0 || n_dst >= ((int64_t)1 << 31) || n_src >= ((int64_t)1 << 31))
        return set_error(CRG_ERR_INVALID, "crg_build_from_coo: bad sizes");
    if (nnz > 0 && (!dst_idx || !src_idx || !area)) return set_error(CRG_ERR_INVALID, "crg_build_from_coo: null triples");
    // host-resident index arrays are range-checked here, device-resident ones by check_coo_kernel below
    const bool dst_idx_dev = is_device_ptr(dst_idx), src_idx_dev = is_device_ptr(src_idx);
    for (int64_t k = 0; k < nnz; ++k)
        if ((!dst_idx_dev && (dst_idx[k] < 0 || dst_idx[k] >= n_dst)) || (!src_idx_dev && (src_idx[k] < 0 || src_idx[k] >= n_src)))
            return set_error(CRG_ERR_INVALID, "crg_build_from_coo: index out of range at entry %lld", (long long)k);
    DeviceGuard guard;
    crg_regridder *R = nullptr;
    CRG_TRY(new_handle(opts, &R, guard));
    R->n_dst = n_dst;
    R->n_src = n_src;
    auto body = [&]() -> int {
        cudaStream_t st = R->stream;
        const double t0 = now_ms();
        Timer tm(st);
        DevBuf<int64_t> dr, dc;
        DevBuf<uint64_t> keys;
        DevBuf<double> vals;
        const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
        CRG_TRY(dr.alloc_tmp(n1, st)); CRG_TRY(dc.alloc_tmp(n1, st)); CRG_TRY(keys.alloc_tmp(n1, st)); CRG_TRY(vals.alloc_tmp(n1, st));
        if (nnz) {
            CRG_CUDA(cudaMemcpyAsync(dr.p, dst_idx, sizeof(int64_t) * nnz, cudaMemcpyDefault, st));
            CRG_CUDA(cudaMemcpyAsync(dc.p, src_idx, sizeof(int64_t) * nnz, cudaMemcpyDefault, st));
            CRG_CUDA(cudaMemcpyAsync(vals.p, area, sizeof(double) * nnz, cudaMemcpyDefault, st));
            if (dst_idx_dev || src_idx_dev) {
                DevBuf<int> bad;
                CRG_TRY(bad.alloc_tmp(1, st));
                CRG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
                check_coo_kernel<<<ceil_div(nnz, 256), 256, 0, st>>>(dr.p, dc.p, nnz, n_dst, n_src, bad.p);
                CRG_LAUNCH_CHECK();
                int hb = 0;
                CRG_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                CRG_CUDA(cudaStreamSynchronize(st));
                if (hb) return set_error(CRG_ERR_INVALID, "crg_build_from_coo: an index is out of range");
            }
            pack_coo_kernel<<<ceil_div(nnz, 256), 256, 0, st>>>(dr.p, dc.p, nnz, keys.p);
            CRG_LAUNCH_CHECK();
        }
        CRG_TRY(R->dst_areas.alloc((size_t)n_dst, st));
        CRG_TRY(R->src_areas.alloc((size_t)n_src, st));
        if (dst_areas && n_dst) CRG_CUDA(cudaMemcpyAsync(R->dst_areas.p, dst_areas, sizeof(double) * n_dst, cudaMemcpyDefault, st));
        else CRG_CUDA(cudaMemsetAsync(R->dst_areas.p, 0, sizeof(double) * (size_t)(n_dst > 0 ? n_dst : 1), st));
        if (src_areas && n_src) CRG_CUDA(cudaMemcpyAsync(R->src_areas.p, src_areas, sizeof(double) * n_src, cudaMemcpyDefault, st));
        else CRG_CUDA(cudaMemsetAsync(R->src_areas.p, 0, sizeof(double) * (size_t)(n_src > 0 ? n_src : 1), st));
        CRG_TRY(tm.mark());
        int a = 0, b = 0, c = 0;
        CRG_TRY(assemble(R, keys, vals, nnz, false, false, R->stream, tm, &a, &b, &c));
        if (R->opts.normalize) CRG_TRY(do_normalize(R));
        CRG_TRY(tm.mark());
        CRG_CUDA(cudaStreamSynchronize(st));
        crg_build_stats &S = R->stats;
        S.n_dst = n_dst; S.n_src = n_src; S.nnz = R->nnz; S.n_candidates = nnz;
        S.ms_sort_csr = tm.ms(a, b); S.ms_sort_csc = tm.ms(b, c);
        S.ms_device = tm.ms(0, (int)tm.ev.size() - 1);
        S.ms_total = now_ms() - t0;
        return CRG_OK;
    };
    int rc;
    {
        ArenaScope arena;
        rc = arena.begin(R->device, R->stream);
        if (rc == CRG_OK) rc = body();
        if (rc != CRG_OK) drain_after_failure(R);
    }
    if (rc != CRG_OK) { destroy_handle(R); return rc; }
    *out = R;
    return CRG_OK;
}

int crg_clip_pairs(const crg_options *opts, const crg_cells *dst, const crg_cells *src, int64_t n_pairs,
                   const int64_t *src_idx, const int64_t *dst_idx, double *area_out) {
    if (!opts) return set_error(CRG_ERR_INVALID, "crg_clip_pairs: null options");
    if (n_pairs < 0 || n_pairs >= ((int64_t)1 << 32)) return set_error(CRG_ERR_INVALID, "crg_clip_pairs: bad pair count");
    if (n_pairs > 0 && (!src_idx || !dst_idx || !area_out)) return set_error(CRG_ERR_INVALID, "crg_clip_pairs: null argument");
    if (opts->manifold != CRG_PLANAR && opts->manifold != CRG_SPHERICAL)
        return set_error(CRG_ERR_INVALID, "crg_clip_pairs: unknown manifold %d", opts->manifold);
    if (!(opts->radius > 0.0)) return set_error(CRG_ERR_INVALID, "crg_clip_pairs: radius must be positive");
    CRG_TRY(validate_cells(dst, "dst"));
    CRG_TRY(validate_cells(src, "src"));
    CRG_TRY(check_device_available());
    if (n_pairs == 0) return CRG_OK;
    DeviceGuard guard;
    CRG_TRY(guard.set(opts->device));
    int dev = 0;
    CRG_CUDA(cudaGetDevice(&dev));
    cudaStream_t st = (cudaStream_t)opts->stream;
    if (!st) CRG_TRY(device_stream(dev, &st));
    ArenaScope arena;
    int rc = arena.begin(dev, st);
    if (rc == CRG_OK)
        rc = opts->manifold == CRG_SPHERICAL ? clip_pairs_impl<3>(opts, dst, src, n_pairs, src_idx, dst_idx, area_out, dev, st)
                                             : clip_pairs_impl<2>(opts, dst, src, n_pairs, src_idx, dst_idx, area_out, dev, st);
    if (rc != CRG_OK) { cudaStreamSynchronize(st); cudaGetLastError(); }
    return rc;
}

int crg_set_areas(crg_regridder *r, const double *dst_areas, const double *src_areas) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    if (dst_areas && r->n_dst) CRG_CUDA(cudaMemcpyAsync(r->dst_areas.p, dst_areas, sizeof(double) * (size_t)r->n_dst, cudaMemcpyDefault, r->stream));
    if (src_areas && r->n_src) CRG_CUDA(cudaMemcpyAsync(r->src_areas.p, src_areas, sizeof(double) * (size_t)r->n_src, cudaMemcpyDefault, r->stream));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_free(crg_regridder *r) {
    destroy_handle(r);
    return CRG_OK;
}

int crg_dims(const crg_regridder *r, int64_t *n_dst, int64_t *n_src, int64_t *nnz) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    if (n_dst) *n_dst = r->n_dst;
    if (n_src) *n_src = r->n_src;
    if (nnz) *nnz = r->nnz;
    return CRG_OK;
}

int crg_stats(const crg_regridder *r, crg_build_stats *s) {
    if (!r || !s) return set_error(CRG_ERR_INVALID, "null argument");
    *s = r->stats;
    return CRG_OK;
}

int crg_areas(const crg_regridder *r, double *dst_areas, double *src_areas) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    CRG_TRY(copy_out(dst_areas, r->dst_areas.p, (size_t)r->n_dst, r->stream));
    CRG_TRY(copy_out(src_areas, r->src_areas.p, (size_t)r->n_src, r->stream));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_export_csc(const crg_regridder *r, int32_t base, int64_t *colptr, int64_t *rowval, double *nzval) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    if (!r->has_At) return set_error(CRG_ERR_INVALID, "crg_export_csc needs build_transpose=1");
    return export_csr_matrix(r, r->At, base, colptr, rowval, nzval);
}

int crg_export_csr(const crg_regridder *r, int32_t base, int64_t *rowptr, int64_t *colval, double *nzval) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    return export_csr_matrix(r, r->A, base, rowptr, colval, nzval);
}

int crg_candidates(const crg_regridder *r, int64_t *src_idx, int64_t *dst_idx) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    if (!r->opts.keep_candidates) return set_error(CRG_ERR_INVALID, "regridder was built without keep_candidates");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    std::vector<int2> tmp((size_t)r->n_cand_kept);
    if (r->n_cand_kept) {
        CRG_CUDA(cudaMemcpyAsync(tmp.data(), r->cand_pairs.p, sizeof(int2) * tmp.size(), cudaMemcpyDeviceToHost, r->stream));
        CRG_CUDA(cudaStreamSynchronize(r->stream));
    }
    for (size_t i = 0; i < tmp.size(); ++i) {
        if (src_idx) src_idx[i] = tmp[i].x;
        if (dst_idx) dst_idx[i] = tmp[i].y;
    }
    return CRG_OK;
}

int crg_normalize(crg_regridder *r) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    CRG_TRY(do_normalize(r));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_maximum(crg_regridder *r, double *out) {
    if (!r || !out) return set_error(CRG_ERR_INVALID, "null argument");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    CRG_TRY(device_maximum(r));
    CRG_CUDA(cudaMemcpyAsync(out, r->scratch_max.p, sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_scale(crg_regridder *r, double divisor) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    if (!(divisor > 0.0) || !std::isfinite(divisor)) return set_error(CRG_ERR_INVALID, "divisor must be positive and finite");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    CRG_TRY(r->scratch_max.alloc(1, r->stream));
    CRG_CUDA(cudaMemcpyAsync(r->scratch_max.p, &divisor, sizeof(double), cudaMemcpyHostToDevice, r->stream));
    CRG_TRY(divide_by_scratch(r));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_apply(crg_regridder *r, int32_t transpose, int32_t divide, double *dst, const double *src, int64_t K,
              int64_t ld_dst, int64_t ld_src, int32_t level_fastest) {
    return apply_impl(r, transpose, divide, dst, src, K, ld_dst, ld_src, level_fastest, true, true);
}

int crg_apply_async(crg_regridder *r, int32_t transpose, int32_t divide, double *dst, const double *src, int64_t K,
                    int64_t ld_dst, int64_t ld_src, int32_t level_fastest) {
    return apply_impl(r, transpose, divide, dst, src, K, ld_dst, ld_src, level_fastest, false, false);
}

int crg_mirror_fold_partners(double *field, int64_t nx, int64_t ny, int64_t K, int64_t ld, int32_t level_fastest,
                             int32_t device, void *stream) {
    if (!field) return set_error(CRG_ERR_INVALID, "crg_mirror_fold_partners: null field");
    if (nx < 4 || (nx & 3) || ny < 1 || K < 1) return set_error(CRG_ERR_INVALID, "crg_mirror_fold_partners: nx must be a positive multiple of 4, ny and K positive");
    const int64_t n = nx * ny;
    if (K == 1 && ld <= 0) ld = level_fastest ? 1 : n;
    if (level_fastest ? ld < K : ld < n) return set_error(CRG_ERR_INVALID, "crg_mirror_fold_partners: leading dimension %lld too small", (long long)ld);
    const int64_t nq = nx / 4, nh = nx / 2, base = (ny - 1) * nx;
    if (!is_device_ptr(field)) {
        for (int64_t k = 0; k < K; ++k)
            for (int64_t i = 0; i < 2 * nq; ++i) {
                const int64_t r = i < nq ? i : nh + (i - nq), q = nx - 1 - r;
                if (level_fastest) field[(base + q) * ld + k] = field[(base + r) * ld + k];
                else field[k * ld + base + q] = field[k * ld + base + r];
            }
        return CRG_OK;
    }
    CRG_TRY(check_device_available());
    DeviceGuard guard;
    CRG_TRY(guard.set(device));
    int dev = 0;
    CRG_CUDA(cudaGetDevice(&dev));
    cudaStream_t st = (cudaStream_t)stream;
    if (!st) CRG_TRY(device_stream(dev, &st));
    mirror_fold_kernel<<<ceil_div(2 * nq * K, 128), 128, 0, st>>>(field, nx, ny, K, ld, level_fastest);
    CRG_LAUNCH_CHECK();
    if (!stream) CRG_CUDA(cudaStreamSynchronize(st));
    return CRG_OK;
}

int crg_set_stream(crg_regridder *r, void *s) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    cudaStream_t ns = s ? (cudaStream_t)s : r->own_stream;
    if (ns == r->stream) return CRG_OK;
    // work already enqueued on the old stream stays ahead of what follows on the new one (no host wait)
    cudaEvent_t ev;
    CRG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, r->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ns, ev, 0);
    cudaEventDestroy(ev);
    CRG_CUDA(e);
    r->stream = ns;
    return CRG_OK;
}

int crg_synchronize(crg_regridder *r) {
    if (!r) return set_error(CRG_ERR_INVALID, "null regridder");
    DeviceGuard guard;
    CRG_TRY(guard.set(r->device));
    CRG_CUDA(cudaStreamSynchronize(r->stream));
    return CRG_OK;
}

int crg_apply_bytes(const crg_regridder *r, int32_t transpose, int32_t divide, int64_t K, int64_t *bytes) {
    if (!r || !bytes) return set_error(CRG_ERR_INVALID, "null argument");
    const int64_t n_out = transpose ? r->n_src : r->n_dst, n_in = transpose ? r->n_dst : r->n_src;
    *bytes = 12 * r->nnz + 4 * (n_out + 1) + (divide ? 8 * n_out : 0) + 8 * K * (n_in + n_out);
    return CRG_OK;
}

int crg_fp64_peak(int32_t device, double *tflops) {
    if (!tflops) return set_error(CRG_ERR_INVALID, "null argument");
    CRG_TRY(check_device_available());
    DeviceGuard guard;
    CRG_TRY(guard.set(device));
    cudaDeviceProp prop;
    int dev = 0;
    CRG_CUDA(cudaGetDevice(&dev));
    CRG_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double *out = nullptr;
    CRG_CUDA(cudaMalloc((void **)&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t a, b;
    CRG_CUDA(cudaEventCreate(&a));
    CRG_CUDA(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CRG_CUDA(cudaEventRecord(a));
        fp64_peak_kernel<<<blocks, threads>>>(out, iters);
        __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
        CRG_CUDA(cudaEventRecord(b));
        CRG_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        CRG_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    *tflops = best;
    return CRG_OK;
}

}  // extern "C"

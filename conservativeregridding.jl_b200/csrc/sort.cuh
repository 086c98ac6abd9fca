// sort.cuh -- hand-written device LSD radix sort of (uint64 key, 64-bit payload) pairs.
//
// Replaces the serial counting sort inside SparseArrays.sparse
// (/root/reference/src/regridder/intersection_areas.jl:115-121).  8-bit digits; per pass:
//   upsweep   : per-tile digit histogram                         (reads keys: 8 B/elem)
//   scan      : exclusive scan of the [256][ntiles] table         (scan.cuh)
//   downsweep : stable rank (warp match + per-warp counters), the tile is staged in shared memory in
//               digit order and written out in contiguous per-digit segments
//               (reads 16 B/elem, writes 16 B/elem)
// so one pass moves 40 B/element; HBM-bound.  The sort is stable, which the assembly uses
// to derive the CSC order from the CSR order with passes over the column bits only.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace crg {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per block
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_upsweep_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                                int shift, int ntiles,
                                                                uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[RS_RADIX];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_downsweep_kernel(
    const uint64_t *__restrict__ keys_in, const uint64_t *__restrict__ vals_in,
    uint64_t *__restrict__ keys_out, uint64_t *__restrict__ vals_out, int64_t n, int shift, int ntiles,
    const uint32_t *__restrict__ hist_scanned) {
    __shared__ uint32_t whist[RS_WARPS][RS_RADIX];
    __shared__ uint32_t gbase[RS_RADIX], dstart[RS_RADIX], scan_tmp[33];
    __shared__ uint64_t skey[RS_TILE], sval[RS_TILE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();

    // warp `wid` owns the contiguous chunk [base, base + 32*RS_ITEMS); item k of lane l is
    // element base + 32*k + l, so loads are coalesced and (warp, k, lane) is memory order.
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)wid * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + 32 * k + lane;
        key[k] = i < n ? keys_in[i] : ~0ull;
    }
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const bool valid = (base + 32 * k + lane) < n;
        const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & 0xFFu) : (0x100u | lane);
        const uint32_t peers = __match_any_sync(CRG_FULL, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = whist[wid][d];
            whist[wid][d] = old + __popc(peers);
        }
        old = __shfl_sync(CRG_FULL, old, leader);
        rank[k] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    uint32_t dcount;
    {   // digit `threadIdx.x`: exclusive scan of the per-warp counts + global base
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = whist[w][d]; whist[w][d] = run; run += c; }
        gbase[d] = hist_scanned[(size_t)d * ntiles + blockIdx.x];
        dcount = run;
    }
    // where the tile's run of digit d starts inside the tile (exclusive scan over the 256 digits)
    {
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<uint32_t>(dcount, scan_tmp, &total);
        dstart[threadIdx.x] = ex;
    }
    __syncthreads();
    // stage the tile in shared memory in digit order (stable), ...
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + 32 * k + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & 0xFFu;
            const uint32_t lp = dstart[d] + whist[wid][d] + rank[k];
            skey[lp] = key[k];
            sval[lp] = vals_in[i];
        }
    }
    __syncthreads();
    // ... then write it out: consecutive threads hold consecutive elements of a digit's run, so the
    // scatter goes out in contiguous segments (one per digit present in the tile)
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int tile_n = (int)min((int64_t)RS_TILE, n - tile_base);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int j = k * RS_THREADS + threadIdx.x;
        if (j < tile_n) {
            const uint64_t kk = skey[j];
            const uint32_t d = (uint32_t)(kk >> shift) & 0xFFu;
            const uint32_t pos = gbase[d] + ((uint32_t)j - dstart[d]);
            keys_out[pos] = kk;
            vals_out[pos] = sval[j];
        }
    }
}

// Sort n (key, val) pairs by the key bits [bit_lo, bit_hi) (stable).  Buffers a/b ping-pong;
// on return *result_in_b tells which pair holds the output.  n must be < 2^32.
inline int radix_sort_pairs(uint64_t *keys_a, uint64_t *vals_a, uint64_t *keys_b, uint64_t *vals_b, int64_t n,
                            int bit_lo, int bit_hi, bool *result_in_b, int *passes_done, cudaStream_t st) {
    *result_in_b = false;
    int passes = 0;
    if (n > 1 && bit_hi > bit_lo) {
        if (n >= ((int64_t)1 << 32)) return set_error(CRG_ERR_NOMEM, "radix sort: %lld elements exceed 2^32", (long long)n);
        const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
        DevBuf<uint32_t> hist;
        CRG_TRY(hist.alloc_tmp((size_t)RS_RADIX * ntiles + 1, st));
        uint64_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
        for (int shift = bit_lo; shift < bit_hi; shift += 8) {
            rs_upsweep_kernel<<<ntiles, RS_THREADS, 0, st>>>(ki, n, shift, ntiles, hist.p);
            CRG_LAUNCH_CHECK();
            CRG_TRY((exclusive_scan<uint32_t, uint32_t>(hist.p, (int64_t)RS_RADIX * ntiles, hist.p, st)));
            rs_downsweep_kernel<<<ntiles, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, ntiles, hist.p);
            CRG_LAUNCH_CHECK();
            uint64_t *t = ki; ki = ko; ko = t;
            t = vi; vi = vo; vo = t;
            ++passes;
        }
        *result_in_b = (passes & 1) != 0;
    }
    if (passes_done) *passes_done = passes;
    return CRG_OK;
}

}  // namespace crg

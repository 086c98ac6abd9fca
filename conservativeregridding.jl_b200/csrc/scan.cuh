// scan.cuh -- device-wide exclusive prefix sum (hand-written; no CUB/Thrust).
//
// out has n+1 entries: out[i] = sum_{j<i} in[j], out[n] = total.  Inputs of at most one 4096-element tile take
// one block; longer ones the single-pass decoupled-look-back kernel below.  HBM-bound: one read + one write.
#pragma once
#include "common.cuh"

namespace crg {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// tile_offsets == nullptr: single tile (n <= SCAN_TILE), writes out[n] = total itself.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const Tin *__restrict__ in, int64_t n,
                                                                 const Tout *__restrict__ tile_offsets,
                                                                 Tout *__restrict__ out) {
    __shared__ Tout sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    Tout v[SCAN_ITEMS];
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        v[k] = i < n ? (Tout)in[i] : Tout(0);
        s += v[k];
    }
    Tout total;
    Tout ex = block_exclusive_scan(s, sm, &total);
    Tout off = tile_offsets ? tile_offsets[blockIdx.x] : Tout(0);
    ex += off;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = off + total;
}

// Single pass over the input with decoupled look-back (Merrill & Garland): a tile takes a ticket, publishes the sum
// of its items, and its first warp walks back over its predecessors' (flag, value) words, 32 at a time -- aggregate or
// inclusive prefix, packed in one 64-bit word so that no fence is needed -- until it meets an inclusive prefix.  One launch and
// 2 x 8 B/item of traffic instead of three launches (reduce / scan of sums / downsweep) and a second read.
// status[0 .. ntiles): tile words, status[ntiles]: the ticket counter; zeroed before the launch.
constexpr unsigned long long SCAN_AGG = 1ull << 62, SCAN_INC = 2ull << 62, SCAN_VAL = (1ull << 62) - 1ull;

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) scan_lookback_kernel(const Tin *__restrict__ in, int64_t n, Tout *__restrict__ out,
                                                                    unsigned long long *__restrict__ status, int ntiles) {
    __shared__ Tout sm[33];
    __shared__ unsigned long long s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(status + ntiles, 1ull);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    Tout v[SCAN_ITEMS];
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? (Tout)in[i] : Tout(0);
        s += v[k];
    }
    Tout total;
    Tout ex = block_exclusive_scan(s, sm, &total);
    if (threadIdx.x < 32) {
        // look-back by the first warp, 32 predecessors at a time (one thread walking back tile by tile made the last
        // tiles of a 1 M-element scan wait for ~100 serial L2 round trips: 15-19 us per scan under ncu)
        const int lane = threadIdx.x;
        volatile unsigned long long *st = status;
        if (lane == 0 && tile > 0) st[tile] = SCAN_AGG | ((unsigned long long)total & SCAN_VAL);
        unsigned long long prefix = 0;
        int j = tile - 1;
        bool done = tile == 0;
        while (!done) {
            const int idx = j - lane;
            unsigned long long w = SCAN_INC;                       // before the first tile: an inclusive prefix of 0
            if (idx >= 0) { do { w = st[idx]; } while ((w >> 62) == 0ull); }
            const unsigned inc = __ballot_sync(CRG_FULL, (w >> 62) == 2ull);
            const int first = inc ? __ffs(inc) - 1 : 31;           // the nearest inclusive prefix ends the walk
            unsigned long long val = lane <= first ? (w & SCAN_VAL) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(CRG_FULL, val, o);
            prefix += val;
            if (inc) done = true; else j -= 32;
        }
        if (lane == 0) {
            st[tile] = SCAN_INC | ((prefix + (unsigned long long)total) & SCAN_VAL);
            s_prefix = prefix;
        }
    }
    __syncthreads();
    const Tout off = (Tout)s_prefix;
    ex += off;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) out[n] = off + total;
}

// Exclusive scan; `out` needs n+1 elements.  In-place (out == in, same type) is allowed.  Values are non-negative
// and below 2^62.
template <typename Tin, typename Tout>
int exclusive_scan(const Tin *in, int64_t n, Tout *out, cudaStream_t st) {
    if (n <= 0) {
        CRG_CUDA(cudaMemsetAsync(out, 0, sizeof(Tout), st));
        return CRG_OK;
    }
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 1) {
        scan_down_kernel<Tin, Tout><<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out);
        CRG_LAUNCH_CHECK();
        return CRG_OK;
    }
    DevBuf<unsigned long long> status;
    CRG_TRY(status.alloc_tmp((size_t)ntiles + 1, st));
    CRG_CUDA(cudaMemsetAsync(status.p, 0, sizeof(unsigned long long) * ((size_t)ntiles + 1), st));
    scan_lookback_kernel<Tin, Tout><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, out, status.p, (int)ntiles);
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

}  // namespace crg

// scan.cuh -- device-wide exclusive prefix sum (hand-written; no CUB/Thrust).
//
// Three-kernel reduce / scan-of-sums / downsweep with 4096-element tiles, recursing on the
// tile sums.  out has n+1 entries: out[i] = sum_{j<i} in[j], out[n] = total.  HBM-bound:
// 2 reads + 1 write of the input per level.
#pragma once
#include "common.cuh"

namespace crg {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const Tin *__restrict__ in, int64_t n,
                                                                   Tout *__restrict__ tile_sums) {
    __shared__ Tout sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += (Tout)in[i];
    }
    s = warp_sum(s);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (wid == 0) {
        Tout t = sm[lane];
        t = warp_sum(t);
        if (lane == 0) tile_sums[blockIdx.x] = t;
    }
}

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

// Exclusive scan; `out` needs n+1 elements.  In-place (out == in, same type) is allowed.
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
    DevBuf<Tout> sums;
    CRG_TRY(sums.alloc_tmp((size_t)ntiles + 1, st));
    scan_reduce_kernel<Tin, Tout><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, sums.p);
    CRG_LAUNCH_CHECK();
    CRG_TRY((exclusive_scan<Tout, Tout>(sums.p, ntiles, sums.p, st)));
    scan_down_kernel<Tin, Tout><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, sums.p, out);
    CRG_LAUNCH_CHECK();
    return CRG_OK;
}

}  // namespace crg

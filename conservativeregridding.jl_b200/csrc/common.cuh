// common.cuh -- error handling, stream-ordered device buffers, warp/block primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/crg_b200.h"

namespace crg {

// ---------------------------------------------------------------------------------------
// error plumbing: nothing throws across the ABI; the message is thread-local
// ---------------------------------------------------------------------------------------
extern thread_local char g_err[512];

int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e),
             what, file, line);
    if (e == cudaErrorMemoryAllocation) return CRG_ERR_NOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return CRG_ERR_NO_DEVICE;
    return CRG_ERR_CUDA;
}

#define CRG_CUDA(expr)                                                      \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return crg::fail_cuda(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define CRG_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != CRG_OK) return _rc; \
    } while (0)

extern unsigned long long g_launches;   // kernels launched by this library (crg_launch_count)
#define CRG_LAUNCH_CHECK()                    \
    do {                                      \
        __atomic_add_fetch(&crg::g_launches, 1ull, __ATOMIC_RELAXED); \
        CRG_CUDA(cudaGetLastError());         \
    } while (0)

// ---------------------------------------------------------------------------------------
// device memory
//   * persistent results: stream-ordered pool (cudaMallocAsync, release threshold = max);
//   * build temporaries: a grow-only per-device arena that is bump-allocated during one build and
//     reset at the start of the next -- no driver call on the hot path, no pool fragmentation
//     jitter (measured: per-phase times varied by milliseconds with pool allocations).  A request
//     that does not fit falls back to the pool and makes the arena grow before the next build.
// ---------------------------------------------------------------------------------------
struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0, want = 0;   // want = bytes requested during the current build
    void *take(size_t bytes) {
        const size_t a = (bytes + 255) & ~(size_t)255;
        want += a;
        if (!base || off + a > cap) return nullptr;
        void *p = base + off;
        off += a;
        return p;
    }
};
extern thread_local Arena *t_arena;      // set while a build runs on this thread

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    bool arena = false;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s), arena(o.arena) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; arena = o.arena; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    int alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        arena = false;
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), stream);
        if (e != cudaSuccess) { p = nullptr; n = 0; return fail_cuda(e, "cudaMallocAsync", __FILE__, __LINE__); }
        return CRG_OK;
    }
    // temporary that dies with the current build
    int alloc_tmp(size_t count, cudaStream_t stream) {
        if (t_arena) {
            void *q = t_arena->take((count ? count : 1) * sizeof(T));
            if (q) { release(); p = (T *)q; n = count; s = stream; arena = true; return CRG_OK; }
        }
        return alloc(count, stream);
    }
    void release() {
        if (p && !arena) cudaFreeAsync(p, s);
        p = nullptr; n = 0; arena = false;
    }
    size_t bytes() const { return n * sizeof(T); }
};

inline bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int ilog2_ceil(uint64_t n) { int b = 0; while (((uint64_t)1 << b) < n) ++b; return b; }

// ---------------------------------------------------------------------------------------
// warp / block primitives
// ---------------------------------------------------------------------------------------
#define CRG_FULL 0xffffffffu

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CRG_FULL, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(CRG_FULL, v, o); v = w > v ? w : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(CRG_FULL, v, o); v = w < v ? w : v; }
    return v;
}
// inclusive scan across a warp
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T w = __shfl_up_sync(CRG_FULL, v, o); if (lane >= o) v += w; }
    return v;
}

// Exclusive scan of one value per thread across the block (blockDim.x multiple of 32, <= 1024).
// `smem` must hold 33 T's.  Returns the exclusive prefix; *total receives the block sum.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *smem, T *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = warp_inclusive_scan(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < nw ? smem[lane] : T(0);
        T winc = warp_inclusive_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[wid] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ double atomic_max_pos_double(double *addr, double v) {
    // valid for non-negative doubles: IEEE ordering == unsigned integer ordering
    return __longlong_as_double((long long)atomicMax((unsigned long long *)addr,
                                                     (unsigned long long)__double_as_longlong(v)));
}
__device__ __forceinline__ void atomic_max_pos_float(float *addr, float v) {
    atomicMax((unsigned int *)addr, __float_as_uint(v));
}

}  // namespace crg

// Shared device/host utilities for the vican_b200 CUDA extension (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "mat3.cuh"

#define VB_CHECK(expr)                                  \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return -(int)_e;         \
    } while (0)
#define VB_KERNEL_CHECK() VB_CHECK(cudaGetLastError())

namespace vb {

constexpr unsigned FULL = 0xffffffffu;
// doubles per gathered node block (padded gather layout): 3 rows x (3 + 1 pad) = 96 bytes, padded to one
// 128-byte line so that the three 32-byte row loads of an edge never straddle two L1 lines
constexpr int GSTRIDE = 16;
constexpr int NUM_SMS_B200 = 148;

// Tally of this library's own kernel launches (vb:: kernels; CUB's are not counted), bumped by the
// host wrappers; bench.py reads it through vb_launch_count() around its timed region.
inline std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}
inline void count_launches(long long n) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

// Small pinned host buffer for scalar read-backs (a device-to-host copy into pageable memory is staged by the
// driver and costs tens of microseconds more per synchronisation than one into pinned memory)
inline long long* pinned_scalars() {
    static thread_local long long* p = nullptr;
    if (p == nullptr && cudaMallocHost((void**)&p, 128 * sizeof(long long)) != cudaSuccess) p = nullptr;
    return p;
}
// copy n bytes (<= 1 KB) from the device to *dst through the pinned buffer and synchronise the stream
inline cudaError_t read_back(void* dst, const void* d_src, size_t n, cudaStream_t st) {
    long long* h = pinned_scalars();
    if (h == nullptr || n > 128 * sizeof(long long)) {
        cudaError_t e = cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    cudaError_t e = cudaMemcpyAsync(h, d_src, n, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    memcpy(dst, h, n);
    return cudaSuccess;
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = NUM_SMS_B200;
    }
    return n;
}

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl_down(double v, int d) { return __shfl_down_sync(FULL, v, d); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(FULL, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// L2 cache policies (createpolicy) for the two classes of traffic in an edge pass:
//  * edge data (blocks, indices) is read exactly once per pass  -> evict-first, no L1 allocation
//  * gathered node blocks (X, W) are re-read by many edges       -> evict-last
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ld_stream(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_keep(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

}  // namespace vb

// Shared helpers for the seggroup_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/seggroup_b200.h"

#define SGB_CHECK_LAUNCH()                                         \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return sgb_cuda_error((int)e__);   \
    } while (0)

#define SGB_CUDA(call)                                             \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return sgb_cuda_error((int)e__);   \
    } while (0)

int sgb_cuda_error(int cuda_code);   // records the code, returns SGB_ERR_CUDA
void sgb_count_launch();             // one kernel of this library was launched (sgb_launch_count)
#define SGB_COUNT_LAUNCH() sgb_count_launch()

static inline int sgb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
#include <atomic>
// Opt a kernel in to the full 227 KB of dynamic shared memory ONCE per (kernel, device).  The attribute is process-wide
// state of the function: setting it to the size of each launch is a race between scene threads (one thread lowers it
// between another thread's set and launch -> "invalid argument"), so every kernel gets the device maximum, set once.
template <auto Kernel>
inline cudaError_t sgb_opt_in_smem() {
    static std::atomic<unsigned> done{0};            // bit per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned bit = 1u << (dev & 31);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, Kernel);
    if (e != cudaSuccess) return e;
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);   // static + dynamic <= opt-in
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
#define SGB_OPT_IN_SMEM(...) SGB_CUDA((sgb_opt_in_smem<__VA_ARGS__>()))
#endif

#ifdef __CUDACC__
#define SGB_FULL_MASK 0xffffffffu

// Monotone map float -> uint32 (a < b  <=>  key(a) < key(b); +NaN sorts above +inf).
// The sign-bit OR is written in PTX: as C, the compiler folds `bits | 0x80000000` on a float into -|f| (FADD with
// modifiers), which canonicalises a NaN to 0x7fffffff and drops it to the BOTTOM of the key order.
__device__ __forceinline__ uint32_t sgb_float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    uint32_t pos;
    asm("or.b32 %0, %1, 0x80000000;" : "=r"(pos) : "r"(u));
    return (u & 0x80000000u) ? ~u : pos;
}
__device__ __forceinline__ float sgb_key_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ float sgb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double sgb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int sgb_warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGB_FULL_MASK, v, o);
    return v;
}
// largest i with offsets[i] <= pos, offsets ascending, offsets[0] = 0, n = number of segments
__device__ __forceinline__ int sgb_upper_segment(const int* __restrict__ offsets, int n, int pos) {
    int lo = 0, hi = n;       // invariant: offsets[lo] <= pos < offsets[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}
#endif

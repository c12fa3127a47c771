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
#define SGB_FULL_MASK 0xffffffffu

// Monotone map float -> uint32 (a < b  <=>  key(a) < key(b); +NaN sorts above +inf).
__device__ __forceinline__ uint32_t sgb_float_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
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

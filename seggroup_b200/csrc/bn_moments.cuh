// BatchNorm (batch statistics) from accumulated first/second moments of the layer INPUT.
#pragma once
#include "common.cuh"

namespace sgb_bn {
constexpr float SLOPE = 0.2f;
constexpr float BN_EPS = 1e-5f;
constexpr int COUT = 64;

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * SLOPE; }

// out[i] = sum_b part[b*n + i] in a fixed order (row groups b = y, y+8, ... summed per thread, then groups 0..7):
// the per-CTA partial sums of the moment / gradient kernels are reduced by n/32 CTAs instead of one thread per entry.
template <typename T>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const T* __restrict__ part, int nb, int n, double* __restrict__ out) {
    __shared__ double s_p[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + x;
    double s = 0;
    if (i < n) for (int b = y; b < nb; b += 8) s += (double)part[(size_t)b * n + i];
    s_p[y][x] = s;
    __syncthreads();
    if (y == 0 && i < n) {
        double t = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += s_p[g][x];
        out[i] = t;
    }
}
template <typename T>
inline void reduce_partials(const T* part, int nb, int n, double* out, cudaStream_t st) {
    { reduce_partials_kernel<T><<<(n + 31) / 32, 256, 0, st>>>(part, nb, n, out); SGB_COUNT_LAUNCH(); }
}

// BN1 statistics from the moments: y = W e = W e' + W e0 (e0 = centre the moments were taken about).
// stats layout [4][64]: mean, invstd, scale = gamma*invstd, beta ; var_out[64] = biased variance.
// moments_out[NE1] = reduced (s', G') in fp64 for the backward pass.
template <int CIN>
__global__ void __launch_bounds__(64)
bn1_finalize_kernel(const double* __restrict__ part, int nb, double M, const float* __restrict__ W1, const float* __restrict__ e0v,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ stats,
                    float* __restrict__ var_out, double* __restrict__ moments_out) {
    constexpr int NE1 = CIN * (CIN + 1) / 2 + CIN;
    __shared__ double s_m[NE1];
    __shared__ double s_cov[CIN][CIN];
    for (int n = threadIdx.x; n < NE1; n += blockDim.x) {
        double s = 0;
        for (int b = 0; b < nb; ++b) s += part[(size_t)b * NE1 + n];
        s_m[n] = s;
        if (moments_out) moments_out[n] = s;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < CIN * (CIN + 1) / 2; n += blockDim.x) {
        int q = n, t = 0;
        while (q >= CIN - t) { q -= CIN - t; ++t; }
        const int u = t + q;
        const double c = s_m[CIN + n] / M - (s_m[t] / M) * (s_m[u] / M);
        s_cov[t][u] = c; s_cov[u][t] = c;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < COUT) {
        double mean = 0, var = 0;
        for (int t = 0; t < CIN; ++t) {
            const double w = (double)W1[c * CIN + t];
            const double e0 = e0v ? (double)e0v[t] : 0.0;
            mean += w * (s_m[t] / M + e0);
            double r = 0;
            for (int u = 0; u < CIN; ++u) r += s_cov[t][u] * (double)W1[c * CIN + u];
            var += w * r;
        }
        if (var < 0) var = 0;
        const double invstd = 1.0 / sqrt(var + (double)BN_EPS);
        stats[c] = (float)mean;
        stats[64 + c] = (float)invstd;
        stats[128 + c] = (float)((double)gamma[c] * invstd);
        stats[192 + c] = beta[c];
        if (var_out) var_out[c] = (float)var;
    }
}

}  // namespace sgb_bn

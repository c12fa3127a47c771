// a9 backward: parameter gradients of MLP2 / MLP3 fused with the point -> segment max pooling.
// (The reference gets these from autograd over materialised [1,64,N,20] tensors; inputs have no grad.)
//
// Upstream gradient g[s,c] reaches exactly one edge activation per (segment, channel): point
// p* = argmax point of the pooling, edge k* = argmax edge of that point (argk, kept by the forward).
// Training-mode BatchNorm couples every edge through the batch statistics:
//     dy = gamma*invstd * (dv - mean(dv) - zhat * mean(dv*zhat))
// so dW = sum_e dy_e x_e^T has a SPARSE part (the S*64 winning edges) and a DENSE part that is a linear
// function of the layer-input moments the forward already accumulated:
//     dW[c,:] = gamma_c invstd_c [ sum_sparse dv (x - xbar)  -  dgamma_c invstd_c (Cov_x W_c) ]
// For MLP3 the gradient also flows into the hidden layer,
//     dh_e = sparse_e - q0/M - Bm (h_e - hbar),   Bm = W2^T diag(gamma2 invstd2^2 dgamma2) W2 / M
// whose dense part needs one pass over all edges (recompute h, 64x64 mat-vec, LeakyReLU mask, accumulate
// sum dv1, sum dv1*zhat1, sum dv1 (e - ebar)^T).  Order of kernels:
//     sparse -> [mid finalize -> dense (tensor cores, edgeconv_bwd_tc.cu)] -> last finalize.
// All cross-warp / cross-block sums run in a fixed order (deterministic).
#include "edgeconv_common.cuh"

namespace sgb_ec {
constexpr int NACC = CIN + 2;          // per hidden channel: dbeta, dgamma, T[18]
constexpr int SEG_CHUNK = 32;          // segments per sparse-kernel block

// ebar[t] = mean edge feature = s'/M + e0
__device__ __forceinline__ void load_ebar(const double* __restrict__ mom1, const float* __restrict__ e0, double M, float* s_ebar) {
    if (threadIdx.x < CIN) s_ebar[threadIdx.x] = (float)(mom1[threadIdx.x] / M + (double)e0[threadIdx.x]);
}

template <bool TWO>
__global__ void __launch_bounds__(WARPS * 32)
bwd_sparse_kernel(const float* __restrict__ g /*[S,64]*/, const int* __restrict__ arg /*[S,64] point ids*/,
                  const unsigned char* __restrict__ argk /*[N,64]*/, int S,
                  const float* __restrict__ x9, const int* __restrict__ knn, const float* __restrict__ W1,
                  const float* __restrict__ stats1, const double* __restrict__ mom1, const float* __restrict__ e0, double M,
                  const float* __restrict__ W2, const float* __restrict__ stats2, const double* __restrict__ mom2,
                  float* __restrict__ part1 /*[gridDim.y*gridDim.x][64*NACC]*/, float* __restrict__ part2 /*[gridDim.x][64][66]*/) {
    __shared__ float s_w1t[CIN][COUT];
    __shared__ float s_ebar[CINP];
    __shared__ float s_red[WARPS][COUT * NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * WARPS + warp;          // pooled-output channel of this warp
    const int c0 = lane * 2;                          // hidden / layer-1 channels of this lane
    for (int i = threadIdx.x; i < COUT * CIN; i += blockDim.x) s_w1t[i % CIN][i / CIN] = __ldg(W1 + i);
    load_ebar(mom1, e0, M, s_ebar);
    __syncthreads();
    const float mu0 = stats1[c0], mu1 = stats1[c0 + 1], is0 = stats1[64 + c0], is1 = stats1[64 + c0 + 1];
    const float ga0 = stats1[128 + c0], ga1 = stats1[128 + c0 + 1], be0 = stats1[192 + c0], be1 = stats1[192 + c0 + 1];
    float w2a = 0.f, w2b = 0.f, mu2 = 0.f, is2 = 0.f, ga2 = 0.f, be2 = 0.f, hb0 = 0.f, hb1 = 0.f;
    if (TWO) {
        w2a = __ldg(W2 + c * COUT + c0); w2b = __ldg(W2 + c * COUT + c0 + 1);
        mu2 = stats2[c]; is2 = stats2[64 + c]; ga2 = stats2[128 + c]; be2 = stats2[192 + c];
        hb0 = (float)(mom2[COUT * COUT + c0] / M); hb1 = (float)(mom2[COUT * COUT + c0 + 1] / M);
    }
    float acc[2][NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
    float db2 = 0.f, dg2 = 0.f, t2a = 0.f, t2b = 0.f;

    const int s_begin = blockIdx.x * SEG_CHUNK, s_end = min(S, s_begin + SEG_CHUNK);
    for (int s = s_begin; s < s_end; ++s) {
        const float gv = __ldg(g + (size_t)s * COUT + c);
        const int p = __ldg(arg + (size_t)s * COUT + c);
        const int k = argk[(size_t)p * COUT + c];
        const int j = __ldg(knn + (size_t)p * KNN + k);
        float e[CIN];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float xi = __ldg(x9 + (size_t)p * 9 + t);
            e[t] = __ldg(x9 + (size_t)j * 9 + t) - xi;
            e[9 + t] = xi;
        }
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int t = 0; t < CIN; ++t) {
            const float2 w = *reinterpret_cast<const float2*>(&s_w1t[t][c0]);
            y0 = fmaf(w.x, e[t], y0); y1 = fmaf(w.y, e[t], y1);
        }
        const float z0 = (y0 - mu0) * is0, z1 = (y1 - mu1) * is1;          // zhat1
        const float v0 = fmaf(y0 - mu0, ga0, be0), v1 = fmaf(y1 - mu1, ga1, be1);
        float dv0, dv1;
        if (TWO) {
            const float h0 = lrelu(v0), h1 = lrelu(v1);
            const float y2 = sgb_warp_sum(fmaf(w2a, h0, w2b * h1));
            const float zz = (y2 - mu2) * is2;
            const float v2 = fmaf(y2 - mu2, ga2, be2);
            const float dv2 = gv * (v2 > 0.f ? 1.f : SLOPE);
            db2 += dv2; dg2 = fmaf(dv2, zz, dg2);
            t2a = fmaf(dv2, h0 - hb0, t2a); t2b = fmaf(dv2, h1 - hb1, t2b);
            const float dy2 = ga2 * dv2;                                    // gamma2*invstd2*dv2 (stats2[128+c] = gamma*invstd)
            dv0 = w2a * dy2 * (v0 > 0.f ? 1.f : SLOPE);
            dv1 = w2b * dy2 * (v1 > 0.f ? 1.f : SLOPE);
        } else {
            dv0 = (c0 == c) ? gv * (v0 > 0.f ? 1.f : SLOPE) : 0.f;
            dv1 = (c0 + 1 == c) ? gv * (v1 > 0.f ? 1.f : SLOPE) : 0.f;
        }
        acc[0][0] += dv0; acc[1][0] += dv1;
        acc[0][1] = fmaf(dv0, z0, acc[0][1]); acc[1][1] = fmaf(dv1, z1, acc[1][1]);
#pragma unroll
        for (int t = 0; t < CIN; ++t) {
            const float d = e[t] - s_ebar[t];
            acc[0][2 + t] = fmaf(dv0, d, acc[0][2 + t]);
            acc[1][2 + t] = fmaf(dv1, d, acc[1][2 + t]);
        }
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) { s_red[warp][c0 * NACC + i] = acc[0][i]; s_red[warp][(c0 + 1) * NACC + i] = acc[1][i]; }
    __syncthreads();
    float* dst = part1 + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (COUT * NACC);
    for (int i = threadIdx.x; i < COUT * NACC; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) sum += s_red[w][i];
        dst[i] = sum;
    }
    if (TWO) {
        float* d2 = part2 + ((size_t)blockIdx.x * COUT + c) * 66;
        d2[c0] = t2a; d2[c0 + 1] = t2b;
        if (lane == 0) { d2[64] = db2; d2[65] = dg2; }
    }
}

// MLP3: from the reduced layer-2 sums t2 [64][66] ([c][j] = sum dv2 (h_j - hbar_j), [c][64] = dbeta2, [c][65] = dgamma2)
// emit dW2 / dgamma2 / dbeta2 and the dense-pass coefficients Bm [64][64], r [64].  One CTA per row (64 CTAs x 64 threads).
__global__ void __launch_bounds__(64)
bwd_mid_kernel(const double* __restrict__ t2, double M, const float* __restrict__ W2, const float* __restrict__ stats2,
               const double* __restrict__ mom2, float* __restrict__ gW2, float* __restrict__ gg2, float* __restrict__ gb2,
               float* __restrict__ coef /*[64*64 + 64]*/) {
    __shared__ float s_w2[COUT][COUT + 1];
    __shared__ double s_hbar[COUT];
    __shared__ double s_d[COUT];                // D_c = gamma2 invstd2^2 dgamma2
    __shared__ double s_q[COUT];                // gamma2 invstd2 dbeta2
    __shared__ double s_red[COUT];
    const int b = blockIdx.x, j = threadIdx.x;
    for (int i = j; i < COUT * COUT; i += COUT) s_w2[i / COUT][i % COUT] = __ldg(W2 + i);
    s_hbar[j] = mom2[COUT * COUT + j] / M;
    {
        const double gi = (double)stats2[128 + j];                // gamma*invstd
        s_d[j] = gi * (double)stats2[64 + j] * t2[j * 66 + 65];
        s_q[j] = gi * t2[j * 66 + 64];
        if (b == 0) { gb2[j] = (float)t2[j * 66 + 64]; gg2[j] = (float)t2[j * 66 + 65]; }
    }
    __syncthreads();
    // Bm[b][j] = (1/M) sum_c W2[c][b] D_c W2[c][j]
    double bm = 0;
    for (int c = 0; c < COUT; ++c) bm += (double)s_w2[c][b] * s_d[c] * (double)s_w2[c][j];
    const float bmf = (float)(bm / M);
    coef[b * COUT + j] = bmf;
    // dW2[c=b][j] = gamma2 invstd2 [ T2[c][j] - dgamma2 invstd2 sum_i CovH[j][i] W2[c][i] ]
    {
        const int c = b;
        double r = 0;
        for (int i = 0; i < COUT; ++i) {
            const double cov = 0.5 * (mom2[j * COUT + i] + mom2[i * COUT + j]) / M - s_hbar[j] * s_hbar[i];
            r += cov * (double)s_w2[c][i];
        }
        gW2[c * COUT + j] = (float)((double)stats2[128 + c] * (t2[c * 66 + j] - t2[c * 66 + 65] * (double)stats2[64 + c] * r));
    }
    // r[b] = -q0'[b]/M + (Bm hbar)[b],  q0'[b] = sum_c W2[c][b] gamma2 invstd2 dbeta2   (Bm as rounded to fp32 for the dense pass)
    s_red[j] = (double)bmf * s_hbar[j];
    __syncthreads();
    if (j == 0) {
        double q0 = 0, bh = 0;
        for (int c = 0; c < COUT; ++c) q0 += (double)s_w2[c][b] * s_q[c];
        for (int i = 0; i < COUT; ++i) bh += s_red[i];
        coef[COUT * COUT + b] = (float)(-q0 / M + bh);
    }
}

// x9 [N,9] -> x12 [N,12] (48-byte rows: three aligned 16-byte loads per gathered row)
__global__ void pad12_kernel(const float* __restrict__ x9, long long n12, float* __restrict__ x12) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n12) return;
    const long long p = i / 12;
    const int q = (int)(i % 12);
    x12[i] = q < 9 ? __ldg(x9 + p * 9 + q) : (q == 11 ? 1.f : 0.f);     // pad lane 11 = 1: "this row was gathered" flag of the tcgen05 producers (a zero-filled row has 0 there)
}

// dW1 / dgamma1 / dbeta1 from the reduced (sparse [+ dense]) sums red [nred][64*NACC]
__global__ void __launch_bounds__(64)
bwd_last_kernel(const double* __restrict__ red, int nred, double M,
                const float* __restrict__ W1, const float* __restrict__ stats1, const double* __restrict__ mom1,
                float* __restrict__ gW1, float* __restrict__ gg1, float* __restrict__ gb1) {
    __shared__ double s_cov[CIN][CIN];
    for (int n = threadIdx.x; n < CIN * (CIN + 1) / 2; n += blockDim.x) {
        int q = n, t = 0;
        while (q >= CIN - t) { q -= CIN - t; ++t; }
        const int u = t + q;
        const double cv = mom1[CIN + n] / M - (mom1[t] / M) * (mom1[u] / M);
        s_cov[t][u] = cv; s_cov[u][t] = cv;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= COUT) return;
    double a[NACC];
    for (int i = 0; i < NACC; ++i) {
        double s = 0;
        for (int b = 0; b < nred; ++b) s += red[(size_t)b * (COUT * NACC) + c * NACC + i];
        a[i] = s;
    }
    gb1[c] = (float)a[0];
    gg1[c] = (float)a[1];
    const double gi = (double)stats1[128 + c], is = (double)stats1[64 + c];
    for (int t = 0; t < CIN; ++t) {
        double r = 0;
        for (int u = 0; u < CIN; ++u) r += s_cov[t][u] * (double)W1[c * CIN + u];
        gW1[c * CIN + t] = (float)(gi * (a[2 + t] - a[1] * is * r));
    }
}

}  // namespace sgb_ec

using namespace sgb_ec;

extern "C" size_t sgb_edgeconv_bwd_ws_bytes(int N, int S, int two_layer) {
    const size_t nchunk = (size_t)sgb_div_up(S, SEG_CHUNK);
    size_t b = nchunk * 8 * COUT * NACC * sizeof(float);
    if (two_layer) b += nchunk * COUT * 66 * sizeof(float) + (COUT * COUT + COUT) * sizeof(float);
    b += (size_t)(2 * COUT * NACC + COUT * 66) * sizeof(double);       // reduced sums
    b = (b + 255) & ~(size_t)255;
    if (two_layer) b += ((sgb_ec2_bwd_tc_part_bytes(N) + 255) & ~(size_t)255) + (size_t)(N > 0 ? N : 0) * 48;   // dense partials (fp64), x12
    return b + 256;
}

// g [S,64] gradient of the pooled features; arg [S,64] argmax point ids (sgb_segment_pool_max_fwd);
// argk [N,64] argmax edge per point/channel (sgb_edgeconv_fwd); stats/mom/e0 as returned by the forward.
// Outputs: gW1 [64,18], gg1/gb1 [64] (+ gW2 [64,64], gg2/gb2 [64] when two_layer).
extern "C" int sgb_edgeconv_bwd(const float* g, const int* arg, const unsigned char* argk, int S, const float* x9, const int* knn, int N,
                                int two_layer, const float* W1, const float* stats1, const double* mom1, const float* e0,
                                const float* W2, const float* stats2, const double* mom2,
                                float* gW1, float* gg1, float* gb1, float* gW2, float* gg2, float* gb2,
                                void* ws, size_t ws_bytes, void* stream) {
    if (S <= 0 || N <= 0) return SGB_ERR_INVALID;
    if (!g || !arg || !argk || !x9 || !knn || !W1 || !stats1 || !mom1 || !e0 || !gW1 || !gg1 || !gb1 || !ws) return SGB_ERR_INVALID;
    if (two_layer && (!W2 || !stats2 || !mom2 || !gW2 || !gg2 || !gb2)) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_edgeconv_bwd_ws_bytes(N, S, two_layer)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunk = sgb_div_up(S, SEG_CHUNK);
    const double M = (double)N * KNN;
    double* red = (double*)ws;                               // [2][64*NACC] sparse / dense sums, then t2 [64*66]
    double* t2 = red + 2 * COUT * NACC;
    float* part1 = (float*)(t2 + COUT * 66);
    float* part2 = part1 + (size_t)nchunk * 8 * COUT * NACC;
    dim3 grid(nchunk, COUT / WARPS);
    if (!two_layer) {
        { bwd_sparse_kernel<false><<<grid, WARPS * 32, 0, st>>>(g, arg, argk, S, x9, knn, W1, stats1, mom1, e0, M, nullptr, nullptr, nullptr,
                                                             part1, nullptr); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(part1, nchunk * 8, COUT * NACC, red, st);
        { bwd_last_kernel<<<1, 64, 0, st>>>(red, 1, M, W1, stats1, mom1, gW1, gg1, gb1); SGB_COUNT_LAUNCH(); }
    } else {
        float* coef = part2 + (size_t)nchunk * COUT * 66;
        double* partD = (double*)(((uintptr_t)(coef + COUT * COUT + COUT) + 255) & ~(uintptr_t)255);
        float* x12 = (float*)((unsigned char*)partD + ((sgb_ec2_bwd_tc_part_bytes(N) + 255) & ~(size_t)255));
        int gd = 0;
        { bwd_sparse_kernel<true><<<grid, WARPS * 32, 0, st>>>(g, arg, argk, S, x9, knn, W1, stats1, mom1, e0, M, W2, stats2, mom2, part1, part2); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(part1, nchunk * 8, COUT * NACC, red, st);
        sgb_bn::reduce_partials(part2, nchunk, COUT * 66, t2, st);
        { bwd_mid_kernel<<<COUT, 64, 0, st>>>(t2, M, W2, stats2, mom2, gW2, gg2, gb2, coef); SGB_COUNT_LAUNCH(); }
        // dense pass over all N*20 edges: the 64x64 mat-vec per edge on the tcgen05 tensor cores (edgeconv_bwd_tc.cu)
        { pad12_kernel<<<sgb_div_up((long long)N * 12, 256), 256, 0, st>>>(x9, (long long)N * 12, x12); SGB_COUNT_LAUNCH(); }
        int rc = sgb_ec2_bwd_tc_dense(x12, knn, N, W1, stats1, mom1, e0, M, coef, partD, &gd, st);
        if (rc) return rc;
        sgb_bn::reduce_partials(partD, gd, COUT * NACC, red + COUT * NACC, st);
        { bwd_last_kernel<<<1, 64, 0, st>>>(red, 2, M, W1, stats1, mom1, gW1, gg1, gb1); SGB_COUNT_LAUNCH(); }
    }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

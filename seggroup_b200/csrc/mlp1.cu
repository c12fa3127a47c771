// a5/a6: structural-layer point MLP on the 64-point segment clouds (seggroup/model.py:39-80).
//   idx   = knn(xyz, k=10) inside each cloud (exact torch-CPU fp32 score expression, see cluster_knn.cu)
//   e_k   = ((xyz_nb - mean_k xyz_nb) * 10, rgb_nb)            (get_graph_feature1, model.py:56-58)
//   m     = max_k lrelu(BN(W e_k))    [S,64ch,64pts]           (BN over the S*64*10 edge activations)
//   Feat  = cat(max_pts m, mean_pts m) [S,128]
// One CTA (64 threads) per cloud, one thread per point; the cloud lives in shared memory.
// Kernel 1 finds the neighbours and accumulates the 6x6 input moments (-> BN statistics, analytic
// backward); kernel 2 recomputes the edges and reduces over k and over the points.
#include "common.cuh"
#include "bn_moments.cuh"

namespace sgb_mlp1 {
constexpr int P = 64;       // points per cloud
constexpr int K = 10;
constexpr int CIN = 6;
constexpr int COUT = 64;
constexpr int NE = CIN * (CIN + 1) / 2 + CIN;   // 27
using sgb_bn::lrelu;

__device__ __forceinline__ float sq_norm_ref(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// e[k][0..5] of point `pt` from the cloud in smem and its neighbour list
__device__ __forceinline__ void build_edges(const float (*cl)[CIN], const int* nb, float (&e)[K][CIN]) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { sx += cl[nb[k]][0]; sy += cl[nb[k]][1]; sz += cl[nb[k]][2]; }
    const float mx = sx / (float)K, my = sy / (float)K, mz = sz / (float)K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float* r = cl[nb[k]];
        e[k][0] = (r[0] - mx) * 10.f; e[k][1] = (r[1] - my) * 10.f; e[k][2] = (r[2] - mz) * 10.f;
        e[k][3] = r[3]; e[k][4] = r[4]; e[k][5] = r[5];
    }
}

__global__ void __launch_bounds__(P)
knn_gram_kernel(const float* __restrict__ clouds, int S, int* __restrict__ idx_out /*[S,64,10]*/,
                double* __restrict__ part /*[S][NE]*/) {
    __shared__ float cl[P][CIN];
    __shared__ float4 s_q[P];
    __shared__ float s_red[2][NE];
    const int s = blockIdx.x, pt = threadIdx.x;
    const float* src = clouds + (size_t)s * P * CIN;
    for (int i = pt; i < P * CIN; i += P) (&cl[0][0])[i] = __ldg(src + i);
    __syncthreads();
    const float xi = cl[pt][0], yi = cl[pt][1], zi = cl[pt][2];
    const float xxi = sq_norm_ref(xi, yi, zi);
    s_q[pt] = make_float4(xi, yi, zi, xxi);
    __syncthreads();
    float sc[K]; int id[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { sc[i] = -INFINITY; id[i] = 0; }
    int filled = 0;
    for (int j = 0; j < P; ++j) {
        const float4 c = s_q[j];
        float m = __fmul_rn(xi, c.x);
        m = __fmaf_rn(yi, c.y, m);
        m = __fmaf_rn(zi, c.z, m);
        const float sco = __fsub_rn(__fsub_rn(-c.w, __fmul_rn(-2.f, m)), xxi);
        if (filled < K || sco > sc[K - 1]) {
            if (filled < K) ++filled;
            sc[K - 1] = sco; id[K - 1] = j;
#pragma unroll
            for (int i = K - 1; i > 0; --i) {
                if (sc[i] > sc[i - 1]) {
                    const float ts = sc[i]; sc[i] = sc[i - 1]; sc[i - 1] = ts;
                    const int ti = id[i]; id[i] = id[i - 1]; id[i - 1] = ti;
                }
            }
        }
    }
    int* o = idx_out + ((size_t)s * P + pt) * K;
#pragma unroll
    for (int i = 0; i < K; ++i) o[i] = id[i];

    float e[K][CIN];
    build_edges(cl, id, e);
    // 27 moment entries per thread over its 10 edges, then a block sum (warp shuffles, 2 warps)
    int n = 0;
#pragma unroll
    for (int t = 0; t < CIN; ++t) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) a += e[k][t];
        a = sgb_warp_sum(a);
        if ((pt & 31) == 0) s_red[pt >> 5][n] = a;
        ++n;
    }
#pragma unroll
    for (int t = 0; t < CIN; ++t) {
#pragma unroll
        for (int u = t; u < CIN; ++u) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) a = fmaf(e[k][t], e[k][u], a);
            a = sgb_warp_sum(a);
            if ((pt & 31) == 0) s_red[pt >> 5][n] = a;
            ++n;
        }
    }
    __syncthreads();
    if (pt < NE) part[(size_t)s * NE + pt] = (double)s_red[0][pt] + (double)s_red[1][pt];
}

// fwd: Feat[s, c] = max_pts m, Feat[s, 64 + c] = mean_pts m; arg_pt[s, c] = first point attaining the max
__global__ void __launch_bounds__(P)
forward_kernel(const float* __restrict__ clouds, const int* __restrict__ idx, int S, const float* __restrict__ W,
               const float* __restrict__ stats, float* __restrict__ feat /*[S,128]*/, int* __restrict__ arg_pt /*[S,64] or null*/) {
    __shared__ float cl[P][CIN];
    __shared__ float s_w[COUT][CIN + 2];      // + mean, scale  (beta in s_b)
    __shared__ float s_b[COUT];
    __shared__ float s_m[P][COUT + 1];
    const int s = blockIdx.x, pt = threadIdx.x;
    const float* src = clouds + (size_t)s * P * CIN;
    for (int i = pt; i < P * CIN; i += P) (&cl[0][0])[i] = __ldg(src + i);
    {
        const int c = pt;
#pragma unroll
        for (int t = 0; t < CIN; ++t) s_w[c][t] = __ldg(W + c * CIN + t);
        s_w[c][CIN] = stats[c]; s_w[c][CIN + 1] = stats[128 + c]; s_b[c] = stats[192 + c];
    }
    __syncthreads();
    int nb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) nb[k] = __ldg(idx + ((size_t)s * P + pt) * K + k);
    float e[K][CIN];
    build_edges(cl, nb, e);
    for (int c = 0; c < COUT; ++c) {
        const float mean = s_w[c][CIN], scale = s_w[c][CIN + 1], beta = s_b[c];
        float best = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float y = 0.f;
#pragma unroll
            for (int t = 0; t < CIN; ++t) y = fmaf(s_w[c][t], e[k][t], y);
            best = fmaxf(best, lrelu(fmaf(y - mean, scale, beta)));
        }
        s_m[pt][c] = best;
    }
    __syncthreads();
    {
        const int c = pt;
        float best = s_m[0][c], sum = s_m[0][c];
        int bi = 0;
        for (int q = 1; q < P; ++q) {
            const float v = s_m[q][c];
            sum += v;
            if (v > best) { best = v; bi = q; }
        }
        feat[(size_t)s * 2 * COUT + c] = best;
        feat[(size_t)s * 2 * COUT + COUT + c] = sum / (float)P;
        if (arg_pt) arg_pt[(size_t)s * COUT + c] = bi;
    }
}

// backward: parameter gradients (the clouds are inputs without gradient).
// g [S,128]: g[:, :64] flows to the arg-max point of every channel, g[:, 64:] / 64 to every point; each point then
// routes to its arg-max edge.  Per channel we need sum dv, sum dv*zhat, sum dv*(e - ebar) (see edgeconv_bwd.cu).
__global__ void __launch_bounds__(P)
backward_kernel(const float* __restrict__ g, const float* __restrict__ clouds, const int* __restrict__ idx, const int* __restrict__ arg_pt,
                int S, const float* __restrict__ W, const float* __restrict__ stats, const double* __restrict__ mom, double M,
                float* __restrict__ part /*[S][64][8]*/) {
    __shared__ float cl[P][CIN];
    __shared__ float s_w[COUT][CIN + 2];
    __shared__ float s_b[COUT];
    __shared__ float s_ebar[CIN];
    __shared__ float s_red[2][COUT][8];
    const int s = blockIdx.x, pt = threadIdx.x;
    const float* src = clouds + (size_t)s * P * CIN;
    for (int i = pt; i < P * CIN; i += P) (&cl[0][0])[i] = __ldg(src + i);
    {
        const int c = pt;
#pragma unroll
        for (int t = 0; t < CIN; ++t) s_w[c][t] = __ldg(W + c * CIN + t);
        s_w[c][CIN] = stats[c]; s_w[c][CIN + 1] = stats[128 + c]; s_b[c] = stats[192 + c];
    }
    if (pt < CIN) s_ebar[pt] = (float)(mom[pt] / M);
    __syncthreads();
    int nb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) nb[k] = __ldg(idx + ((size_t)s * P + pt) * K + k);
    float e[K][CIN];
    build_edges(cl, nb, e);
    for (int c = 0; c < COUT; ++c) {
        const float mean = s_w[c][CIN], scale = s_w[c][CIN + 1], beta = s_b[c];
        float best = -INFINITY, bv = 0.f, by = 0.f;
        float eb[CIN];
#pragma unroll
        for (int t = 0; t < CIN; ++t) eb[t] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float y = 0.f;
#pragma unroll
            for (int t = 0; t < CIN; ++t) y = fmaf(s_w[c][t], e[k][t], y);
            const float v = fmaf(y - mean, scale, beta);
            const float a = lrelu(v);
            if (a > best) {
                best = a; bv = v; by = y;
#pragma unroll
                for (int t = 0; t < CIN; ++t) eb[t] = e[k][t];
            }
        }
        float dm = __ldg(g + (size_t)s * 2 * COUT + COUT + c) * (1.f / (float)P);
        if (__ldg(arg_pt + (size_t)s * COUT + c) == pt) dm += __ldg(g + (size_t)s * 2 * COUT + c);
        const float dv = dm * (bv > 0.f ? 1.f : sgb_bn::SLOPE);
        const float zh = (by - mean) * stats[64 + c];
        float vals[8];
        vals[0] = dv; vals[1] = dv * zh;
#pragma unroll
        for (int t = 0; t < CIN; ++t) vals[2 + t] = dv * (eb[t] - s_ebar[t]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float r = sgb_warp_sum(vals[i]);
            if ((pt & 31) == 0) s_red[pt >> 5][c][i] = r;
        }
    }
    __syncthreads();
    for (int i = pt; i < COUT * 8; i += P) part[(size_t)s * COUT * 8 + i] = (&s_red[0][0][0])[i] + (&s_red[1][0][0])[i];
}

__global__ void __launch_bounds__(512)
backward_reduce_kernel(const float* __restrict__ part, int S, double* __restrict__ sums /*[64*8]*/) {
    const int i = threadIdx.x;
    double a = 0;
    for (int s = 0; s < S; ++s) a += (double)part[(size_t)s * COUT * 8 + i];
    sums[i] = a;
}

__global__ void __launch_bounds__(64)
backward_finalize_kernel(const double* __restrict__ sums, double M, const float* __restrict__ W, const float* __restrict__ stats,
                         const double* __restrict__ mom, float* __restrict__ gW, float* __restrict__ gg, float* __restrict__ gb) {
    __shared__ double s_cov[CIN][CIN];
    for (int n = threadIdx.x; n < CIN * (CIN + 1) / 2; n += blockDim.x) {
        int q = n, t = 0;
        while (q >= CIN - t) { q -= CIN - t; ++t; }
        const int u = t + q;
        const double cv = mom[CIN + n] / M - (mom[t] / M) * (mom[u] / M);
        s_cov[t][u] = cv; s_cov[u][t] = cv;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= COUT) return;
    const double db = sums[c * 8], dg = sums[c * 8 + 1];
    gb[c] = (float)db; gg[c] = (float)dg;
    const double gi = (double)stats[128 + c], is = (double)stats[64 + c];
    for (int t = 0; t < CIN; ++t) {
        double r = 0;
        for (int u = 0; u < CIN; ++u) r += s_cov[t][u] * (double)W[c * CIN + u];
        gW[c * CIN + t] = (float)(gi * (sums[c * 8 + 2 + t] - dg * is * r));
    }
}
}  // namespace sgb_mlp1

using namespace sgb_mlp1;

extern "C" size_t sgb_mlp1_ws_bytes(int S) { return (size_t)((S > 0 ? S : 0) + 1) * NE * sizeof(double); }

// clouds [S,64,6] (output of sgb_cluster_cloud_transform) -> feat [S,128]; knn_idx [S,64,10] local ids;
// stats [4][64], var [64], mom [27] as in sgb_edgeconv_fwd.
extern "C" int sgb_mlp1_fwd(const float* clouds, int S, const float* W, const float* gamma, const float* beta,
                            float* feat, int* knn_idx, int* arg_pt, float* stats, float* var, double* mom,
                            void* ws, size_t ws_bytes, void* stream) {
    if (S <= 0) return S == 0 ? SGB_OK : SGB_ERR_INVALID;
    if (!clouds || !W || !gamma || !beta || !feat || !knn_idx || !stats || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_mlp1_ws_bytes(S)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* part = (double*)ws;
    if (!mom) mom = part + (size_t)S * NE;                 // inference: nobody keeps the moments
    { knn_gram_kernel<<<S, P, 0, st>>>(clouds, S, knn_idx, part); SGB_COUNT_LAUNCH(); }
    sgb_bn::reduce_partials(part, S, NE, mom, st);
    { sgb_bn::bn1_finalize_kernel<CIN><<<1, 64, 0, st>>>(mom, 1, (double)S * P * K, W, nullptr, gamma, beta, stats, var, nullptr); SGB_COUNT_LAUNCH(); }
    { forward_kernel<<<S, P, 0, st>>>(clouds, knn_idx, S, W, stats, feat, arg_pt); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" size_t sgb_mlp1_bwd_ws_bytes(int S) { return (size_t)(S > 0 ? S : 0) * COUT * 8 * sizeof(float) + COUT * 8 * sizeof(double) + 64; }

// g [S,128] -> gW [64,6], ggamma [64], gbeta [64]
extern "C" int sgb_mlp1_bwd(const float* g, const float* clouds, const int* knn_idx, const int* arg_pt, int S, const float* W,
                            const float* stats, const double* mom, float* gW, float* gg, float* gb,
                            void* ws, size_t ws_bytes, void* stream) {
    if (S <= 0) return SGB_ERR_INVALID;
    if (!g || !clouds || !knn_idx || !arg_pt || !W || !stats || !mom || !gW || !gg || !gb || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_mlp1_bwd_ws_bytes(S)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)ws;
    float* part = (float*)((unsigned char*)ws + COUT * 8 * sizeof(double));
    const double M = (double)S * P * K;
    { backward_kernel<<<S, P, 0, st>>>(g, clouds, knn_idx, arg_pt, S, W, stats, mom, M, part); SGB_COUNT_LAUNCH(); }
    sgb_bn::reduce_partials(part, S, COUT * 8, sums, st);
    { backward_finalize_kernel<<<1, 64, 0, st>>>(sums, M, W, stats, mom, gW, gg, gb); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

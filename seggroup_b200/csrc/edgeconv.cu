// a9: EdgeConv point-feature blocks MLP2 / MLP3 (seggroup/model.py:83-138), forward.
//
//   e_ij = (x_j - x_i, x_i)  in R^18, j in knn(i) (k = 20)
//   MLP2: out_i = max_j lrelu(BN1(W1 e_ij))
//   MLP3: out_i = max_j lrelu(BN2(W2 lrelu(BN1(W1 e_ij))))
// BatchNorm always uses the statistics of the current scene (the reference never calls .eval()).
//
// The reference materialises [1,18,N,20] and [1,64,N,20] (5 KB/point).  Here nothing edge-sized ever
// reaches HBM: batch statistics come from second moments accumulated on the fly,
//   pass A : G  = sum_e e' e'^T, s = sum_e e'   (e' = e centred on the scene mean, 18x18)  -> BN1 mean/var
//   pass A2: H  = sum_e h h^T,   t = sum_e h    (h = lrelu(BN1(W1 e)), 64x64, MLP3 only)   -> BN2 mean/var
//            because mean(W h) = W t / M and E[(W h)^2] = W H W^T / M;
//   pass B : recompute the edge MLP from the gathered rows, max over k              -> out [N,64]
// (the same moments are exactly what the analytic backward needs, see edgeconv_bwd.cu).
// One warp per point, lanes own 2 of the 64 output channels; the 20 gathered edge vectors of a point are
// staged in shared memory by lanes 0..19 (one 36 B row each) and read back as broadcast float4.
// Partial sums are fp32 per point / per warp, fp64 across points, reduced in a fixed order (deterministic).
// Algorithmic HBM bytes per point: 36 (x9) + 80 (knn) + 256 (out) per pass B; passes A/A2 read 116 B/point.
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

namespace sgb_ec {
// ------------------------------------------------------------------------------------------------
// scene mean of x9 (centre for the moment accumulation): partial sums per block, fixed-order finish
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) x9_mean_partial(const float* __restrict__ x9, int N, double* __restrict__ part) {
    __shared__ double sm[8][9];
    double acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] += (double)__ldg(x9 + (size_t)p * 9 + t);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = sgb_warp_sum(acc[t]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int t = 0; t < 9; ++t) sm[threadIdx.x >> 5][t] = acc[t];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
        part[(size_t)blockIdx.x * 9 + threadIdx.x] = s;
    }
}
__global__ void x9_mean_finish(const double* __restrict__ part, int nb, int N, float* __restrict__ ctr) {
    if (threadIdx.x < 9) {
        double s = 0;
        for (int b = 0; b < nb; ++b) s += part[(size_t)b * 9 + threadIdx.x];
        ctr[threadIdx.x] = 0.f;
        ctr[9 + threadIdx.x] = (float)(s / (double)N);
    }
}

// ------------------------------------------------------------------------------------------------
// pass A: first and second moments of the (centred) edge features
// entry n < CIN: sum e_n ; entry CIN + pair(t,u), t <= u : sum e_t e_u
// ------------------------------------------------------------------------------------------------
constexpr int NE1_PER_LANE = (NE1 + 31) / 32;   // 6

// One point per lane.  (The first version staged every edge of a point in shared memory, one warp per point, and paid two
// LDS per FMA for all 189 entries: 230 us at 150k points, l1tex 91 %.)  An edge feature is e_k = (d_k, c) with
// d_k = x_j - x_i and c = x_i - centre CONSTANT over the 20 edges of a point, so per point
//     sum_k e_k e_k^T = [ sum_k d_k d_k^T   (sum_k d_k) c^T ]      sum_k e_k = ( sum_k d_k, 20 c )
//                       [       .               20 c c^T    ]
// Only the 45 entries of the d d^T block are per-edge work.  A lane owns one point: it gathers its 20 neighbour rows
// (48-byte padded rows, three 16-byte loads each, several neighbours in flight), keeps d d^T and sum d in registers, and
// the 189 per-point values are summed over the warp's 32 points by shuffles; lane n % 32 keeps entry n in fp64.
__global__ void __launch_bounds__(WARPS * 32)
gram1_pt_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ ctr,
                double* __restrict__ part /*[grid][NE1]*/) {
    __shared__ double s_red[WARPS][NE1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float c9[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) c9[t] = __ldg(ctr + 9 + t);
    double acc[NE1_PER_LANE];
#pragma unroll
    for (int m = 0; m < NE1_PER_LANE; ++m) acc[m] = 0.0;

    for (int p0 = (blockIdx.x * WARPS + warp) * 32; p0 < N; p0 += gridDim.x * WARPS * 32) {
        const int p = p0 + lane;
        const bool valid = p < N;
        float xi[9], c[9], s[9], dd[45];
#pragma unroll
        for (int t = 0; t < 9; ++t) { xi[t] = 0.f; c[t] = 0.f; s[t] = 0.f; }
#pragma unroll
        for (int n = 0; n < 45; ++n) dd[n] = 0.f;
        if (valid) {
            const float4* r = reinterpret_cast<const float4*>(x12 + (size_t)p * 12);
            const float4 a0 = __ldg(r), a1 = __ldg(r + 1), a2 = __ldg(r + 2);
            xi[0] = a0.x; xi[1] = a0.y; xi[2] = a0.z; xi[3] = a0.w; xi[4] = a1.x; xi[5] = a1.y; xi[6] = a1.z; xi[7] = a1.w; xi[8] = a2.x;
#pragma unroll
            for (int t = 0; t < 9; ++t) c[t] = xi[t] - c9[t];
            const int4* kr = reinterpret_cast<const int4*>(knn + (size_t)p * KNN);      // 80-byte rows: 16-byte aligned
#pragma unroll 1
            for (int k4 = 0; k4 < KNN / 4; ++k4) {
                const int4 jj = __ldg(kr + k4);
                const int js[4] = {jj.x, jj.y, jj.z, jj.w};
                float4 b0[4], b1[4], b2[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {                                           // 12 independent 16-byte gathers in flight
                    const float4* rj = reinterpret_cast<const float4*>(x12 + (size_t)js[q] * 12);
                    b0[q] = __ldg(rj); b1[q] = __ldg(rj + 1); b2[q] = __ldg(rj + 2);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float d[9] = {b0[q].x - xi[0], b0[q].y - xi[1], b0[q].z - xi[2], b0[q].w - xi[3], b1[q].x - xi[4],
                                        b1[q].y - xi[5], b1[q].z - xi[6], b1[q].w - xi[7], b2[q].x - xi[8]};
                    int n = 0;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        s[t] += d[t];
#pragma unroll
                        for (int u = t; u < 9; ++u) { dd[n] = fmaf(d[t], d[u], dd[n]); ++n; }
                    }
                }
            }
        }
        // the 189 per-point values in moment order (entry n < 18: sum e_n; then pairs (t, u), t <= u, row by row), summed
        // over the 32 points of the warp; lane n % 32 accumulates entry n
        int n = 0;
        auto emit = [&](float v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGB_FULL_MASK, v, o);
            if ((n & 31) == lane) acc[n >> 5] += (double)v;
            ++n;
        };
#pragma unroll
        for (int t = 0; t < 9; ++t) emit(s[t]);
#pragma unroll
        for (int t = 0; t < 9; ++t) emit(valid ? (float)KNN * c[t] : 0.f);
        {
            int q = 0;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
                for (int u = t; u < 9; ++u) emit(dd[q++]);
#pragma unroll
                for (int u = 0; u < 9; ++u) emit(s[t] * c[u]);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
                for (int u = t; u < 9; ++u) emit((float)KNN * (c[t] * c[u]));
            }
        }
    }
#pragma unroll
    for (int m = 0; m < NE1_PER_LANE; ++m) { const int n = lane + 32 * m; if (n < NE1) s_red[warp][n] = acc[m]; }
    __syncthreads();
    for (int n = threadIdx.x; n < NE1; n += blockDim.x) {
        double sum = 0;
        for (int w = 0; w < WARPS; ++w) sum += s_red[w][n];
        part[(size_t)blockIdx.x * NE1 + n] = sum;
    }
}

// x9 [N,9] -> x12 [N,12] (48-byte rows: three aligned 16-byte loads per gathered row)
__global__ void pad_rows12_kernel(const float* __restrict__ x9, long long n12, float* __restrict__ x12) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n12) return;
    const long long p = i / 12;
    const int q = (int)(i % 12);
    x12[i] = q < 9 ? __ldg(x9 + p * 9 + q) : (q == 11 ? 1.f : 0.f);     // pad lane 11 = 1: "this row was gathered" flag of the tcgen05 producers (a zero-filled row has 0 there)
}

// ------------------------------------------------------------------------------------------------
// pass B: out[p, c] = max_k of the edge MLP
// ------------------------------------------------------------------------------------------------
// Round 2: rows come from the 48-byte padded copy (three aligned 16-byte loads instead of nine scalar ones), BatchNorm-1 is folded
// into the weights, and the two channels of a lane advance with ONE packed FFMA2 per input feature (weights as (w_c0, w_c0+1) pairs,
// the staged edge feature as the broadcast scalar operand).  The edge vector is (x_j - x_i, x_i): its second half is the same for
// the 20 edges of a point, so  y_k = [bias + Wc x_i] + Wd (x_j - x_i)  costs 9 FFMA2 per edge and lane plus 9 per point instead of
// 18 per edge, and only the 9 differences are staged.  LeakyReLU is increasing, so max_k lrelu(y_k) = lrelu(max_k y_k): one
// activation per point (first maximum kept, as torch.max).  Two points per warp iteration keep two independent chains in flight.
constexpr int CD = 9;                     // difference features per edge
constexpr int CDP = 12;                   // staged row: 9 differences + 3 pad (48 B, three 16-byte loads)
__global__ void __launch_bounds__(WARPS * 32)
forward_max_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
                   const float* __restrict__ stats1, float* __restrict__ out, unsigned char* __restrict__ argk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float (*s_e)[2][KNN][CDP] = reinterpret_cast<float (*)[2][KNN][CDP]>(smem_raw);        // [warp][point of the pair][edge][12]
    __shared__ __align__(16) float s_c[WARPS][2][CDP];                                      // centre rows x_i of the pair
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = lane * 2;
    float2 wd[CD], wc[CD];
    const float sc0 = stats1[128 + c0], sc1 = stats1[128 + c0 + 1];
#pragma unroll
    for (int t = 0; t < CD; ++t) {
        wd[t] = make_float2(__ldg(W1 + c0 * CIN + t) * sc0, __ldg(W1 + (c0 + 1) * CIN + t) * sc1);
        wc[t] = make_float2(__ldg(W1 + c0 * CIN + CD + t) * sc0, __ldg(W1 + (c0 + 1) * CIN + CD + t) * sc1);
    }
    const float2 bias = make_float2(fmaf(-sc0, stats1[c0], stats1[192 + c0]), fmaf(-sc1, stats1[c0 + 1], stats1[192 + c0 + 1]));
    for (int p0 = (blockIdx.x * WARPS + warp) * 2; p0 < N; p0 += gridDim.x * WARPS * 2) {
        __syncwarp();
        // stage the 2 x 20 difference vectors: lane l < 20 fetches neighbour l of point p0, lanes 20..31 + a second pass neighbours of p0 + 1
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int slot = pass * 32 + lane;                 // 0..39 used
            if (slot < 2 * KNN) {
                const int pp = slot / KNN, k = slot % KNN;
                const int p = p0 + pp;
                if (p < N) {
                    const float4* xi = reinterpret_cast<const float4*>(x12 + (size_t)p * 12);
                    const float4* xj = reinterpret_cast<const float4*>(x12 + (size_t)__ldg(knn + (size_t)p * KNN + k) * 12);
                    const float4 a0 = __ldg(xi), a1 = __ldg(xi + 1), a2 = __ldg(xi + 2);
                    const float4 b0 = __ldg(xj), b1 = __ldg(xj + 1), b2 = __ldg(xj + 2);
                    float4* d = reinterpret_cast<float4*>(&s_e[warp][pp][k][0]);
                    d[0] = make_float4(b0.x - a0.x, b0.y - a0.y, b0.z - a0.z, b0.w - a0.w);
                    d[1] = make_float4(b1.x - a1.x, b1.y - a1.y, b1.z - a1.z, b1.w - a1.w);
                    d[2] = make_float4(b2.x - a2.x, 0.f, 0.f, 0.f);
                    if (k == 0) {
                        float4* c = reinterpret_cast<float4*>(&s_c[warp][pp][0]);
                        c[0] = a0; c[1] = a1; c[2] = a2;
                    }
                }
            }
        }
        __syncwarp();
        float2 base[2] = {bias, bias};                       // bias + Wc x_i: once per point
#pragma unroll
        for (int t = 0; t < CD; ++t) {
            const float e0 = s_c[warp][0][t], e1 = s_c[warp][1][t];
            sgb_tc::ffma2(base[0], wc[t], make_float2(e0, e0));
            sgb_tc::ffma2(base[1], wc[t], make_float2(e1, e1));
        }
        float2 best[2] = {make_float2(-INFINITY, -INFINITY), make_float2(-INFINITY, -INFINITY)};
        int bk[2][2] = {{0, 0}, {0, 0}};
#pragma unroll 4
        for (int k = 0; k < KNN; ++k) {
            float2 y[2] = {base[0], base[1]};
#pragma unroll
            for (int t4 = 0; t4 < CDP / 4; ++t4) {
                const float4 v0 = *reinterpret_cast<const float4*>(&s_e[warp][0][k][t4 * 4]);
                const float4 v1 = *reinterpret_cast<const float4*>(&s_e[warp][1][k][t4 * 4]);
                const float e0[4] = {v0.x, v0.y, v0.z, v0.w}, e1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t4 * 4 + u;
                    if (t < CD) {
                        sgb_tc::ffma2(y[0], wd[t], make_float2(e0[u], e0[u]));
                        sgb_tc::ffma2(y[1], wd[t], make_float2(e1[u], e1[u]));
                    }
                }
            }
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                if (y[pp].x > best[pp].x) { best[pp].x = y[pp].x; bk[pp][0] = k; }
                if (y[pp].y > best[pp].y) { best[pp].y = y[pp].y; bk[pp][1] = k; }
            }
        }
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const int p = p0 + pp;
            if (p < N) {
                *reinterpret_cast<float2*>(out + (size_t)p * COUT + c0) = make_float2(lrelu(best[pp].x), lrelu(best[pp].y));
                if (argk) *reinterpret_cast<uchar2*>(argk + (size_t)p * COUT + c0) = make_uchar2((unsigned char)bk[pp][0], (unsigned char)bk[pp][1]);
            }
        }
    }
}

inline int persistent_grid(int N) {
    int g = sgb_div_up(N, WARPS);
    const int cap = 148 * 4;
    return g < cap ? (g < 1 ? 1 : g) : cap;
}
}  // namespace sgb_ec

using namespace sgb_ec;

// workspace layout (bytes): [ctr 16 floats][mean partials 256*9 dbl][reduced gram1 192 dbl][gram1 partials grid*NE1 dbl][tcgen05 workspace (two_layer)][x12]
extern "C" size_t sgb_edgeconv_ws_bytes(int N, int two_layer) {
    const size_t g = (size_t)persistent_grid(N);
    size_t b = 128 + 256 * 9 * 8 + 192 * 8 + g * NE1 * 8;
    if (two_layer) { b = (b + 255) & ~(size_t)255; b += sgb_ec2_tc_ws_bytes(N); }
    b = (b + 255) & ~(size_t)255;
    b += (size_t)(N > 0 ? N : 0) * 48;                 // x12: 48-byte padded rows for the 16-byte gathers
    return b;
}

// stats1/stats2: [4][64] float (mean, invstd, gamma*invstd, beta); var1/var2: [64] biased batch variance;
// mom1: [189] double, mom2: [4160] double — moments kept for the backward pass (may be NULL for inference).
extern "C" int sgb_edgeconv_fwd(const float* x9, const int* knn, int N, int two_layer,
                                const float* W1, const float* gamma1, const float* beta1,
                                const float* W2, const float* gamma2, const float* beta2,
                                float* out, unsigned char* argk, float* stats1, float* var1, double* mom1,
                                float* stats2, float* var2, double* mom2, float* ctr_out /*[18] = e0*/,
                                void* ws, size_t ws_bytes, void* stream) {
    if (N <= 0) return N == 0 ? SGB_OK : SGB_ERR_INVALID;
    if (!x9 || !knn || !W1 || !gamma1 || !beta1 || !out || !stats1 || !ws) return SGB_ERR_INVALID;
    if (two_layer && (!W2 || !gamma2 || !beta2 || !stats2)) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_edgeconv_ws_bytes(N, two_layer)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w8 = (unsigned char*)ws;
    float* ctr = (float*)w8;
    double* mpart = (double*)(w8 + 128);
    double* g1red = mpart + 256 * 9;      // [NE1 -> 192] reduced first-layer moments when the caller keeps none
    double* g1part = g1red + 192;
    const int grid = persistent_grid(N);
    const double M = (double)N * KNN;

    const int mb = N < 256 * 256 ? sgb_div_up(N, 256) : 256;
    { x9_mean_partial<<<mb, 256, 0, st>>>(x9, N, mpart); SGB_COUNT_LAUNCH(); }
    { x9_mean_finish<<<1, 32, 0, st>>>(mpart, mb, N, ctr); SGB_COUNT_LAUNCH(); }
    float* x12 = (float*)(w8 + (((sgb_edgeconv_ws_bytes(N, two_layer) - (size_t)N * 48)) & ~(size_t)255));
    { pad_rows12_kernel<<<sgb_div_up((long long)N * 12, 256), 256, 0, st>>>(x9, (long long)N * 12, x12); SGB_COUNT_LAUNCH(); }
    {
        const int gp = sgb_div_up(N, WARPS * 32) < 148 * 2 ? sgb_div_up(N, WARPS * 32) : 148 * 2;    // <= grid: the partials fit
        gram1_pt_kernel<<<gp, WARPS * 32, 0, st>>>(x12, knn, N, ctr, g1part); SGB_COUNT_LAUNCH();
        sgb_bn::reduce_partials(g1part, gp, NE1, mom1 ? mom1 : g1red, st);
    }
    double* m1red = mom1 ? mom1 : g1red;
    { bn1_finalize_kernel<CIN><<<1, 64, 0, st>>>(m1red, 1, M, W1, ctr, gamma1, beta1, stats1, var1, nullptr); SGB_COUNT_LAUNCH(); }
    if (ctr_out) SGB_CUDA(cudaMemcpyAsync(ctr_out, ctr, 18 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (two_layer) {
        // second layer on the tcgen05 tensor cores (mom2 != NULL: also the hidden-layer second moments for the backward)
        size_t off = 128 + 256 * 9 * 8 + 192 * 8 + (size_t)grid * NE1 * 8;
        off = (off + 255) & ~(size_t)255;
        return sgb_ec2_tc_forward(x12, knn, N, W1, stats1, W2, gamma2, beta2, out, argk, stats2, var2, mom2, w8 + off, st);
    }
    {                                      // single layer (MLP2): recompute the edge layer per neighbour, max over k
        const size_t smB = sizeof(float) * WARPS * 2 * KNN * CDP;
        SGB_OPT_IN_SMEM(forward_max_kernel);
        { forward_max_kernel<<<grid, WARPS * 32, smB, st>>>(x12, knn, N, W1, stats1, out, argk); SGB_COUNT_LAUNCH(); }
    }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

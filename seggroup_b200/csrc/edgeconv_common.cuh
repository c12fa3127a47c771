// Shared pieces of the EdgeConv forward / backward kernels (MLP2 / MLP3, seggroup/model.py:83-138).
#pragma once
#include "common.cuh"
#include "bn_moments.cuh"

namespace sgb_ec {
constexpr int CIN = 18;        // edge feature width
constexpr int CINP = 20;       // padded row (float4 aligned)
constexpr int KNN = 20;
constexpr int COUT = 64;
constexpr int WARPS = 8;
constexpr int NE1 = CIN * (CIN + 1) / 2 + CIN;      // 189 first+second moment entries
constexpr int NE2 = COUT * COUT + COUT;             // 4160 (full 64x64 + 64)
using sgb_bn::lrelu;
using sgb_bn::SLOPE;
using sgb_bn::BN_EPS;
using sgb_bn::bn1_finalize_kernel;

// lanes 0..KNN-1 gather one neighbour row each and write e_k to the per-warp staging area.
__device__ __forceinline__ void stage_edges(const float* __restrict__ x9, const int* __restrict__ knn, int p, int lane,
                                            float (*se)[CINP], const float* xi, const float* ctr) {
    if (lane < KNN) {
        const int j = __ldg(knn + (size_t)p * KNN + lane);
        const float* xj = x9 + (size_t)j * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            se[lane][t] = __ldg(xj + t) - xi[t];
            se[lane][9 + t] = ctr ? xi[t] - ctr[t] : xi[t];
        }
        se[lane][18] = 0.f; se[lane][19] = 0.f;
    }
    __syncwarp();
}

// y[c0], y[c0+1] for edge k from the staged row
__device__ __forceinline__ void conv1(const float (*se)[CINP], int k, const float (&w)[2][CIN], float& y0, float& y1) {
    y0 = 0.f; y1 = 0.f;
#pragma unroll
    for (int t4 = 0; t4 < CINP / 4; ++t4) {
        const float4 v = *reinterpret_cast<const float4*>(&se[k][t4 * 4]);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t4 * 4 + u;
            if (t < CIN) { y0 = fmaf(w[0][t], e[u], y0); y1 = fmaf(w[1][t], e[u], y1); }
        }
    }
}

}  // namespace sgb_ec

// tensor-core second layer of MLP3 (edgeconv_tc.cu)
size_t sgb_ec2_tc_ws_bytes(int N);
int sgb_ec2_tc_forward(const float* x12, const int* knn, int N, const float* W1, const float* stats1, const float* W2,
                       const float* gamma2, const float* beta2, float* out, unsigned char* argk, float* stats2, float* var2,
                       double* mom2, void* ws, cudaStream_t st);

// tensor-core dense pass of the MLP3 backward (edgeconv_bwd_tc.cu)
size_t sgb_ec2_bwd_tc_part_bytes(int N);
int sgb_ec2_bwd_tc_dense(const float* x12, const int* knn, int N, const float* W1, const float* stats1, const double* mom1,
                         const float* e0, double M, const float* coef, double* part, int* nparts, cudaStream_t st);

// a5/a7: exact kNN inside clusters (seggroup/model.py:30-36, 512-522).
//
// One thread per query, queries laid out in cluster-major member order so a 128-thread CTA touches one
// large cluster or a few adjacent small ones.  Candidates stream through shared memory as
// (x, y, z, |x|^2) float4 tiles (one broadcast LDS.128 per candidate); every thread keeps its k best
// in registers (fully unrolled insertion, strict '>' so the earlier member wins ties).
// The score is evaluated with explicit __fmul_rn/__fadd_rn/__fmaf_rn so it is bit-identical to the
// fp32 expression torch-CPU evaluates (probed: matmul over K=3 is an fma chain x,y,z; sum(x**2) is
// (x*x + y*y) + z*z) — the ranking is cancellation-prone, so the exact expression matters.
// HBM traffic: 16*N (xyz + order) + 4*N*k (output); candidate re-reads are served by L2/smem.
#include "common.cuh"

namespace {
constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 512;

__device__ __forceinline__ float sq_norm_ref(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
cluster_knn_kernel(const float* __restrict__ xyz, int stride, int N, const int* __restrict__ order,
                   const int* __restrict__ cl_off, int S, int* __restrict__ knn) {
    __shared__ float4 tile[KNN_TILE];
    __shared__ int s_range[2];
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool valid = q < N;
    int lo = 0, hi = 0, pid = 0;
    float xi = 0.f, yi = 0.f, zi = 0.f, xxi = 0.f;
    if (valid) {
        const int c = sgb_upper_segment(cl_off, S, q);
        lo = __ldg(cl_off + c);
        hi = __ldg(cl_off + c + 1);
        pid = __ldg(order + q);
        const float* p = xyz + (size_t)pid * stride;
        xi = __ldg(p); yi = __ldg(p + 1); zi = __ldg(p + 2);
        xxi = sq_norm_ref(xi, yi, zi);
    }
    const int n = hi - lo;
    if (threadIdx.x == 0) s_range[0] = lo;
    const int last = min(blockIdx.x * KNN_THREADS + KNN_THREADS, N) - 1;
    if (q == last) s_range[1] = hi;
    __syncthreads();
    const int r0 = s_range[0], r1 = s_range[1];
    const bool search = valid && n > K;

    float sc[K];
    int id[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { sc[i] = -INFINITY; id[i] = -1; }
    int filled = 0;

    for (int t0 = r0; t0 < r1; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, r1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += KNN_THREADS) {
            const float* p = xyz + (size_t)__ldg(order + t0 + i) * stride;
            const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
            tile[i] = make_float4(x, y, z, sq_norm_ref(x, y, z));
        }
        __syncthreads();
        if (search) {
            const int a = max(lo, t0) - t0, b = min(hi, t0 + tn) - t0;
            for (int j = a; j < b; ++j) {
                const float4 cnd = tile[j];
                float m = __fmul_rn(xi, cnd.x);
                m = __fmaf_rn(yi, cnd.y, m);
                m = __fmaf_rn(zi, cnd.z, m);
                const float inner = __fmul_rn(-2.f, m);
                const float s = __fsub_rn(__fsub_rn(-cnd.w, inner), xxi);
                // the first K candidates always enter (torch.topk keeps NaN/-inf rows well-defined the same way)
                if (filled < K || s > sc[K - 1]) {
                    if (filled < K) ++filled;
                    sc[K - 1] = s; id[K - 1] = t0 + j;
#pragma unroll
                    for (int i = K - 1; i > 0; --i) {
                        if (sc[i] > sc[i - 1]) {
                            const float ts = sc[i]; sc[i] = sc[i - 1]; sc[i - 1] = ts;
                            const int ti = id[i]; id[i] = id[i - 1]; id[i - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (!valid) return;
    int* out = knn + (size_t)pid * K;
    if (search) {
#pragma unroll
        for (int i = 0; i < K; ++i) out[i] = __ldg(order + id[i]);
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) out[i] = (i < n) ? __ldg(order + lo + i) : 0;
    }
}
}  // namespace

extern "C" int sgb_cluster_knn(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                               int k, int* knn, void* stream) {
    if (N < 0 || S < 0 || stride < 3) return SGB_ERR_INVALID;
    if (N == 0) return SGB_OK;
    if (!xyz || !order || !cl_off || !knn || S == 0) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sgb_div_up(N, KNN_THREADS);
    if (k == 20) { cluster_knn_kernel<20><<<grid, KNN_THREADS, 0, st>>>(xyz, stride, N, order, cl_off, S, knn); SGB_COUNT_LAUNCH(); }
    else if (k == 10) { cluster_knn_kernel<10><<<grid, KNN_THREADS, 0, st>>>(xyz, stride, N, order, cl_off, S, knn); SGB_COUNT_LAUNCH(); }
    else if (k == 16) { cluster_knn_kernel<16><<<grid, KNN_THREADS, 0, st>>>(xyz, stride, N, order, cl_off, S, knn); SGB_COUNT_LAUNCH(); }
    else return SGB_ERR_UNSUPPORTED;
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

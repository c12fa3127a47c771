// Dense fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with the
// TF32 x 3 split of tc_common.cuh:  C[M,N] = A[M,K] * B[N,K]^T  (the shape of nn.Linear / a 1x1 convolution).
// Used for the dense heads of the segment graph (GCN fc 192x192 / 256x256, model.py:146-151) and as the unit test of
// the descriptor / TMEM / mbarrier plumbing that the fused EdgeConv and KPConv kernels build on.
//
// One CTA (128 threads) owns a 128-row tile of C and all N <= 256 columns; K is consumed in chunks of 32:
//   all threads: global -> registers -> (hi, lo) split -> canonical no-swizzle smem tiles -> fence.proxy.async
//   thread 0   : 4 k-steps x 3 tcgen05.mma (hi*hi, lo*hi, hi*lo) per chunk, tcgen05.commit -> mbarrier
//   all threads: wait on the mbarrier before the tiles are overwritten; after the last chunk every warp reads its
//                32 TMEM lanes with tcgen05.ld and stores the rows of C.
// Measured with a one-instruction descriptor probe during bring-up (round 1): kind::tf32 reads MN-major operands ONLY in the
// SWIZZLE_128B_BASE32B layout (every other layout type yields zeros), whose swizzle differs from all K-major layouts, so
// one shared-memory tile cannot serve both X W^T and X^T X; this library keeps its tf32 operands K-major.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace sgb_tc;
constexpr int TG_THREADS = 128;
constexpr int TG_KC = 32;                   // K columns per chunk (8 x 16 B)

__global__ void __launch_bounds__(TG_THREADS)
gemm_tf32x3_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K, int tmem_cols, int relu,
                   int k_per_split) {
    extern __shared__ __align__(128) unsigned char tg_smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    unsigned char* sA_hi = tg_smem;                          // [128 x 32] fp32
    unsigned char* sA_lo = sA_hi + 128 * TG_KC * 4;
    unsigned char* sB_hi = sA_lo + 128 * TG_KC * 4;          // [N x 32]
    unsigned char* sB_lo = sB_hi + (size_t)N * TG_KC * 4;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * 128;
    // split-K (gridDim.y > 1): this CTA contracts columns [k_begin, k_end) and writes its own [M,N] slab; the caller sums the slabs
    const int k_begin = blockIdx.y * k_per_split, k_end = min(K, k_begin + k_per_split);
    C += (size_t)blockIdx.y * M * N;

    if (warp == 0) tmem_alloc(&s_tmem, (uint32_t)tmem_cols);
    if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = make_idesc_tf32(128, N, false, false);

    uint32_t phase = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += TG_KC) {
        // ---- stage the chunk (zero-filled past M, N, K)
        {
            const int r = tid, gm = m0 + r;
#pragma unroll
            for (int c4 = 0; c4 < TG_KC / 4; ++c4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int k = k0 + c4 * 4;
                if (gm < M && k < k_end) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * K + k));
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const float4 l = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
                const uint32_t off = tile_off(r, c4 * 4, 128);
                *reinterpret_cast<float4*>(sA_hi + off) = h;
                *reinterpret_cast<float4*>(sA_lo + off) = l;
            }
        }
        for (int r = tid; r < N; r += TG_THREADS) {
#pragma unroll
            for (int c4 = 0; c4 < TG_KC / 4; ++c4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int k = k0 + c4 * 4;
                if (k < k_end) v = __ldg(reinterpret_cast<const float4*>(B + (size_t)r * K + k));
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const float4 l = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
                const uint32_t off = tile_off(r, c4 * 4, N);
                *reinterpret_cast<float4*>(sB_hi + off) = h;
                *reinterpret_cast<float4*>(sB_lo + off) = l;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (warp == 0 && elect_one_sync()) {
            fence_after_sync();
#pragma unroll
            for (int i = 0; i < TG_KC / 8; ++i) {
                const uint32_t a_off = (uint32_t)(2 * i) * (128 * 16), b_off = (uint32_t)(2 * i) * (uint32_t)(N * 16);
                const uint64_t dah = make_desc(smem_u32(sA_hi) + a_off, 128 * 16, 128);
                const uint64_t dal = make_desc(smem_u32(sA_lo) + a_off, 128 * 16, 128);
                const uint64_t dbh = make_desc(smem_u32(sB_hi) + b_off, (uint32_t)N * 16, 128);
                const uint64_t dbl = make_desc(smem_u32(sB_lo) + b_off, (uint32_t)N * 16, 128);
                mma_tf32(tmem, dah, dbh, idesc, k0 > k_begin || i > 0);
                mma_tf32(tmem, dal, dbh, idesc, true);
                mma_tf32(tmem, dah, dbl, idesc, true);
            }
            mma_commit(&s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
    }
    fence_after_sync();
    // ---- epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows m0 + 32w + lane
    const int row = m0 + warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (relu) {                                            // fused epilogue of the GCN layer: relu(fc(.)), model.py:151
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (v[q] > 0.f || v[q] != v[q]) ? v[q] : 0.f;      // NaN propagates, as in torch.relu
        }
        if (row < M) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(C + (size_t)row * N + c0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

}  // namespace

extern "C" int sgb_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, void* stream) {
    return sgb_linear_tf32x3(A, B, C, M, N, K, 0, stream);
}
extern "C" int sgb_linear_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, int relu, void* stream) {
    if (M < 0 || N <= 0 || K <= 0) return SGB_ERR_INVALID;
    if (M == 0) return SGB_OK;
    if (!A || !B || !C) return SGB_ERR_INVALID;
    if (N > 256 || (N & 15) || (K & 3)) return SGB_ERR_UNSUPPORTED;
    int cols = 32;
    while (cols < N) cols <<= 1;
    const size_t smem = (size_t)(2 * 128 + 2 * N) * TG_KC * 4;
    SGB_OPT_IN_SMEM(gemm_tf32x3_kernel);
    gemm_tf32x3_kernel<<<sgb_div_up(M, 128), TG_THREADS, smem, (cudaStream_t)stream>>>(A, B, C, M, N, K, cols, relu, K); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
// Split-K variant for "tall" contractions (the weight gradient dW = dZ^T X of the GCN layer: M, N <= 256, K = number of
// clusters): Cpart [ksplit, M, N], slab s = A[:, s*kps : (s+1)*kps] B[:, same]^T with kps = k_per_split (a multiple of 32);
// the caller sums the slabs (fixed order -> deterministic).
extern "C" int sgb_gemm_tf32x3_splitk(const float* A, const float* B, float* Cpart, int M, int N, int K, int k_per_split, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || k_per_split <= 0 || (k_per_split & 31)) return SGB_ERR_INVALID;
    if (!A || !B || !Cpart) return SGB_ERR_INVALID;
    if (N > 256 || (N & 15) || (K & 3)) return SGB_ERR_UNSUPPORTED;
    int cols = 32;
    while (cols < N) cols <<= 1;
    const size_t smem = (size_t)(2 * 128 + 2 * N) * TG_KC * 4;
    SGB_OPT_IN_SMEM(gemm_tf32x3_kernel);
    dim3 grid(sgb_div_up(M, 128), sgb_div_up(K, k_per_split));
    gemm_tf32x3_kernel<<<grid, TG_THREADS, smem, (cudaStream_t)stream>>>(A, B, Cpart, M, N, K, cols, 0, k_per_split); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

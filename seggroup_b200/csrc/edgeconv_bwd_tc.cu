// a9 backward, dense pass of MLP3 on the tensor cores (SIMT version: bwd_dense_kernel in edgeconv_bwd.cu).
//
// Training-mode BatchNorm-2 couples every edge to the batch statistics, so the gradient that reaches the hidden layer has a
// dense part over ALL N*20 edges (edgeconv_bwd.cu, header):
//     dv1[e, c] = (r[c] - sum_j Bm[c, j] h[e, j]) * lrelu'(v1[e, c]),      h[e, :] = lrelu(BN1(W1 e_e))
// and the first-layer parameter gradients need   A0[c] = sum_e dv1[e, c],   T[c, t] = sum_e dv1[e, c] (e_e[t] - ebar[t])
// (the dgamma sum follows from T:  sum_e dv1 zhat1 = invstd1[c] * W1[c, :] . T[c, :], because mean(W1 e) = W1 ebar).
// The 64x64 mat-vec per edge (8.2 kFLOP x 3 M edges) is the same GEMM shape as the forward second layer, so this kernel is
// the forward tcgen05 kernel (edgeconv_tc.cu) with Bm in place of W2:  D[64 channels, 160 edges] = Bm * H^T, kind::tf32 x 3.
// Differences from the forward kernel:
//   * the producers also leave, per edge, the sign bits of the 64 hidden activations (2 words) and the centred edge vector
//     (18 floats) in the stage, for the epilogue;
//   * the epilogue owns the stage until it is done with it (bar_empty is arrived by the epilogue warps, not by the MMA commit):
//     thread = channel (lanes 0..15 of each warp hold the M = 64 accumulator rows), the partner lane 16 + l takes half of the
//     18 columns of T for the same channel (dv1 by shuffle), accumulators fp32 per 8 tiles, fp64 across;
//   * outputs: per-CTA partial sums part[cta][64][20] = (A0, sum dv1 zhat1, T[18]) in fp64, reduced in a fixed order.
// Measured (profiles/r01i_ncu_full_train_150k.txt): 827 us at 150k points, l1tex 88 %, tensor pipe 16.5 %: the kernel is
// shared-memory-bandwidth bound, and the largest consumer is the epilogue's broadcast reads of the centred edge vectors
// (every warp re-reads the 72 bytes of every edge, two distinct addresses per wavefront).  Spreading the epilogue over eight
// warps with pipelined TMEM loads did not help (0.93 -> 0.96 ms: same shared-memory traffic); the next step is to run
// T = dv1^T (e - ebar) as a second tcgen05 GEMM over the edges (K = edges, as the forward Gram accumulator does), which needs
// 80-edge tiles to fit the dv1 operand next to the H tiles.
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

namespace sgb_ecbt {
using namespace sgb_tc;
using sgb_ec::CIN;
using sgb_ec::COUT;
using sgb_ec::KNN;
using sgb_bn::lrelu;
using sgb_bn::SLOPE;

constexpr int EPI_WARPS = 4, PROD_WARPS = 10;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;     // 480
constexpr int PROD_THREADS = PROD_WARPS * 32;                  // 320
constexpr int TE = 160;                                        // edges per tile
constexpr int PTS = TE / KNN;                                  // 8 points per tile
constexpr int PPE = PROD_THREADS / TE;                         // 2 producer threads per edge
constexpr int CPT = 16 / PPE;                                  // 8 chunks of 4 hidden channels per producer thread
constexpr int BM_BYTES = COUT * COUT * 4;
constexpr int TILE_BYTES = TE * COUT * 4;                      // one K-major H tile (hi or lo)
constexpr int DROW = 20;                                       // floats per edge row of the centred edge vectors
constexpr int D_BYTES = TE * DROW * 4;
constexpr int MASK_BYTES = TE * 2 * 4;
constexpr int STAGE_BYTES = 2 * TILE_BYTES + D_BYTES + MASK_BYTES;        // 96,000: a multiple of 128
constexpr int W1T_STRIDE = 80;
constexpr int NACC = CIN + 2;
constexpr int FLUSH_TILES = 8;
constexpr int TMEM_COLS = 512;
constexpr int Z_COL = 256;

constexpr int off_bm_hi = 0;
constexpr int off_bm_lo = off_bm_hi + BM_BYTES;
constexpr int off_stage0 = off_bm_lo + BM_BYTES;
constexpr int off_w1t = off_stage0 + 2 * STAGE_BYTES;
constexpr int off_b1 = off_w1t + CIN * W1T_STRIDE * 4;
constexpr int off_ebar = off_b1 + COUT * 4;
constexpr int off_bars = off_ebar + 32 * 4;
constexpr int off_tmem_slot = off_bars + 8 * 8;
constexpr int SMEM_TOTAL = off_tmem_slot + 16;
static_assert(STAGE_BYTES % 128 == 0, "stage alignment");
static_assert(SMEM_TOTAL + 128 <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(THREADS, 1)
ec2_bwd_tc_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
                  const float* __restrict__ stats1, const double* __restrict__ mom1, const float* __restrict__ e0, double M,
                  const float* __restrict__ coef /*[64*64 Bm][64 r]*/, double* __restrict__ part /*[grid][64*NACC]*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + off_bars);         // [2] producers -> MMA
    uint64_t* bar_empty = bar_full + 2;                                      // [2] epilogue -> producers (stage + TMEM buffer free)
    uint64_t* bar_tfull = bar_full + 4;                                      // [2] MMA (commit) -> epilogue
    uint64_t* bar_tempty = bar_full + 6;                                     // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + off_tmem_slot);
    float* s_w1t = reinterpret_cast<float*>(sm + off_w1t);
    float* s_b1 = reinterpret_cast<float*>(sm + off_b1);
    float* s_ebar = reinterpret_cast<float*>(sm + off_ebar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int per = N / gridDim.x, rem = N % gridDim.x;
    const int p_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int p_end = p_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
    const long long g_begin = (long long)p_begin * KNN, g_end = (long long)p_end * KNN;
    const int ntiles = (int)((g_end - g_begin + TE - 1) / TE);

    // ---- one-time setup: Bm (symmetric) -> K-major canonical tiles, hi / lo split
    for (int i = tid; i < COUT * COUT; i += THREADS) {
        const int c = i / COUT, j = i % COUT;
        const float w = __ldg(coef + i);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, j, COUT);
        *reinterpret_cast<float*>(sm + off_bm_hi + off) = hi;
        *reinterpret_cast<float*>(sm + off_bm_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < COUT * CIN; i += THREADS) {
        const int c = i / CIN, q = i % CIN;
        s_w1t[q * W1T_STRIDE + (c >> 4) * 20 + (c & 15)] = __ldg(W1 + i) * stats1[128 + c];
    }
    for (int i = tid; i < COUT; i += THREADS) s_b1[i] = fmaf(-stats1[128 + i], stats1[i], stats1[192 + i]);
    if (tid < CIN) s_ebar[tid] = (float)(mom1[tid] / M + (double)e0[tid]);
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_empty[0], EPI_WARPS); mbar_init(&bar_empty[1], EPI_WARPS);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], EPI_WARPS); mbar_init(&bar_tempty[1], EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp > MMA_WARP) {
        // ================= producers (as the forward kernel) + sign bits and centred edge vectors for the epilogue
        const int pw = warp - (MMA_WARP + 1);
        const int er = pw * (32 / PPE) + lane / PPE;
        const int part_id = lane % PPE;
        float en[CIN];
        int j_next = 0;
        bool v_next = false;
        auto issue_index = [&](int t) -> int {
            const long long g = g_begin + (long long)t * TE + er;
            return (t < ntiles && g < g_end) ? __ldg(knn + g) : -1;
        };
        auto issue_rows = [&](int t, int j) {
            const long long g = g_begin + (long long)t * TE + er;
            v_next = j >= 0;
            if (v_next) {
                const float4* xi = reinterpret_cast<const float4*>(x12 + (size_t)(g / KNN) * 12);
                const float4* xj = reinterpret_cast<const float4*>(x12 + (size_t)j * 12);
                const float4 a0 = __ldg(xi), a1 = __ldg(xi + 1), a2 = __ldg(xi + 2);
                const float4 b0 = __ldg(xj), b1 = __ldg(xj + 1), b2 = __ldg(xj + 2);
                en[0] = b0.x - a0.x; en[1] = b0.y - a0.y; en[2] = b0.z - a0.z; en[3] = b0.w - a0.w;
                en[4] = b1.x - a1.x; en[5] = b1.y - a1.y; en[6] = b1.z - a1.z; en[7] = b1.w - a1.w;
                en[8] = b2.x - a2.x;
                en[9] = a0.x; en[10] = a0.y; en[11] = a0.z; en[12] = a0.w;
                en[13] = a1.x; en[14] = a1.y; en[15] = a1.z; en[16] = a1.w; en[17] = a2.x;
            }
        };
        issue_rows(0, issue_index(0));
        j_next = issue_index(1);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            float2 ee[CIN];
            float ec[CIN];
            const bool valid = v_next;
#pragma unroll
            for (int q = 0; q < CIN; ++q) { ee[q] = make_float2(en[q], en[q]); ec[q] = valid ? en[q] - s_ebar[q] : 0.f; }
            issue_rows(t + 1, j_next);
            j_next = issue_index(t + 2);
            mbar_wait(&bar_empty[st], ph ^ 1u);
            unsigned char* dst_hi = sm + off_stage0 + st * STAGE_BYTES;
            unsigned char* dst_lo = dst_hi + TILE_BYTES;
            float* dst_d = reinterpret_cast<float*>(dst_lo + TILE_BYTES) + er * DROW;
            uint32_t* dst_m = reinterpret_cast<uint32_t*>(dst_lo + TILE_BYTES + D_BYTES) + er * 2;
            uint32_t bits = 0;
#pragma unroll 2
            for (int i4 = 0; i4 < CPT; ++i4) {
                const int c4 = part_id * CPT + i4;
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float4 b = *reinterpret_cast<const float4*>(s_b1 + c4 * 4);
                    float2 y01 = make_float2(b.x, b.y), y23 = make_float2(b.z, b.w);
                    const float* wrow = s_w1t + (c4 >> 2) * 20 + (c4 & 3) * 4;
#pragma unroll
                    for (int q = 0; q < CIN; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(wrow + q * W1T_STRIDE);
                        ffma2(y01, make_float2(w.x, w.y), ee[q]);
                        ffma2(y23, make_float2(w.z, w.w), ee[q]);
                    }
                    bits |= (y01.x > 0.f ? 1u : 0u) << (4 * i4) | (y01.y > 0.f ? 2u : 0u) << (4 * i4) |
                            (y23.x > 0.f ? 4u : 0u) << (4 * i4) | (y23.y > 0.f ? 8u : 0u) << (4 * i4);
                    y = make_float4(lrelu(y01.x), lrelu(y01.y), lrelu(y23.x), lrelu(y23.y));
                }
                const float4 hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                const float4 lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
                const uint32_t off = (uint32_t)c4 * (TE * 16) + (uint32_t)(er >> 3) * 128 + (uint32_t)(er & 7) * 16;
                *reinterpret_cast<float4*>(dst_hi + off) = hi;
                *reinterpret_cast<float4*>(dst_lo + off) = lo;
            }
            dst_m[part_id] = bits;                      // sign bits of hidden channels 32 part_id .. 32 part_id + 31
            if (part_id == 0) {                         // centred edge vector: [0..7] = t 0..7, [8..15] = t 8..15, [16], [17]
                float4* d4 = reinterpret_cast<float4*>(dst_d);
                d4[0] = make_float4(ec[0], ec[1], ec[2], ec[3]);
                d4[1] = make_float4(ec[4], ec[5], ec[6], ec[7]);
                d4[2] = make_float4(ec[8], ec[9], ec[10], ec[11]);
                d4[3] = make_float4(ec[12], ec[13], ec[14], ec[15]);
                d4[4] = make_float4(ec[16], ec[17], 0.f, 0.f);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[st]);
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer
        const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
        const uint32_t a_hi = smem_u32(sm + off_bm_hi), a_lo = smem_u32(sm + off_bm_lo);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_full[st], ph);
            mbar_wait(&bar_tempty[st], ph ^ 1u);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t b_hi = smem_u32(sm + off_stage0 + st * STAGE_BYTES), b_lo = b_hi + TILE_BYTES;
                const uint32_t d = tmem + (uint32_t)(st * Z_COL);
#pragma unroll
                for (int i = 0; i < COUT / 8; ++i) {
                    const uint32_t ao = (uint32_t)(2 * i) * (COUT * 16), bo = (uint32_t)(2 * i) * (TE * 16);
                    const uint64_t dah = make_desc(a_hi + ao, COUT * 16, 128), dal = make_desc(a_lo + ao, COUT * 16, 128);
                    const uint64_t dbh = make_desc(b_hi + bo, TE * 16, 128), dbl = make_desc(b_lo + bo, TE * 16, 128);
                    mma_tf32(d, dah, dbh, idesc, i > 0);
                    mma_tf32(d, dal, dbh, idesc, true);
                    mma_tf32(d, dah, dbl, idesc, true);
                }
                mma_commit(&bar_tfull[st]);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: channel c = 16 warp + (lane & 15); lane half h = lane >> 4 takes columns
        //                   t = 8h .. 8h + 7 and 16 + h of T for that channel
        const int c = warp * 16 + (lane & 15);
        const int half = lane >> 4;
        const float rc = __ldg(coef + COUT * COUT + c);
        const uint32_t word = (uint32_t)(c >> 5), bit = (uint32_t)(c & 31);
        float a0 = 0.f, tt[9];
        double A0 = 0.0, TT[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { tt[i] = 0.f; TT[i] = 0.0; }
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_tfull[st], ph);
            fence_after_sync();
            const long long g0 = g_begin + (long long)t * TE;
            const int npts = (int)min((long long)PTS, (g_end - g0) / KNN);
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * Z_COL);
            const unsigned char* stage = sm + off_stage0 + st * STAGE_BYTES + 2 * TILE_BYTES;
            const float* s_d = reinterpret_cast<const float*>(stage);
            const uint32_t* s_m = reinterpret_cast<const uint32_t*>(stage + D_BYTES);
#pragma unroll 1
            for (int pp = 0; pp < npts; ++pp) {
                float v[16], u[4];
                tmem_ld16(taddr + (uint32_t)(pp * KNN), v);
                tmem_ld4(taddr + (uint32_t)(pp * KNN + 16), u);
#pragma unroll
                for (int k = 0; k < KNN; ++k) {
                    const int e = pp * KNN + k;
                    const float z = k < 16 ? v[k] : u[k - 16];
                    const bool posv = (s_m[e * 2 + word] >> bit) & 1u;
                    float dv = (rc - z) * (posv ? 1.f : SLOPE);          // meaningful on lanes 0..15 (accumulator rows)
                    dv = __shfl_sync(SGB_FULL_MASK, dv, lane & 15);
                    a0 += dv;
                    const float4 d0 = *reinterpret_cast<const float4*>(s_d + e * DROW + half * 8);
                    const float4 d1 = *reinterpret_cast<const float4*>(s_d + e * DROW + half * 8 + 4);
                    const float d2 = s_d[e * DROW + 16 + half];
                    tt[0] = fmaf(dv, d0.x, tt[0]); tt[1] = fmaf(dv, d0.y, tt[1]); tt[2] = fmaf(dv, d0.z, tt[2]); tt[3] = fmaf(dv, d0.w, tt[3]);
                    tt[4] = fmaf(dv, d1.x, tt[4]); tt[5] = fmaf(dv, d1.y, tt[5]); tt[6] = fmaf(dv, d1.z, tt[6]); tt[7] = fmaf(dv, d1.w, tt[7]);
                    tt[8] = fmaf(dv, d2, tt[8]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bar_tempty[st]); mbar_arrive(&bar_empty[st]); }
            if ((t % FLUSH_TILES) == FLUSH_TILES - 1 || t == ntiles - 1) {
                A0 += (double)a0; a0 = 0.f;
#pragma unroll
                for (int i = 0; i < 9; ++i) { TT[i] += (double)tt[i]; tt[i] = 0.f; }
            }
        }
        // part[cta][c][0] = A0, [1] = sum dv1 zhat1 = invstd1 * W1[c,:].T[c,:], [2 + t] = T[c][t]
        double* dst = part + ((size_t)blockIdx.x * COUT + c) * NACC;
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int tcol = i < 8 ? half * 8 + i : 16 + half;
            dst[2 + tcol] = TT[i];
            dot += (double)__ldg(W1 + c * CIN + tcol) * TT[i];
        }
        dot += __shfl_xor_sync(SGB_FULL_MASK, dot, 16);
        if (half == 0) { dst[0] = A0; dst[1] = (double)stats1[64 + c] * dot; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

inline int grid_for(int N) {
    const int tiles = sgb_div_up((long long)N * KNN, TE);
    return tiles < 148 ? (tiles < 1 ? 1 : tiles) : 148;
}
}  // namespace sgb_ecbt

size_t sgb_ec2_bwd_tc_part_bytes(int N) { return (size_t)sgb_ecbt::grid_for(N) * sgb_ec::COUT * sgb_ecbt::NACC * sizeof(double); }

// Dense pass on the tensor cores.  x12: 48-byte padded rows; coef: Bm [64][64] then r [64] (bwd_mid_kernel);
// part: sgb_ec2_bwd_tc_part_bytes(N) bytes; *nparts <- number of per-CTA partial rows written.
int sgb_ec2_bwd_tc_dense(const float* x12, const int* knn, int N, const float* W1, const float* stats1, const double* mom1,
                         const float* e0, double M, const float* coef, double* part, int* nparts, cudaStream_t st) {
    using namespace sgb_ecbt;
    const int grid = grid_for(N);
    SGB_OPT_IN_SMEM(ec2_bwd_tc_kernel);
    { ec2_bwd_tc_kernel<<<grid, THREADS, SMEM_TOTAL + 128, st>>>(x12, knn, N, W1, stats1, mom1, e0, M, coef, part); SGB_COUNT_LAUNCH(); }
    *nparts = grid;
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

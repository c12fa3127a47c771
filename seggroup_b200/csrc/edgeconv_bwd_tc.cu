// a9 backward, dense pass of MLP3 on the tensor cores (SIMT version: bwd_dense_kernel in edgeconv_bwd.cu).
//
// Training-mode BatchNorm-2 couples every edge to the batch statistics, so the gradient that reaches the hidden layer has a
// dense part over ALL N*20 edges (edgeconv_bwd.cu, header):
//     dv1[e, c] = (r[c] - sum_j Bm[c, j] h[e, j]) * lrelu'(v1[e, c]),      h[e, :] = lrelu(BN1(W1 e_e))
// and the first-layer parameter gradients need   A0[c] = sum_e dv1[e, c],   T[c, t] = sum_e dv1[e, c] (e_e[t] - ebar[t])
// (the dgamma sum follows from T:  sum_e dv1 zhat1 = invstd1[c] * W1[c, :] . T[c, :], because mean(W1 e) = W1 ebar).
//
// Round 2 design — BOTH contractions run on tcgen05 (kind::tf32 x 3, tc_common.cuh):
//   (1) z = Bm h per edge:   D_z[64 channels, 80 edges] = Bm[64, 64] * H[80, 64]^T          (as the forward second layer)
//   (2) T and A0:            D_T[64 channels, 24]      += dv1^T[64, 80 edges] * EC[24, 80 edges]^T, K = the edges of the tile,
//       EC rows 0..17 = centred edge vector, row 18 = 1 for a real edge (-> A0), rows 19..23 = 0.
// Round 1 accumulated (2) on the CUDA cores in the epilogue: 9 FMA + 3 shared-memory loads + a shuffle per edge and thread,
// one warp per scheduler, latency bound (profiles/r02o: 400 instructions per point and warp); the shared-memory reads of the
// centred vectors plus the producers' weight reads put the kernel at 88 % l1tex.  Now the epilogue only turns z into dv1
// (20 values per point in registers -> hi / lo split -> 16-byte stores of 4 consecutive edges into a K-major operand tile) and
// the tensor cores do the rest; D_T stays in TMEM for 32 tiles, is flushed to fp64 accumulators, and double buffered.
// Producers: as the forward kernel (edgeconv_tc.cu) — first-layer weights in registers, cp.async gather ring 3 tiles ahead —
// plus, per edge, the sign bits of the hidden activations (a byte per 4-channel chunk) and the transposed centred edge vector.
// Pipelines (mbarriers): H stage full / empty (empty = tcgen05.commit after the T MMAs of the tile), z accumulators full /
// empty (2 TMEM buffers), dv1 tile full / empty (1 buffer), T accumulator segment full / empty (2 TMEM buffers).
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

#ifndef SGB_ABL
#define SGB_ABL 0      // role-ablation timing experiments (tools/ablate.sh): results are WRONG for any value but 0
#endif

namespace sgb_ecbt {
using namespace sgb_tc;
using sgb_ec::CIN;
using sgb_ec::COUT;
using sgb_ec::KNN;
using sgb_bn::lrelu;
using sgb_bn::SLOPE;

constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;     // 416
constexpr int PROD_THREADS = PROD_WARPS * 32;                  // 256
constexpr int TE = 80;                                         // edges per tile
constexpr int PTS = TE / KNN;                                  // 4 points per tile
constexpr int BLOCKS = TE / 8;                                 // 8-edge blocks per tile; a group of 4 producer warps takes every 2nd one
constexpr int TN = 24;                                         // columns of the T accumulator: 18 + A0 + 5 zero
constexpr int BM_BYTES = COUT * COUT * 4;
constexpr int TILE_BYTES = TE * COUT * 4;                      // one K-major H tile (hi or lo); also one dv1^T tile [64 rows][80 edges]
constexpr int EC_BYTES = TN * TE * 4;                          // one K-major EC tile [24 rows][80 edges] (hi or lo)
constexpr int MASK_BYTES = TE * 16;                            // one byte per (edge, 4-channel chunk): sign bits of the hidden activations
constexpr int HSTAGE_BYTES = 2 * TILE_BYTES;                   // H hi / lo: free again as soon as the z MMAs of the tile have completed
constexpr int ESTAGE_BYTES = 2 * EC_BYTES + MASK_BYTES;        // EC hi / lo + sign bytes: live until the T MMAs of the tile have completed
constexpr int ERING = 3;
constexpr int DVH_BYTES = COUT * (TE / 2) * 4;                 // dv1^T of HALF a tile (2 points = 40 edges), hi or lo
constexpr int RING = 3;                                        // tiles in flight in the gather ring
constexpr int RAW_BYTES = TE * 48 + PTS * 48 + TE * 8;         // x_j rows, x_i rows, validity words, neighbour indices of a later tile
constexpr int NACC = CIN + 2;
constexpr int ACCN = 20;                                       // fp64 accumulator columns kept per channel (18 + A0 + pad)
constexpr int FLUSH = 32;                                      // tiles per T segment (fp32 in TMEM), then fp64
constexpr int TMEM_COLS = 512;
constexpr int Z_COL = 256;                                     // z accumulators at columns 0 and 256
constexpr int T_COL0 = 128, T_COLS = 64;                       // T accumulators at columns 128 and 192

constexpr int off_bm_hi = 0;
constexpr int off_bm_lo = off_bm_hi + BM_BYTES;
constexpr int off_h0 = off_bm_lo + BM_BYTES;                   // 2 H stages
constexpr int off_e0 = off_h0 + 2 * HSTAGE_BYTES;              // 3 EC / mask stages
constexpr int off_dv = off_e0 + ERING * ESTAGE_BYTES;          // [half][hi, lo]
constexpr int off_raw0 = off_dv + 4 * DVH_BYTES;
constexpr int off_wc = off_raw0 + RING * RAW_BYTES;            // centre-half first-layer weights [16 chunks][9] float4 (BN1 folded in)
constexpr int off_ctr = off_wc + 16 * 9 * 16;                  // per producer warp: centre rows [PTS][4 parts] float4 of the current tile
constexpr int off_ebar = off_ctr + PROD_WARPS * PTS * 4 * 16;
constexpr int off_bars = off_ebar + 32 * 4;                    // 19 mbarriers
constexpr int off_tmem_slot = off_bars + 20 * 8;
constexpr int SMEM_TOTAL = off_tmem_slot + 16;
static_assert(HSTAGE_BYTES % 128 == 0 && ESTAGE_BYTES % 128 == 0 && DVH_BYTES % 128 == 0, "stage alignment");
static_assert(SMEM_TOTAL + 128 <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ bool mbar_test(uint64_t* mbar, uint32_t parity) {       // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    return ok != 0;
}

// byte offset of (row r, K index e) in a canonical no-swizzle K-major tile with R rows (tc_common.cuh tile_off with c = e)
__device__ __forceinline__ uint32_t kmajor_off(int r, int e, int R) { return (uint32_t)((e >> 2) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (e & 3) * 4); }

__global__ void __launch_bounds__(THREADS, 1)      // 13 warps: one scheduler hosts 4 of them -> 128 registers per thread at most
ec2_bwd_tc_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
                  const float* __restrict__ stats1, const double* __restrict__ mom1, const float* __restrict__ e0, double M,
                  const float* __restrict__ coef /*[64*64 Bm][64 r]*/, double* __restrict__ part /*[grid][64*NACC]*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + off_bars);         // [2] producers -> MMA
    uint64_t* bar_hempty = bar_full + 2;                                     // [2] MMA (commit after the z MMAs) -> producers: H tiles free
    uint64_t* bar_tfull = bar_full + 4;                                      // [2] MMA (commit) -> epilogue: z ready
    uint64_t* bar_tempty = bar_full + 6;                                     // [2] epilogue -> MMA: z buffer free
    uint64_t* bar_dvfull = bar_full + 8;                                     // [2] epilogue -> MMA: half of the dv1 tile written
    uint64_t* bar_dvempty = bar_full + 10;                                   // [2] MMA (commit) -> epilogue: that half consumed
    uint64_t* bar_gfull = bar_full + 12;                                     // [2] MMA (commit) -> epilogue: T segment complete
    uint64_t* bar_gempty = bar_full + 14;                                    // [2] epilogue -> MMA: T accumulator flushed
    uint64_t* bar_eempty = bar_full + 16;                                    // [3] MMA (commit after the T MMAs) -> producers: EC / mask stage free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + off_tmem_slot);
    float* s_ebar = reinterpret_cast<float*>(sm + off_ebar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int per = N / gridDim.x, rem = N % gridDim.x;
    const int p_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int p_end = p_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
    const long long g_begin = (long long)p_begin * KNN, g_end = (long long)p_end * KNN;
    const int ntiles = (int)((g_end - g_begin + TE - 1) / TE);

    // ---- one-time setup: Bm (symmetric) -> K-major canonical tiles, hi / lo split; constant rows of the EC tiles; accumulators
    for (int i = tid; i < COUT * COUT; i += THREADS) {
        const int c = i / COUT, j = i % COUT;
        const float w = __ldg(coef + i);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, j, COUT);
        *reinterpret_cast<float*>(sm + off_bm_hi + off) = hi;
        *reinterpret_cast<float*>(sm + off_bm_lo + off) = tf32_hi(w - hi);
    }
    for (int s2 = 0; s2 < ERING; ++s2) {                                      // rows 18..23 of the EC tiles of every stage: zero (row 18 is
        unsigned char* ec = sm + off_e0 + s2 * ESTAGE_BYTES;                  // rewritten per edge by the producers, 19..23 stay zero)
        for (int i = tid; i < 6 * TE; i += THREADS) {
            const int r = 18 + i / TE, e = i % TE;
            *reinterpret_cast<float*>(ec + kmajor_off(r, e, TN)) = 0.f;
            *reinterpret_cast<float*>(ec + EC_BYTES + kmajor_off(r, e, TN)) = 0.f;
        }
    }
    for (int i = tid; i < 16 * 9; i += THREADS) {                             // centre half of the first layer: columns 9..17 of W1, BN1 scale folded in
        const int ch = i / 9, q = i % 9;
        float w[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) w[s4] = __ldg(W1 + (4 * ch + s4) * CIN + 9 + q) * stats1[128 + 4 * ch + s4];
        *reinterpret_cast<float4*>(sm + off_wc + i * 16) = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (tid < CIN) s_ebar[tid] = (float)(mom1[tid] / M + (double)e0[tid]);
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_hempty[0], 1); mbar_init(&bar_hempty[1], 1);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], EPI_WARPS); mbar_init(&bar_tempty[1], EPI_WARPS);
        mbar_init(&bar_dvfull[0], EPI_WARPS); mbar_init(&bar_dvfull[1], EPI_WARPS);
        mbar_init(&bar_dvempty[0], 1); mbar_init(&bar_dvempty[1], 1);
        mbar_init(&bar_eempty[0], 1); mbar_init(&bar_eempty[1], 1); mbar_init(&bar_eempty[2], 1);
        mbar_init(&bar_gfull[0], 1); mbar_init(&bar_gfull[1], 1);
        mbar_init(&bar_gempty[0], EPI_WARPS); mbar_init(&bar_gempty[1], EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp > MMA_WARP) {
        // ================= producers (as the forward kernel) + sign bits and transposed centred edge vectors
        const int pw = warp - (MMA_WARP + 1);           // 0..7
        const int ptid = pw * 32 + lane;                // gather role: edge row `ptid` (< TE), point row `ptid` (< PTS)
        const int grp = pw >> 2;
        const int c4 = (lane >> 3) * 4 + (pw & 3);      // my chunk of the hidden vector: channels 4 c4 .. 4 c4 + 3
        const int eb = lane & 7;
        const int part = lane >> 3;
        // as the forward kernel: only the 9 difference columns of the first layer are per-edge work (weights in registers); the centre
        // half  bias + Wc x_i  is evaluated once per point and tile
        constexpr int CD = 9;
        float2 w01[CD], w23[CD];
        float4 bias;
        {
            float sc[4], bb[4];
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int c = 4 * c4 + s4;
                sc[s4] = stats1[128 + c];
                bb[s4] = fmaf(-stats1[128 + c], stats1[c], stats1[192 + c]);
            }
#pragma unroll
            for (int q = 0; q < CD; ++q) {
                w01[q] = make_float2(__ldg(W1 + (4 * c4 + 0) * CIN + q) * sc[0], __ldg(W1 + (4 * c4 + 1) * CIN + q) * sc[1]);
                w23[q] = make_float2(__ldg(W1 + (4 * c4 + 2) * CIN + q) * sc[2], __ldg(W1 + (4 * c4 + 3) * CIN + q) * sc[3]);
            }
            bias = make_float4(bb[0], bb[1], bb[2], bb[3]);
        }
        const float4* wc = reinterpret_cast<const float4*>(sm + off_wc) + c4 * 9;
        unsigned char* ctr = sm + off_ctr + pw * (PTS * 4 * 16);
        // my EC rows: row c4 = (x_j[c4] - x_i[c4]) - ebar[c4] for c4 < 9, x_i[c4 - 9] - ebar[c4] otherwise; rows 16 + c4 for c4 < 3
        const uint32_t ec_qj = (uint32_t)(c4 < 9 ? c4 : 0), ec_qi = (uint32_t)(c4 < 9 ? c4 : c4 - 9);
        const float ebar0 = s_ebar[c4], ebar1 = c4 < 2 ? s_ebar[16 + c4] : 0.f;
        const bool gatherer = ptid < TE, pgatherer = ptid < PTS;
        auto edge_in_range = [&](int t) -> bool {
            const long long g = g_begin + (long long)t * TE + ptid;
            return gatherer && t < ntiles && g < g_end;
        };
        auto cp16 = [&](void* dst, const float* src, bool valid) {
            const uint32_t n = valid ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
        };
        auto issue_index_copy = [&](int t) {
            if (edge_in_range(t)) {
                unsigned char* slot = sm + off_raw0 + (t % RING) * RAW_BYTES + TE * 48 + PTS * 48 + TE * 4 + ptid * 4;
                const int* src = knn + (g_begin + (long long)t * TE + ptid);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(slot)), "l"(src) : "memory");
            }
        };
        auto issue_rows = [&](int t, int j) {
            unsigned char* raw = sm + off_raw0 + (t % RING) * RAW_BYTES;
            if (gatherer) {
                const bool v = j >= 0;
                const float* src = x12 + (size_t)(v ? j : 0) * 12;
                unsigned char* dst = raw + ptid * 48;
                cp16(dst, src, v); cp16(dst + 16, src + 4, v); cp16(dst + 32, src + 8, v);      // !v: zero fill -> pad lane 11 = 0 marks the edge invalid
            }
            if (pgatherer) {
                const long long pt = g_begin / KNN + (long long)t * PTS + ptid;
                const bool v = t < ntiles && pt < (long long)p_end;
                const float* src = x12 + (size_t)(v ? pt : 0) * 12;
                unsigned char* dst = raw + TE * 48 + ptid * 48;
                cp16(dst, src, v); cp16(dst + 16, src + 4, v); cp16(dst + 32, src + 8, v);
            }
            issue_index_copy(t + RING - 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto staged_index = [&](int t) -> int {
            if (!edge_in_range(t)) return -1;
            return *reinterpret_cast<const volatile int*>(sm + off_raw0 + (t % RING) * RAW_BYTES + TE * 48 + PTS * 48 + TE * 4 + ptid * 4);
        };
#pragma unroll 1
        for (int tt = 0; tt < RING - 1; ++tt) issue_rows(tt, edge_in_range(tt) ? __ldg(knn + (g_begin + (long long)tt * TE + ptid)) : -1);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            asm volatile("cp.async.wait_group %0;" :: "n"(RING - 2) : "memory");
            asm volatile("bar.sync 1, %0;" :: "n"(PROD_THREADS) : "memory");
            issue_rows(t + RING - 1, staged_index(t + RING - 1));
            const unsigned char* raw = sm + off_raw0 + (t % RING) * RAW_BYTES;
            if (eb < PTS) {                              // lane (part, eb): centre row of point eb of the tile for my chunk c4
                const float4* ri = reinterpret_cast<const float4*>(raw + TE * 48 + eb * 48);
                const float4 a0 = ri[0], a1 = ri[1], a2 = ri[2];
                const float xi[CD] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
                float2 c01 = make_float2(bias.x, bias.y), c23 = make_float2(bias.z, bias.w);
#pragma unroll
                for (int q = 0; q < CD; ++q) {
                    const float4 w = wc[q];
                    const float2 xx = make_float2(xi[q], xi[q]);
                    ffma2(c01, make_float2(w.x, w.y), xx);
                    ffma2(c23, make_float2(w.z, w.w), xx);
                }
                *reinterpret_cast<float4*>(ctr + (eb * 4 + part) * 16) = make_float4(c01.x, c01.y, c23.x, c23.y);
            }
            __syncwarp();
            mbar_wait(&bar_hempty[st], ph ^ 1u);
            mbar_wait(&bar_eempty[t % ERING], ((uint32_t)(t / ERING) & 1u) ^ 1u);
            unsigned char* dst_hi = sm + off_h0 + st * HSTAGE_BYTES;
            unsigned char* dst_lo = dst_hi + TILE_BYTES;
            unsigned char* dst_ech = sm + off_e0 + (t % ERING) * ESTAGE_BYTES;
            unsigned char* dst_ecl = dst_ech + EC_BYTES;
            unsigned char* dst_mt = dst_ecl + EC_BYTES;
            struct Blk { float4 b0, b1, b2, a0, a1, a2, cc; float ej, ei; };      // ej / ei: the entries of x_j / x_i my EC row is made of
            const uint32_t raw_s = smem_u32(raw), ctr_s = smem_u32(ctr);
            auto load_blk = [&](int blk, Blk& B) {
                const int er = blk * 8 + eb;
                const int pt = (er * 205) >> 12;        // er / KNN, exact for er < 1039
                const uint32_t rj = raw_s + (uint32_t)(er * 48);
                const uint32_t ri = raw_s + (uint32_t)(TE * 48 + pt * 48);
                B.b0 = lds128_ordered(rj); B.b1 = lds128_ordered(rj + 16); B.b2 = lds128_ordered(rj + 32);
                B.a0 = lds128_ordered(ri); B.a1 = lds128_ordered(ri + 16); B.a2 = lds128_ordered(ri + 32);
                B.cc = lds128_ordered(ctr_s + (uint32_t)((pt * 4 + part) * 16));
                B.ej = lds32_ordered(rj + ec_qj * 4);
                B.ei = lds32_ordered(ri + ec_qi * 4);
            };
            auto finish_blk = [&](int blk, const Blk& B) {
                const int er = blk * 8 + eb;
                const float ev[CD] = {B.b0.x - B.a0.x, B.b0.y - B.a0.y, B.b0.z - B.a0.z, B.b0.w - B.a0.w, B.b1.x - B.a1.x, B.b1.y - B.a1.y,
                                      B.b1.z - B.a1.z, B.b1.w - B.a1.w, B.b2.x - B.a2.x};
                const bool valid = B.b2.w != 0.f;        // pad lane 11: 1 in a gathered row, 0 in a zero-filled one
                float2 y01 = make_float2(B.cc.x, B.cc.y), y23 = make_float2(B.cc.z, B.cc.w);
#pragma unroll
                for (int q = 0; q < ((SGB_ABL & 1) ? 1 : CD); ++q) {
                    const float2 ee = make_float2(ev[q], ev[q]);
                    ffma2(y01, w01[q], ee);
                    ffma2(y23, w23[q], ee);
                }
                uint32_t bits = 0;
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    bits = (y01.x > 0.f ? 1u : 0u) | (y01.y > 0.f ? 2u : 0u) | (y23.x > 0.f ? 4u : 0u) | (y23.y > 0.f ? 8u : 0u);
                    y = make_float4(lrelu(y01.x), lrelu(y01.y), lrelu(y23.x), lrelu(y23.y));
                }
                const float4 hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                const float4 lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
                const uint32_t off = (uint32_t)c4 * (TE * 16) + (uint32_t)(er >> 3) * 128 + (uint32_t)(er & 7) * 16;
                *reinterpret_cast<float4*>(dst_hi + off) = hi;
                *reinterpret_cast<float4*>(dst_lo + off) = lo;
                dst_mt[er * 16 + c4] = (unsigned char)bits;     // sign bits of hidden channels 4 c4 .. 4 c4 + 3
                // EC column of this edge (centred edge vector, then 1 for a real edge -> A0), spread over the 16 threads of the edge:
                // thread c4 writes row c4, threads 0..2 also rows 16..18.  The two entries row c4 is made of were fetched with the
                // block's rows (load_blk), rows 16 / 17 are x_i[7] / x_i[8] (already in registers), the centres e-bar are per-thread
                // constants: no shared-memory load sits between the arithmetic and these stores any more (ncu source view: 11 % of
                // the producers' samples waited on them).
                if (!(SGB_ABL & 16)) {
                    const float vm = valid ? 1.f : 0.f;
                    {
                        const float v = ((c4 < 9 ? B.ej - B.ei : B.ei) - ebar0) * vm;
                        const float vh = tf32_hi(v);
                        const uint32_t o = kmajor_off(c4, er, TN);
                        *reinterpret_cast<float*>(dst_ech + o) = vh;
                        *reinterpret_cast<float*>(dst_ecl + o) = tf32_hi(v - vh);
                    }
                    if (c4 < 3) {
                        const float x = c4 == 0 ? B.a1.w : B.a2.x;                     // x_i[7], x_i[8]
                        const float v = (c4 == 2 ? 1.f : x - ebar1) * vm;
                        const float vh = tf32_hi(v);
                        const uint32_t o = kmajor_off(16 + c4, er, TN);
                        *reinterpret_cast<float*>(dst_ech + o) = vh;
                        *reinterpret_cast<float*>(dst_ecl + o) = tf32_hi(v - vh);
                    }
                }
            };
            Blk A, Bq;
            int blk = grp;
            load_blk(blk, A);
#pragma unroll 1
            for (; blk + 2 < BLOCKS; blk += 4) {
                load_blk(blk + 2, Bq);
                finish_blk(blk, A);
                if (blk + 4 < BLOCKS) load_blk(blk + 4, A);
                finish_blk(blk + 2, Bq);
            }
            if (blk < BLOCKS) finish_blk(blk, A);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[st]);
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer (one thread): an event loop over two independent streams of work — the z MMAs of the next
        // tile whose H stage is full, and the T MMAs of the next half tile whose dv1 operand the epilogue has finished — so that
        // neither waits for the other's inputs (non-blocking mbarrier tests)
        // The whole warp walks the loop with warp-uniform control flow (barrier tests agreed on by vote) and ONE elected lane issues:
        // under `if (lane == 0)` the compiler wrapped every tcgen05.mma / commit in its own ELECT + BRA.U.ANY loop.
        {
            const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
            const uint32_t idesc_t = make_idesc_tf32(64, TN, false, false);
            const uint32_t a_hi = smem_u32(sm + off_bm_hi), a_lo = smem_u32(sm + off_bm_lo);
            // Descriptors are loop-invariant up to the start-address field: bases once, a K step = an add of (bytes >> 4) to the low word
            // (shared-memory addresses < 256 KB: the 14-bit field cannot carry).  The issuing thread is ONE instruction stream; at 40 tensor
            // cycles per instruction (N = 80) the descriptor arithmetic between two tcgen05.mma was the pacing item of the kernel.
            auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
            const uint64_t dBh = make_desc(a_hi, COUT * 16, 128), dBl = make_desc(a_lo, COUT * 16, 128);
            uint64_t dHh[2], dHl[2], dDh[2], dDl[2], dEh[ERING], dEl[ERING];
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                const uint32_t hh = smem_u32(sm + off_h0 + b2 * HSTAGE_BYTES), dv = smem_u32(sm + off_dv + b2 * 2 * DVH_BYTES);
                dHh[b2] = make_desc(hh, TE * 16, 128); dHl[b2] = make_desc(hh + TILE_BYTES, TE * 16, 128);
                dDh[b2] = make_desc(dv, COUT * 16, 128); dDl[b2] = make_desc(dv + DVH_BYTES, COUT * 16, 128);
            }
#pragma unroll
            for (int b3 = 0; b3 < ERING; ++b3) {
                const uint32_t ec = smem_u32(sm + off_e0 + b3 * ESTAGE_BYTES);
                dEh[b3] = make_desc(ec, TN * 16, 128); dEl[b3] = make_desc(ec + EC_BYTES, TN * 16, 128);
            }
            int nz = 0, nT = 0;                         // next tile for z, next HALF tile (2 u + h) for T
            while (nT < 2 * ntiles) {
                bool did = false;
                if (nz < ntiles) {
                    const int st = nz & 1;
                    const uint32_t ph = (uint32_t)(nz >> 1) & 1u;
                    const bool ok = mbar_test(&bar_full[st], ph) && mbar_test(&bar_tempty[st], ph ^ 1u);
                    if (__all_sync(SGB_FULL_MASK, ok)) {
                        fence_after_sync();
                        const uint64_t hh = st ? dHh[1] : dHh[0], hl = st ? dHl[1] : dHl[0];
                        const uint32_t d = tmem + (uint32_t)(st * Z_COL);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int i = 0; i < COUT / 8; ++i) {
                                const uint64_t dah = adv(dBh, 2 * i * (COUT * 16)), dal = adv(dBl, 2 * i * (COUT * 16));
                                const uint64_t dbh = adv(hh, 2 * i * (TE * 16)), dbl = adv(hl, 2 * i * (TE * 16));
                                mma_tf32(d, dah, dbh, idesc, i > 0);
                                if (!(SGB_ABL & 4)) {
                                    mma_tf32(d, dal, dbh, idesc, true);
                                    mma_tf32(d, dah, dbl, idesc, true);
                                }
                            }
                            mma_commit(&bar_tfull[st]);
                            mma_commit(&bar_hempty[st]);
                        }
                        __syncwarp();
                        ++nz;
                        did = true;
                    }
                }
                {
                    const int u = nT >> 1, h = nT & 1;
                    const int seg = u / FLUSH, gb = seg & 1;
                    bool ok = mbar_test(&bar_dvfull[h], (uint32_t)u & 1u);
                    if (ok && h == 0 && u % FLUSH == 0) ok = mbar_test(&bar_gempty[gb], ((uint32_t)(seg >> 1) & 1u) ^ 1u);
                    if (__all_sync(SGB_FULL_MASK, ok)) {
                        fence_after_sync();
                        const int er = u % ERING;
                        const uint64_t eh0 = er == 0 ? dEh[0] : (er == 1 ? dEh[1] : dEh[2]), el0 = er == 0 ? dEl[0] : (er == 1 ? dEl[1] : dEl[2]);
                        const uint64_t eh = adv(eh0, (uint32_t)(h * (TE / 8)) * (TN * 16)), el = adv(el0, (uint32_t)(h * (TE / 8)) * (TN * 16));
                        const uint64_t vh = h ? dDh[1] : dDh[0], vl = h ? dDl[1] : dDl[0];
                        const uint32_t d = tmem + (uint32_t)(T_COL0 + gb * T_COLS);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int i = 0; i < ((SGB_ABL & 8) ? 1 : TE / 16); ++i) {           // 8 edges (K) per instruction = 2 chunks of 4; 5 per half tile
                                const uint64_t dah = adv(vh, 2 * i * (COUT * 16)), dal = adv(vl, 2 * i * (COUT * 16));
                                const uint64_t dbh = adv(eh, 2 * i * (TN * 16)), dbl = adv(el, 2 * i * (TN * 16));
                                mma_tf32(d, dah, dbh, idesc_t, (u % FLUSH) > 0 || h > 0 || i > 0);
                                if (!(SGB_ABL & 4)) {
                                    mma_tf32(d, dal, dbh, idesc_t, true);
                                    mma_tf32(d, dah, dbl, idesc_t, true);
                                }
                            }
                            mma_commit(&bar_dvempty[h]);
                            if (h == 1) {
                                mma_commit(&bar_eempty[u % ERING]);
                                if ((u + 1) % FLUSH == 0 || u == ntiles - 1) mma_commit(&bar_gfull[gb]);
                            }
                        }
                        __syncwarp();
                        ++nT;
                        did = true;
                    }
                }
                if (!did) __nanosleep(32);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: channel c = 16 warp + lane on lanes 0..15 (the rows of an M = 64 accumulator); z -> dv1 -> operand tile
        const int c = warp * 16 + (lane & 15);
        const bool owner = lane < 16;
        const float rc = __ldg(coef + COUT * COUT + c);
        const uint32_t mbyte = (uint32_t)(c >> 2), bit = (uint32_t)(c & 3);
        // T / A0 accumulators of my channel across the CTA's segments: unevaluated fp32 pairs in registers (two-sum), converted to
        // fp64 once at the end (round 2a kept fp64 accumulators in 10 KB of shared memory; DADD is slow on this part)
        float acc_h[ACCN], acc_l[ACCN];
#pragma unroll
        for (int q = 0; q < ACCN; ++q) { acc_h[q] = 0.f; acc_l[q] = 0.f; }
        auto two_sum = [](float& h, float& l, float x) {
            const float s_ = h + x;
            const float bb = s_ - h;
            l += (h - (s_ - bb)) + (x - bb);
            h = s_;
        };
        unsigned char* dv_row = sm + off_dv + (c >> 3) * 128 + (c & 7) * 16;          // my row in a [64][40] K-major half tile
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_tfull[st], ph);
            fence_after_sync();
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * Z_COL);
            const unsigned char* s_m = sm + off_e0 + (t % ERING) * ESTAGE_BYTES + 2 * EC_BYTES;
#pragma unroll 1
            for (int pp = 0; pp < PTS; ++pp) {
                const int h = pp >> 1;                                  // half of the dv1 tile this point belongs to
                if ((pp & 1) == 0) mbar_wait(&bar_dvempty[h], ((uint32_t)t & 1u) ^ 1u);      // the T MMAs of tile t - 1 are done with this half
                uint32_t v[16], u[4];
                tmem_ld16_issue(taddr + (uint32_t)(pp * KNN), v);
                tmem_ld4_issue(taddr + (uint32_t)(pp * KNN + 16), u);
                tmem_ld_wait();
                tmem_ld_pin20(v, u);
                float dv[KNN];
#pragma unroll
                for (int k = 0; k < ((SGB_ABL & 2) ? 1 : KNN); ++k) {
                    const int e = pp * KNN + k;
                    const float z = __uint_as_float(k < 16 ? v[k] : u[k - 16]);
                    const bool posv = (s_m[e * 16 + mbyte] >> bit) & 1u;
                    dv[k] = (rc - z) * (posv ? 1.f : SLOPE);            // padding edges: their EC column is zero, so any finite value is inert
                }
                if (owner) {
#pragma unroll
                    for (int q = 0; q < ((SGB_ABL & 2) ? 0 : KNN / 4); ++q) {                   // 4 consecutive edges = one 16-byte chunk of row c
                        const float4 d4 = make_float4(dv[4 * q], dv[4 * q + 1], dv[4 * q + 2], dv[4 * q + 3]);
                        const float4 hi = make_float4(tf32_hi(d4.x), tf32_hi(d4.y), tf32_hi(d4.z), tf32_hi(d4.w));
                        const float4 lo = make_float4(tf32_hi(d4.x - hi.x), tf32_hi(d4.y - hi.y), tf32_hi(d4.z - hi.z), tf32_hi(d4.w - hi.w));
                        const uint32_t o = (uint32_t)(h * 2 * DVH_BYTES) + (uint32_t)((pp & 1) * (KNN / 4) + q) * (COUT * 16);
                        *reinterpret_cast<float4*>(dv_row + o) = hi;
                        *reinterpret_cast<float4*>(dv_row + o + DVH_BYTES) = lo;
                    }
                }
                if (pp & 1) {                                           // half complete: hand it to the tensor cores
                    fence_async_smem();                                 // generic-proxy writes -> tcgen05.mma operand reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_dvfull[h]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[st]);
            if ((t + 1) % FLUSH == 0 || t == ntiles - 1) {
                // flush the finished T segment into the fp64 accumulators (thread = accumulator row)
                const int seg = t / FLUSH, gb = seg & 1;
                mbar_wait(&bar_gfull[gb], (uint32_t)(seg >> 1) & 1u);
                fence_after_sync();
                const uint32_t gaddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(T_COL0 + gb * T_COLS);
                float g16[16], g4a[4], g4b[4];
                tmem_ld16(gaddr, g16);
                tmem_ld4(gaddr + 16, g4a);
                tmem_ld4(gaddr + 20, g4b);
#pragma unroll
                for (int q = 0; q < 16; ++q) two_sum(acc_h[q], acc_l[q], g16[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q) two_sum(acc_h[16 + q], acc_l[16 + q], g4a[q]);      // columns 20..23 are structurally zero
                (void)g4b;
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_gempty[gb]);
            }
        }
        // part[cta][c][0] = A0, [1] = sum dv1 zhat1 = invstd1 * W1[c,:].T[c,:], [2 + t] = T[c][t]
        if (owner) {
            double* dst = part + ((size_t)blockIdx.x * COUT + c) * NACC;
            double dot = 0.0;
#pragma unroll
            for (int i = 0; i < CIN; ++i) {
                const double a_i = (double)acc_h[i] + (double)acc_l[i];
                dst[2 + i] = a_i;
                dot += (double)__ldg(W1 + c * CIN + i) * a_i;
            }
            dst[0] = (double)acc_h[CIN] + (double)acc_l[CIN];
            dst[1] = (double)stats1[64 + c] * dot;
        }
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

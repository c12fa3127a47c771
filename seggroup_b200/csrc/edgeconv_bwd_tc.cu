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

#ifdef SGB_ROLE_CLOCKS
#include <cstdio>
#define WAITC(i, x) do { const long long t0_ = clock64(); x; wclk[i] += clock64() - t0_; } while (0)
#else
#define WAITC(i, x) x
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
constexpr int Z_COL = 80;                                      // TMEM: z accumulators at columns 0 and 80
constexpr int T_COL0 = 160, T_COLS = 32;                       //       T accumulators at 160 and 192 (24 columns used)
constexpr int D1_COL = 224;                                    //       first-layer accumulators at 224 and 288 (64 columns each)
constexpr int BM_COL = 352;                                    //       Bm hi | lo as the A operand of the z MMAs (2 x 64 columns)
constexpr int KE = 24;                                         // K of the first layer: 9 differences + 9 centre + 1 + 5 zero
constexpr int E_BYTES = TE * 20 * 4;                           // one K-major E tile (hi or lo): K chunks 0..4; K = 20..23 is one shared chunk of zeros
constexpr int EZ_BYTES = 128 * 16;
constexpr int W1_BYTES = COUT * KE * 4;

constexpr int off_w1_hi = 0;                                   // W1s = (scale1 W1 | folded BN1 bias | 0), hi / lo
constexpr int off_w1_lo = off_w1_hi + W1_BYTES;
constexpr int off_h0 = off_w1_lo + W1_BYTES;                   // 2 H stages
constexpr int off_bm_hi = off_h0;                              // setup only (aliases the H stages): Bm hi / lo and the identity that moves them to TMEM
constexpr int off_bm_lo = off_bm_hi + BM_BYTES;
constexpr int off_ident = off_bm_lo + BM_BYTES;
constexpr int off_e0 = off_h0 + 2 * HSTAGE_BYTES;              // 3 EC / mask stages
constexpr int off_dv = off_e0 + ERING * ESTAGE_BYTES;          // [half][hi, lo]
constexpr int off_et0 = off_dv + 4 * DVH_BYTES;                // E tiles: [buffer 0: hi, lo][buffer 1: hi, lo][zero chunk]
constexpr int off_ezero = off_et0 + 4 * E_BYTES;
constexpr int off_raw0 = off_ezero + EZ_BYTES;                 // (the M = 128 read of K chunk 4 runs 48 rows into whatever follows the tile)
constexpr int off_ebar = off_raw0 + RING * RAW_BYTES;
constexpr int off_bars = off_ebar + 32 * 4;                    // 27 mbarriers
constexpr int off_tmem_slot = off_bars + 28 * 8;
constexpr int SMEM_TOTAL = off_tmem_slot + 16;
static_assert(HSTAGE_BYTES % 128 == 0 && ESTAGE_BYTES % 128 == 0 && DVH_BYTES % 128 == 0, "stage alignment");
static_assert(SMEM_TOTAL + 128 <= 227 * 1024, "shared memory budget");
static_assert(3 * BM_BYTES <= 2 * HSTAGE_BYTES, "setup tiles fit into the H stages");
static_assert(off_et0 % 16 == 0 && E_BYTES % 16 == 0 && off_bars % 8 == 0, "alignment");

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
    uint64_t* bar_xfull = bar_full + 19;                                     // [2] producers -> MMA: E tile written
    uint64_t* bar_xempty = bar_full + 21;                                    // [2] MMA (commit after the first-layer MMAs) -> producers
    uint64_t* bar_d1full = bar_full + 23;                                    // [2] MMA (same commit) -> producers: D1 complete
    uint64_t* bar_d1empty = bar_full + 25;                                   // [2] producers -> MMA: D1 read into registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + off_tmem_slot);
    float* s_ebar = reinterpret_cast<float*>(sm + off_ebar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SGB_ROLE_CLOCKS
    long long wclk[32]; for (int i = 0; i < 32; ++i) wclk[i] = 0; const long long tstart = clock64();
#endif

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
    for (int i = tid; i < COUT * COUT; i += THREADS) {                        // identity: Bm reaches TMEM as Bm * I (edgeconv_tc.cu)
        const int n = i / COUT, j = i % COUT;
        *reinterpret_cast<float*>(sm + off_ident + tile_off(n, j, COUT)) = (n == j) ? 1.f : 0.f;
    }
    for (int i = tid; i < COUT * KE; i += THREADS) {                          // W1s [c][k]: BN1 scale folded in, column 18 = folded bias
        const int c = i / KE, k = i % KE;
        const float sc = stats1[128 + c];
        float w = 0.f;
        if (k < CIN) w = __ldg(W1 + c * CIN + k) * sc;
        else if (k == CIN) w = fmaf(-sc, stats1[c], stats1[192 + c]);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, k, COUT);
        *reinterpret_cast<float*>(sm + off_w1_hi + off) = hi;
        *reinterpret_cast<float*>(sm + off_w1_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < 128; i += THREADS) *reinterpret_cast<float4*>(sm + off_ezero + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
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
        for (int b2 = 0; b2 < 2; ++b2) {
            mbar_init(&bar_xfull[b2], PROD_WARPS); mbar_init(&bar_xempty[b2], 1);
            mbar_init(&bar_d1full[b2], 1); mbar_init(&bar_d1empty[b2], PROD_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp > MMA_WARP) {
        // ================= producers, as ec2_tc1_kernel (edgeconv_tc.cu): the first layer runs on the tensor cores,
        //     D1[128 edge rows, 64 channels] = E[128, 24] * W1s[64, 24]^T,   E = (x_j - x_i, x_i, valid, 0...),
        // and a producer thread owns one edge ROW of D1: sign bits, LeakyReLU, hi / lo split, 16-byte stores into the K-major H tile.
        // Per tile:  [gather ring] -> E(t) (one thread per edge) and EC(t) (one thread per (row, 4 edges)) -> H(t - 1), mask(t - 1).
        // Round 2 evaluated the first layer on the CUDA cores, 16 threads per edge: the producers were busy 96 % of the kernel
        // (cycle counters around every wait, -DSGB_ROLE_CLOCKS) with everybody else waiting for them.
        const int pw = warp - (MMA_WARP + 1);           // 0..7
        const int ptid = pw * 32 + lane;                // gather / E role: edge row `ptid` (< TE), point row `ptid` (< PTS)
        const int quad = warp & 3;                      // the TMEM lanes this warp may read: 32 quad .. 32 quad + 31
        const int chh = pw >> 2;                        // H role: channels 32 chh .. 32 chh + 31 of edge row 32 quad + lane
        const int hrow = quad * 32 + lane;
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
        auto split4 = [](const float4 y, float4& hi, float4& lo) {
            hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
            lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
        };
        const uint32_t d1addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D1_COL + 32 * chh);
#pragma unroll 1
        for (int tt = 0; tt < RING - 1; ++tt) issue_rows(tt, edge_in_range(tt) ? __ldg(knn + (g_begin + (long long)tt * TE + ptid)) : -1);
#pragma unroll 1
        for (int t = 0; t <= ntiles; ++t) {
            if (t < ntiles) {
                WAITC(30, asm volatile("cp.async.wait_group %0;" :: "n"(RING - 2) : "memory"));
                WAITC(31, asm volatile("bar.sync 1, %0;" :: "n"(PROD_THREADS) : "memory"));
                issue_rows(t + RING - 1, staged_index(t + RING - 1));
                const unsigned char* raw = sm + off_raw0 + (t % RING) * RAW_BYTES;
                // ---- E(t): the edge vectors of tile t as the A operand of the first layer
                const int xb = t & 1;
                WAITC(0, mbar_wait(&bar_xempty[xb], (((uint32_t)t >> 1) & 1u) ^ 1u));
                if (gatherer) {
                    unsigned char* e_hi = sm + off_et0 + xb * (2 * E_BYTES);
                    unsigned char* e_lo = e_hi + E_BYTES;
                    const int pt = (ptid * 205) >> 12;      // ptid / KNN
                    const float4* rj = reinterpret_cast<const float4*>(raw + ptid * 48);
                    const float4* ri = reinterpret_cast<const float4*>(raw + TE * 48 + pt * 48);
                    const float4 b0 = rj[0], b1 = rj[1], b2 = rj[2], a0 = ri[0], a1 = ri[1], a2 = ri[2];
                    const float one = b2.w != 0.f ? 1.f : 0.f;      // 1 for a gathered row, 0 for a zero-filled one (edge outside the CTA's range)
                    float4 ch[5];
                    ch[0] = make_float4(b0.x - a0.x, b0.y - a0.y, b0.z - a0.z, b0.w - a0.w);
                    ch[1] = make_float4(b1.x - a1.x, b1.y - a1.y, b1.z - a1.z, b1.w - a1.w);
                    ch[2] = make_float4(b2.x - a2.x, a0.x, a0.y, a0.z);
                    ch[3] = make_float4(a0.w, a1.x, a1.y, a1.z);
                    ch[4] = make_float4(a1.w, a2.x, 1.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const float4 y = make_float4(ch[q].x * one, ch[q].y * one, ch[q].z * one, ch[q].w * one);
                        float4 hi, lo;
                        split4(y, hi, lo);
                        *reinterpret_cast<float4*>(e_hi + q * (TE * 16) + ptid * 16) = hi;
                        *reinterpret_cast<float4*>(e_lo + q * (TE * 16) + ptid * 16) = lo;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_xfull[xb]);
            }
            if (t >= 1) {
                // ---- H(u), mask(u): hidden activations and their sign bits of tile u = t - 1 from the first-layer accumulator
                const int u = t - 1, su = u & 1;
                WAITC(2, mbar_wait(&bar_d1full[su], ((uint32_t)u >> 1) & 1u));
                fence_after_sync();
                uint32_t va[16], vb[16];
                tmem_ld16_issue(d1addr + (uint32_t)(su * 64), va);
                tmem_ld16_issue(d1addr + (uint32_t)(su * 64 + 16), vb);
                tmem_ld_wait();
                tmem_ld_pin16(va);
                tmem_ld_pin16(vb);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_d1empty[su]);
                WAITC(3, mbar_wait(&bar_hempty[su], (((uint32_t)(u >> 1)) & 1u) ^ 1u));
                unsigned char* dst_hi = sm + off_h0 + su * HSTAGE_BYTES;
                unsigned char* dst_lo = dst_hi + TILE_BYTES;
                unsigned char* dst_mt = sm + off_e0 + (u % ERING) * ESTAGE_BYTES + 2 * EC_BYTES;      // this stage was acquired with EC(u)
                const bool rowv = hrow < TE;
                uint32_t mlo = 0, mhi = 0;              // sign bits: one byte per 4-channel chunk (a zero row — edge out of range — gives 0 and h = 0)
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    float y[4];
                    uint32_t bits = 0;
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const int i = 4 * i4 + s4;
                        const float pre = __uint_as_float(i < 16 ? va[i] : vb[i - 16]);
                        bits |= pre > 0.f ? (1u << s4) : 0u;
                        y[s4] = lrelu(pre);
                    }
                    if (i4 < 4) mlo |= bits << (8 * i4); else mhi |= bits << (8 * (i4 - 4));
                    float4 hi, lo;
                    split4(make_float4(y[0], y[1], y[2], y[3]), hi, lo);
                    if (rowv) {
                        const uint32_t off = (uint32_t)(8 * chh + i4) * (TE * 16) + (uint32_t)hrow * 16;
                        *reinterpret_cast<float4*>(dst_hi + off) = hi;
                        *reinterpret_cast<float4*>(dst_lo + off) = lo;
                    }
                }
                if (rowv) *reinterpret_cast<uint2*>(dst_mt + hrow * 16 + 8 * chh) = make_uint2(mlo, mhi);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[su]);
            }
            if (t < ntiles) {
                // ---- EC(t), LAST (its stage is freed by the T MMAs of tile t - 3: waiting here does not hold back H(t - 1)): the centred edge vectors with the edges along K, [24 rows][80 edges]: one work item = (row r < 19, 4 consecutive
                // edges — one point, one 16-byte chunk), consecutive lanes = consecutive rows (conflict-free loads and stores)
                const unsigned char* raw = sm + off_raw0 + (t % RING) * RAW_BYTES;
                WAITC(1, mbar_wait(&bar_eempty[t % ERING], ((uint32_t)(t / ERING) & 1u) ^ 1u));
                {
                    unsigned char* dst_ech = sm + off_e0 + (t % ERING) * ESTAGE_BYTES;
                    unsigned char* dst_ecl = dst_ech + EC_BYTES;
#pragma unroll 1
                    for (int it = ptid; it < 19 * (TE / 4); it += PROD_THREADS) {
                        const int g = it / 19, r = it - g * 19;
                        const int e0i = 4 * g, pt = g / (KNN / 4);
                        const float* rj = reinterpret_cast<const float*>(raw + e0i * 48);
                        const float* ri = reinterpret_cast<const float*>(raw + TE * 48 + pt * 48);
                        const float eb_ = r < CIN ? s_ebar[r] : 0.f;
                        const float xi = r < CIN ? ri[r < 9 ? r : r - 9] : 0.f;
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float vm = rj[q * 12 + 11] != 0.f ? 1.f : 0.f;
                            const float ev = r < 9 ? rj[q * 12 + r] - xi : (r < CIN ? xi : 1.f);
                            v[q] = (ev - eb_) * vm;
                        }
                        float4 hi, lo;
                        split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
                        const uint32_t o = (uint32_t)g * (TN * 16) + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
                        *reinterpret_cast<float4*>(dst_ech + o) = hi;
                        *reinterpret_cast<float4*>(dst_ecl + o) = lo;
                    }
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
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
            // Bm (hi, lo) into TMEM in the layout of an M = 64 accumulator: D = Bm * I (exact).  The tensor pipe runs in issue order, so every
            // later MMA — and the first write into H stage 0, which waits for the first-layer MMA of tile 0 — comes after these.
            if (elect_one_sync()) {
                const uint32_t idw = make_idesc_tf32(64, COUT, false, false);
                const uint32_t idn = smem_u32(sm + off_ident);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                    for (int i = 0; i < COUT / 8; ++i) {
                        const uint32_t o = (uint32_t)(2 * i) * (COUT * 16);
                        mma_tf32(tmem + (uint32_t)(BM_COL + h2 * 64), make_desc((h2 ? a_lo : a_hi) + o, COUT * 16, 128), make_desc(idn + o, COUT * 16, 128), idw, i > 0);
                    }
            }
            __syncwarp();
            const uint32_t idesc1 = make_idesc_tf32(128, COUT, false, false);
            const uint32_t ez = smem_u32(sm + off_ezero);
            const uint64_t dW1h = make_desc(smem_u32(sm + off_w1_hi), COUT * 16, 128), dW1l = make_desc(smem_u32(sm + off_w1_lo), COUT * 16, 128);
            uint64_t dXh[2], dXl[2], dXh2[2], dXl2[2];
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                const uint32_t eh = smem_u32(sm + off_et0 + b2 * (2 * E_BYTES)), el = eh + E_BYTES;
                dXh[b2] = make_desc(eh, TE * 16, 128);
                dXl[b2] = make_desc(el, TE * 16, 128);
                dXh2[b2] = make_desc(eh + 4 * (TE * 16), ez - (eh + 4 * (TE * 16)), 128);      // K chunk 4 paired with the shared zero chunk
                dXl2[b2] = make_desc(el + 4 * (TE * 16), ez - (el + 4 * (TE * 16)), 128);
            }
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
            int n1 = 0, nz = 0, nT = 0;                 // next tile for the first layer, for z, next HALF tile (2 u + h) for T
            while (nT < 2 * ntiles) {
                bool did = false;
                if (n1 < ntiles) {
                    const int xb = n1 & 1;
                    const uint32_t ph1 = ((uint32_t)n1 >> 1) & 1u;
                    const bool ok = mbar_test(&bar_xfull[xb], ph1) && mbar_test(&bar_d1empty[xb], ph1 ^ 1u);
                    if (__all_sync(SGB_FULL_MASK, ok)) {
                        fence_after_sync();
                        const uint64_t eh = xb ? dXh[1] : dXh[0], el = xb ? dXl[1] : dXl[0], eh2 = xb ? dXh2[1] : dXh2[0], el2 = xb ? dXl2[1] : dXl2[0];
                        const uint32_t d1 = tmem + (uint32_t)(D1_COL + xb * 64);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int i = 0; i < KE / 8; ++i) {
                                const uint64_t dah = i < 2 ? adv(eh, 2 * i * (TE * 16)) : eh2, dal = i < 2 ? adv(el, 2 * i * (TE * 16)) : el2;
                                const uint64_t dbh = adv(dW1h, 2 * i * (COUT * 16)), dbl = adv(dW1l, 2 * i * (COUT * 16));
                                mma_tf32(d1, dah, dbh, idesc1, i > 0);
                                mma_tf32(d1, dal, dbh, idesc1, true);
                                mma_tf32(d1, dah, dbl, idesc1, true);
                            }
                            mma_commit(&bar_d1full[xb]);
                            mma_commit(&bar_xempty[xb]);
                        }
                        __syncwarp();
                        ++n1;
                        did = true;
                    }
                }
                if (nz < ntiles) {
                    // (issuing the 24 z instructions in 4 groups with the other streams polled in between measured SLOWER, 0.72 -> 0.88 ms per
                    // entry at 150k points, as did a round-robin of the three products in the forward kernel: batches stay whole)
                    const int st = nz & 1;
                    const uint32_t ph = (uint32_t)(nz >> 1) & 1u;
                    const bool ok = mbar_test(&bar_full[st], ph) && mbar_test(&bar_tempty[st], ph ^ 1u);
                    if (__all_sync(SGB_FULL_MASK, ok)) {
                        fence_after_sync();
                        const uint64_t hh = st ? dHh[1] : dHh[0], hl = st ? dHl[1] : dHl[0];
                        const uint32_t d = tmem + (uint32_t)(st * Z_COL), wb = tmem + (uint32_t)BM_COL;
                        if (elect_one_sync()) {
#pragma unroll
                            for (int i = 0; i < COUT / 8; ++i) {                           // A = Bm from TMEM: 8 columns (K) per instruction
                                const uint64_t dbh = adv(hh, 2 * i * (TE * 16)), dbl = adv(hl, 2 * i * (TE * 16));
                                mma_tf32_ts(d, wb + (uint32_t)(8 * i), dbh, idesc, i > 0);
                                if (!(SGB_ABL & 4)) {
                                    mma_tf32_ts(d, wb + (uint32_t)(64 + 8 * i), dbh, idesc, true);
                                    mma_tf32_ts(d, wb + (uint32_t)(8 * i), dbl, idesc, true);
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
                if (!did) { WAITC(29, __nanosleep(32)); }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: channel c = 16 warp + lane on lanes 0..15 (the rows of an M = 64 accumulator); z -> dv1 -> operand tile
        const int c = warp * 16 + (lane & 15);
        const bool owner = lane < 16;
        const float rc = __ldg(coef + COUT * COUT + c);
        const uint32_t mbyte = (uint32_t)(c >> 2), bit = (uint32_t)(c & 3);
        const int hp = lane >> 4;
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
            WAITC(2, mbar_wait(&bar_tfull[st], ph));
            fence_after_sync();
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * Z_COL);
            const unsigned char* s_m = sm + off_e0 + (t % ERING) * ESTAGE_BYTES + 2 * EC_BYTES;
            // An M = 64 accumulator keeps its rows on lanes 0..15 of the quadrant: the .16x32bx2 loads hand the upper half-warp the SECOND
            // point of a pair (20 columns further) of the same 16 channels, so one pass of the loop body turns a whole half tile into dv1.
#pragma unroll 1
            for (int h = 0; h < PTS / 2; ++h) {                         // half of the dv1 tile = points 2 h (lanes 0..15) and 2 h + 1 (lanes 16..31)
                WAITC(3, mbar_wait(&bar_dvempty[h], ((uint32_t)t & 1u) ^ 1u));      // the T MMAs of tile t - 1 are done with this half
                const int pp = 2 * h + hp;
                uint32_t v[16], u[4];
                tmem_ld16_halves_issue<KNN>(taddr + (uint32_t)(2 * h * KNN), v);
                tmem_ld4_halves_issue<KNN>(taddr + (uint32_t)(2 * h * KNN + 16), u);
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
#pragma unroll
                for (int q = 0; q < ((SGB_ABL & 2) ? 0 : KNN / 4); ++q) {                   // 4 consecutive edges = one 16-byte chunk of row c
                    const float4 d4 = make_float4(dv[4 * q], dv[4 * q + 1], dv[4 * q + 2], dv[4 * q + 3]);
                    const float4 hi = make_float4(tf32_hi(d4.x), tf32_hi(d4.y), tf32_hi(d4.z), tf32_hi(d4.w));
                    const float4 lo = make_float4(tf32_hi(d4.x - hi.x), tf32_hi(d4.y - hi.y), tf32_hi(d4.z - hi.z), tf32_hi(d4.w - hi.w));
                    const uint32_t o = (uint32_t)(h * 2 * DVH_BYTES) + (uint32_t)(hp * (KNN / 4) + q) * (COUT * 16);
                    *reinterpret_cast<float4*>(dv_row + o) = hi;
                    *reinterpret_cast<float4*>(dv_row + o + DVH_BYTES) = lo;
                }
                fence_async_smem();                                     // generic-proxy writes -> tcgen05.mma operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_dvfull[h]);              // half complete: hand it to the tensor cores
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[st]);
            if ((t + 1) % FLUSH == 0 || t == ntiles - 1) {
                // flush the finished T segment into the fp64 accumulators (thread = accumulator row)
                const int seg = t / FLUSH, gb = seg & 1;
                WAITC(4, mbar_wait(&bar_gfull[gb], (uint32_t)(seg >> 1) & 1u));
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
#ifdef SGB_ROLE_CLOCKS
    if (blockIdx.x == 3 && lane == 0) { const long long tot = clock64() - tstart; for (int i = 0; i < 32; ++i) if (wclk[i]) printf("bwd warp %2d wait %2d : %6.1f%%  (total %lld clk, %d tiles)\n", warp, i, 100.0 * wclk[i] / tot, tot, ntiles); }
#endif
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

// a9 on the tensor cores: BOTH EdgeConv layers of MLP3 (seggroup/model.py:121-138) in one warp-specialised tcgen05 kernel.
//
//   z[c, e] = sum_j W2[c, j] h[e, j],   h[e, :] = lrelu(BN1(W1 e_e)),   e = (x_j - x_i, x_i),  j in knn(i)
//
// Per tile of TE edges (TE / 20 points) three products run on the tensor cores, all kind::tf32 with the TF32 x 3 split of tc_common.cuh
// (fp32-level accuracy: the merge decisions downstream are discrete):
//   (1) first layer    D1[128 edge rows, 64 channels] = E[128, 24] * W1s[64, 24]^T      (K = 9 differences + 9 centre + 1 + 5 zero)
//       E = (x_j - x_i, x_i, valid, 0...) is assembled by ONE thread per edge (hi / lo split, five 16-byte stores; K = 20..23 is one
//       shared chunk of zeros reached through the descriptor's leading-dimension offset); W1s = (scale1 W1 | beta1 - scale1 mean1 | 0):
//       the constant-1 column carries the folded BatchNorm-1 bias, an all-zero row (edge outside the CTA's range) gives h = lrelu(0) = 0.
//       Round 2 had this layer on the CUDA cores: 16 producer threads per edge, each re-reading the edge's two raw rows, ~1,700
//       instructions per edge against ~350 now.
//   (2) second layer   D[64 channels, TE edges] += W2[64, 64] * H[TE, 64]^T.  The edges of D1 lie on the TMEM LANES, so a producer thread
//       owns one edge row: tcgen05.ld of 32 of its 64 pre-activations (two warps per TMEM quadrant, one per channel half), LeakyReLU,
//       hi / lo split, 16-byte stores into the K-major H tile (consecutive lanes = consecutive rows: conflict-free).  W2 (hi | lo) is the
//       A operand and lives in TMEM for the whole kernel (written once as W2 * I by the tensor core itself, so it has the layout of an
//       M = 64 accumulator by construction).  Orientation: channels on the TMEM lanes, edges on the columns, so that one epilogue thread owns
//       one channel and both reductions over edges are thread-local:
//         * BatchNorm-2 batch statistics  sum z, sum z^2  (fp32 per tile, two-sum pairs across tiles, fixed order);
//         * max / min over the 20 neighbours of a point.  BN (per-channel affine) followed by LeakyReLU is monotone, so
//               max_k lrelu(BN2(z_k)) = lrelu(BN2(max_k z_k))  if gamma2 >= 0,   lrelu(BN2(min_k z_k))  otherwise:
//           ONE pass over the edges produces the statistics and the (extreme, arg) candidates; a light second kernel applies BN2 + LeakyReLU.
//       An M = 64 accumulator keeps its rows on lanes 0..15 of each TMEM quadrant: the .16x32bx2 load shape gives the upper half-warp
//       the NEXT point (20 columns further) of the same channels, so every lane of the epilogue works.
//   (3) GRAM variant (training): the backward pass needs the second moments of the hidden activations, sum_e h h^T and sum_e h
//       (analytic BatchNorm backward, edgeconv_bwd.cu) — an edge contraction, i.e. the edges must lie along K.  kind::tf32 only walks
//       MN-major operands in one swizzle mode that no K-major layout shares (descriptor probe, round 1), so the producers store h a
//       second time, transposed ([hidden row][edges], K-major SWIZZLE_64B, 32 consecutive edges of one row per store), and
//           G[128, 80] += [H_lo^T ; H_hi^T] (K = edges) x [H_hi^T ; 1 ; 0]^T
//       collects lo*hi (rows 0..63), hi*hi (rows 64..127); the column of ones gives sum h; hi*lo follows by symmetry.  G is flushed to
//       global fp32 slots every FLUSH tiles and summed in fp64 in a fixed order (TMEM accumulates in fp32).
//
// Warp roles (416 threads, 1 CTA / SM, persistent over a contiguous range of points):
//   warps 0-3   epilogue (one per TMEM quadrant);
//   warp  4     TMEM allocation + the single MMA-issuing thread.  Its descriptors are loop-invariant up to the start-address field (a K
//               step is one add to the low word); per iteration it issues  MMA1(t), Gram(t - 1), MMA2(t - 1);
//   warps 5-12  producers: (a) gather — thread g < TE copies the (48-byte padded) row x_j of edge g, thread p < TE / 20 the row x_i of
//               point p, into a shared ring RING - 1 tiles ahead with cp.async (neighbour indices travel through the ring as well);
//               (b) E(t); (c) H(t - 1) from D1, so that the first-layer MMA of tile t overlaps the second half of the step.
// Pipelines (mbarriers): E tiles and D1 accumulators (2 each), H stages (2), z accumulators (2), the transposed tile (1: the Gram MMAs
// are issued first), the Gram accumulator (1).  Tiles: 120 edges (rows 120..127 of the M = 128 operand are whatever follows in shared
// memory — their D1 rows are never read) without the Gram, 80 edges with it (shared-memory budget of the transposed tile).
//
// Measured (profiles/r03*): tcgen05.mma runs at N / 2 cycles per instruction for M = 64 and at the shared-memory fetch time
// (32 M + 32 N bytes at 128 B / cycle) for M = 128, dependent or not, A from shared memory or TMEM (tools/ubench/mma_rate.cu):
// 15.6 tensor cycles per edge without the Gram, 23.9 with it.  Cycle counters around every wait (-DSGB_ROLE_CLOCKS): the issuing
// thread is busy 84-88 % of the kernel, the producers 65-90 %, the epilogue ~45 %.
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

#ifdef SGB_ROLE_CLOCKS
#include <cstdio>
#define WAITC(i, x) do { const long long t0_ = clock64(); x; wc[i] += clock64() - t0_; } while (0)
#else
#define WAITC(i, x) x
#endif
namespace sgb_ectc {
using namespace sgb_tc;
using sgb_ec::CIN;
using sgb_ec::COUT;
using sgb_ec::KNN;
using sgb_bn::lrelu;

constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32; // 256
#ifndef SGB_EPI1
#define SGB_EPI1 4
#endif
constexpr int EPI1 = SGB_EPI1, NH1 = EPI1 / 4, MMA1W = EPI1;         // epilogue warps (NH1 per TMEM quadrant), the issuing warp follows, then the producers
constexpr int THREADS1 = (EPI1 + 1 + PROD_WARPS) * 32;         // 416
constexpr int W2_BYTES = COUT * COUT * 4;     // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int FLUSH = 32;                     // tiles per Gram segment
constexpr int GN = 80;                        // Gram accumulator columns: 64 hidden + ones + 15 zero rows
constexpr int GR = 144;                       // rows of the transposed tile: 64 lo + 64 hi + 16 extra

// byte offset of (row r, edge e) in the transposed K-major SWIZZLE_64B tile (atoms of 8 rows x 16 edges, 512 B)
__device__ __forceinline__ uint32_t ht_off(int r, int e) {
    return (uint32_t)((e >> 4) * (GR * 64) + (r >> 3) * 512 + (r & 7) * 64 + ((((e & 15) >> 2) ^ ((r >> 1) & 3)) << 4) + (e & 3) * 4);
}

template <bool GRAM> struct Cfg1 {
    static constexpr int TE = GRAM ? 80 : 120;
    static constexpr int PTS = TE / KNN;
    static constexpr int KE = 24;                                  // K of the first layer
    static constexpr int TILE_BYTES = TE * COUT * 4;               // one K-major H tile (hi or lo)
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
    static constexpr int HT_BYTES = GRAM ? GR * TE * 4 : 0;        // transposed tile (lo, hi, ones / zero rows), one buffer
    static constexpr int E_BYTES = TE * 20 * 4;                    // one K-major E tile (hi or lo): chunks 0..4; the zero chunk (K = 20..23) is shared
    static constexpr int EZ_BYTES = 128 * 16;                      // the shared zero chunk, M = 128 rows
    static constexpr int W1_BYTES = COUT * KE * 4;
    static constexpr int w2_hi = 0;
    static constexpr int w2_lo = w2_hi + W2_BYTES;
    static constexpr int w1_hi = w2_lo + W2_BYTES;
    static constexpr int w1_lo = w1_hi + W1_BYTES;
    static constexpr int stage0 = w1_lo + W1_BYTES;
    static constexpr int ht0 = stage0 + 2 * STAGE_BYTES;
    static constexpr int e0 = ht0 + HT_BYTES;                      // E tiles: [buffer 0: hi, lo][buffer 1: hi, lo][zero chunk]
    static constexpr int ezero = e0 + 4 * E_BYTES;
    static constexpr int RING = GRAM ? 4 : 3;
    static constexpr int RAW_BYTES = TE * 48 + PTS * 48 + TE * 8;
    static constexpr int raw0 = ezero + EZ_BYTES;                  // (the M = 128 read of chunk 4 runs (128 - TE) * 16 bytes into whatever follows the tile)
    static constexpr int bars = raw0 + RING * RAW_BYTES;           // 21 mbarriers
    static constexpr int tmem_slot = bars + 22 * 8;
    static constexpr int total = tmem_slot + 16;
    static constexpr int Z_COL = GRAM ? 80 : 128;                  // TMEM: z accumulators at 0 and Z_COL
    static constexpr int D1_COL = GRAM ? 160 : 256;                //       first-layer accumulators (2 x 64 columns)
    static constexpr int G_COL0 = 288;                             //       Gram accumulator (80 columns, one buffer)
    static constexpr int W_COL = GRAM ? 368 : 384;                 //       W2 hi | lo as the A operand of the second layer (2 x 64 columns)
};
static_assert(Cfg1<true>::ht0 % 1024 == 0 && Cfg1<true>::stage0 % 1024 == 0, "swizzled tile alignment");
static_assert(Cfg1<true>::bars % 8 == 0 && Cfg1<false>::bars % 8 == 0, "barrier alignment");
static_assert(Cfg1<true>::total + 1024 <= 227 * 1024 && Cfg1<false>::total + 1024 <= 227 * 1024, "shared memory budget");
static_assert(Cfg1<true>::e0 % 16 == 0 && Cfg1<false>::e0 % 16 == 0 && Cfg1<true>::E_BYTES % 16 == 0 && Cfg1<false>::E_BYTES % 16 == 0, "descriptor alignment");

template <bool ARG, bool GRAM>
__global__ void __launch_bounds__(THREADS1, 1)     // 13 warps: one scheduler hosts 4 of them -> 128 registers per thread at most
ec2_tc1_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
               const float* __restrict__ stats1, const float* __restrict__ W2, const float* __restrict__ gamma2,
               float* __restrict__ zsel, unsigned char* __restrict__ ksel, double* __restrict__ part /*[grid][128]*/,
               float* __restrict__ gslots /*[grid][nflush][128*GN]*/, int nflush) {
    using C = Cfg1<GRAM>;
    constexpr int TE = C::TE;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + C::bars);          // [2] producers -> MMA: H stage written
    uint64_t* bar_empty = bar_full + 2;                                      // [2] MMA (commit after the second-layer MMAs) -> producers
    uint64_t* bar_tfull = bar_full + 4;                                      // [2] MMA (commit) -> epilogue
    uint64_t* bar_tempty = bar_full + 6;                                     // [2] epilogue -> MMA
    uint64_t* bar_gfull = bar_full + 8;                                      // [2] MMA (commit) -> epilogue: Gram segment complete
    uint64_t* bar_gempty = bar_full + 10;                                    // [2] epilogue -> MMA: Gram accumulator flushed
    uint64_t* bar_efull = bar_full + 12;                                     // [2] producers -> MMA: E tile written
    uint64_t* bar_eempty = bar_full + 14;                                    // [2] MMA (commit after the first-layer MMAs) -> producers
    uint64_t* bar_d1full = bar_full + 16;                                    // [2] MMA (same commit) -> producers: D1 complete
    uint64_t* bar_d1empty = bar_full + 18;                                   // [2] producers -> MMA: D1 read into registers
    uint64_t* bar_htempty = bar_full + 20;                                   // MMA (commit after the Gram MMAs) -> producers: transposed tile free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SGB_ROLE_CLOCKS
    long long wc[32]; for (int i = 0; i < 32; ++i) wc[i] = 0; const long long tstart = clock64();
#endif

    const int per = N / gridDim.x, rem = N % gridDim.x;
    const int p_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int p_end = p_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
    const long long g_begin = (long long)p_begin * KNN, g_end = (long long)p_end * KNN;
    const int ntiles = (int)((g_end - g_begin + TE - 1) / TE);

    // ---- one-time setup
    for (int i = tid; i < COUT * COUT; i += THREADS1) {              // W2 [c][j] -> K-major canonical tile (rows = c, K = j)
        const int c = i / COUT, j = i % COUT;
        const float w = __ldg(W2 + i);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, j, COUT);
        *reinterpret_cast<float*>(sm + C::w2_hi + off) = hi;
        *reinterpret_cast<float*>(sm + C::w2_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < COUT * C::KE; i += THREADS1) {             // W1s [c][k]: BN1 scale folded in, column 18 = folded bias
        const int c = i / C::KE, k = i % C::KE;
        const float sc = stats1[128 + c];
        float w = 0.f;
        if (k < CIN) w = __ldg(W1 + c * CIN + k) * sc;
        else if (k == CIN) w = fmaf(-sc, stats1[c], stats1[192 + c]);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, k, COUT);
        *reinterpret_cast<float*>(sm + C::w1_hi + off) = hi;
        *reinterpret_cast<float*>(sm + C::w1_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < 128; i += THREADS1)                       // K = 20..23 of every E tile: one shared chunk of zeros
        *reinterpret_cast<float4*>(sm + C::ezero + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < COUT * COUT; i += THREADS1) {             // identity (aliases H stage 0 until the first tile): W2 reaches TMEM as W2 * I
        const int n = i / COUT, j = i % COUT;
        *reinterpret_cast<float*>(sm + C::stage0 + tile_off(n, j, COUT)) = (n == j) ? 1.f : 0.f;
    }
    if (GRAM) {                                                     // extra rows of the transposed tile: ones, then zeros
        unsigned char* ht = sm + C::ht0;
        for (int i = tid; i < 16 * TE; i += THREADS1) {
            const int r = 128 + i / TE, e = i % TE;
            *reinterpret_cast<float*>(ht + ht_off(r, e)) = (r == 128) ? 1.f : 0.f;
        }
    }
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], EPI1); mbar_init(&bar_tempty[1], EPI1);
        mbar_init(&bar_gfull[0], 1); mbar_init(&bar_gfull[1], 1);
        mbar_init(&bar_gempty[0], EPI1); mbar_init(&bar_gempty[1], EPI1);
        for (int b2 = 0; b2 < 2; ++b2) {
            mbar_init(&bar_efull[b2], PROD_WARPS); mbar_init(&bar_eempty[b2], 1);
            mbar_init(&bar_d1full[b2], 1); mbar_init(&bar_d1empty[b2], PROD_WARPS);
        }
        mbar_init(bar_htempty, 1);
        mbar_fence_init();
    }
    if (warp == MMA1W) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp > MMA1W) {
        // ================= producers
        const int pw = warp - (MMA1W + 1);              // 0..7
        const int ptid = pw * 32 + lane;                // gather / E role: edge row `ptid` of the tile (if < TE)
        const int quad = warp & 3;                      // the TMEM lanes this warp may read: 32 quad .. 32 quad + 31
        const int chh = pw >> 2;                        // H role: channels 32 chh .. 32 chh + 31 of edge row 32 quad + lane
        const int hrow = quad * 32 + lane;
        constexpr int RING = C::RING;
        const bool gatherer = ptid < TE;
        const bool pgatherer = ptid < C::PTS;
        auto edge_in_range = [&](int t) -> bool {
            const long long g = g_begin + (long long)t * TE + ptid;
            return gatherer && t < ntiles && g < g_end;
        };
        auto cp16 = [&](void* dst, const float* src, bool valid) {       // 16-byte async copy, zero fill when !valid
            const uint32_t n = valid ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
        };
        auto issue_index_copy = [&](int t) {            // index of my edge of tile t -> index slot t % RING
            if (edge_in_range(t)) {
                unsigned char* slot = sm + C::raw0 + (t % RING) * C::RAW_BYTES + TE * 48 + C::PTS * 48 + TE * 4 + ptid * 4;
                const int* src = knn + (g_begin + (long long)t * TE + ptid);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(slot)), "l"(src) : "memory");
            }
        };
        auto issue_rows = [&](int t, int j) {           // j < 0: edge out of range.  Every thread commits one (possibly empty) group per call
            unsigned char* raw = sm + C::raw0 + (t % RING) * C::RAW_BYTES;
            if (gatherer) {
                const bool v = j >= 0;
                const float* src = x12 + (size_t)(v ? j : 0) * 12;
                unsigned char* dst = raw + ptid * 48;
                cp16(dst, src, v); cp16(dst + 16, src + 4, v); cp16(dst + 32, src + 8, v);      // !v: zero fill -> pad lane 11 = 0 marks the edge invalid
            }
            if (pgatherer) {
                const long long pt = g_begin / KNN + (long long)t * C::PTS + ptid;
                const bool v = t < ntiles && pt < (long long)p_end;
                const float* src = x12 + (size_t)(v ? pt : 0) * 12;
                unsigned char* dst = raw + TE * 48 + ptid * 48;
                cp16(dst, src, v); cp16(dst + 16, src + 4, v); cp16(dst + 32, src + 8, v);
            }
            issue_index_copy(t + RING - 1);             // lands before rows(t + RING - 1) are issued (same group as rows(t))
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto staged_index = [&](int t) -> int {         // read back my own copy (complete: its group has been waited for)
            if (!edge_in_range(t)) return -1;
            return *reinterpret_cast<const volatile int*>(sm + C::raw0 + (t % RING) * C::RAW_BYTES + TE * 48 + C::PTS * 48 + TE * 4 + ptid * 4);
        };
        auto split4 = [](const float4 y, float4& hi, float4& lo) {
            hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
            lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
        };
        const uint32_t d1addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(C::D1_COL + 32 * chh);
#pragma unroll 1
        for (int tt = 0; tt < RING - 1; ++tt) issue_rows(tt, edge_in_range(tt) ? __ldg(knn + (g_begin + (long long)tt * TE + ptid)) : -1);
#pragma unroll 1
        for (int t = 0; t <= ntiles; ++t) {
            if (t < ntiles) {
                // ---- E(t): the edge vectors of tile t as the A operand of the first layer
                WAITC(30, asm volatile("cp.async.wait_group %0;" :: "n"(RING - 2) : "memory"));
                WAITC(31, asm volatile("bar.sync 1, %0;" :: "n"(PROD_THREADS) : "memory"));
                issue_rows(t + RING - 1, staged_index(t + RING - 1));                    // refill the slot tile t - 1 used
                const int eb = t & 1;
                WAITC(0, mbar_wait(&bar_eempty[eb], (((uint32_t)t >> 1) & 1u) ^ 1u));
                if (gatherer) {
                    unsigned char* e_hi = sm + C::e0 + eb * (2 * C::E_BYTES);
                    unsigned char* e_lo = e_hi + C::E_BYTES;
                    const unsigned char* raw = sm + C::raw0 + (t % RING) * C::RAW_BYTES;
                    const int pt = (ptid * 205) >> 12;      // ptid / KNN
                    const float4* rj = reinterpret_cast<const float4*>(raw + ptid * 48);
                    const float4* ri = reinterpret_cast<const float4*>(raw + TE * 48 + pt * 48);
                    const float4 b0 = rj[0], b1 = rj[1], b2 = rj[2], a0 = ri[0], a1 = ri[1], a2 = ri[2];
                    const bool v = b2.w != 0.f;             // 1 for a gathered row, 0 for a zero-filled one (edge outside the CTA's range)
                    const float one = v ? 1.f : 0.f;
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
                if (lane == 0) mbar_arrive(&bar_efull[eb]);
            }
            if (t >= 1) {
                // ---- H(u): hidden activations of tile u = t - 1 from the first-layer accumulator
                const int u = t - 1, su = u & 1;
                WAITC(1, mbar_wait(&bar_d1full[su], ((uint32_t)u >> 1) & 1u));
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
                WAITC(2, mbar_wait(&bar_empty[su], (((uint32_t)(u >> 1)) & 1u) ^ 1u));
                unsigned char* dst_hi = sm + C::stage0 + su * C::STAGE_BYTES;
                unsigned char* dst_lo = dst_hi + C::TILE_BYTES;
                const bool rowv = hrow < TE;
                if (GRAM) WAITC(3, mbar_wait(bar_htempty, ((uint32_t)u & 1u) ^ 1u));      // the Gram MMAs of tile u - 1 were issued before its second layer
                // transposed copy of the Gram variant: rows 0..63 lo, 64..127 hi (row = channel), 32 consecutive edges of one row per store
                unsigned char* dst_t = sm + C::ht0 + (uint32_t)chh * (4 * 512);
                const uint32_t e1 = (uint32_t)((hrow >> 4) * (GR * 64) + (hrow & 3) * 4), eq = (uint32_t)(((hrow & 15) >> 2) << 4);
                unsigned char* htb[4];                  // the four swizzle classes of a row: everything else is an immediate offset
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) htb[k4] = dst_t + e1 + (eq ^ (uint32_t)(k4 << 4));
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    float y[4];
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const int i = 4 * i4 + s4;
                        y[s4] = lrelu(__uint_as_float(i < 16 ? va[i] : vb[i - 16]));
                    }
                    float4 hi, lo;
                    split4(make_float4(y[0], y[1], y[2], y[3]), hi, lo);
                    if (rowv) {
                        const uint32_t off = (uint32_t)(8 * chh + i4) * (TE * 16) + (uint32_t)hrow * 16;
                        *reinterpret_cast<float4*>(dst_hi + off) = hi;
                        *reinterpret_cast<float4*>(dst_lo + off) = lo;
                        if (GRAM) {                         // (the warp's two halves write neighbouring 16-edge atoms, 9216 bytes apart: a 2-way bank
                            // conflict per store; swapping channel pairs in the upper half removes it but costs a select per store and measured slower)
                            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                            for (int s4 = 0; s4 < 4; ++s4) {
                                const int ci = 4 * i4 + s4;                  // channel 32 chh + ci = row of the transposed tile
                                unsigned char* q = htb[(ci >> 1) & 3] + ((ci >> 3) * 512 + (ci & 7) * 64);
                                *reinterpret_cast<float*>(q) = lv[s4];
                                *reinterpret_cast<float*>(q + 8 * 512) = hv[s4];
                            }
                        }
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[su]);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == MMA1W) {
        // ================= MMA issuer
        const uint32_t idesc1 = make_idesc_tf32(128, COUT, false, false);
        const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
        const uint32_t idesc_g = make_idesc_tf32(128, GN, false, false);
        const uint32_t a_hi = smem_u32(sm + C::w2_hi), a_lo = smem_u32(sm + C::w2_lo);
        const uint32_t w1h = smem_u32(sm + C::w1_hi), w1l = smem_u32(sm + C::w1_lo);
        const uint32_t ez = smem_u32(sm + C::ezero);
        // W2 (hi, lo) into TMEM in the layout of an M = 64 accumulator: D = W2 * I (exact: tf32 values times one).  The tensor pipe runs in
        // issue order, so every later MMA — and the first write into H stage 0, which waits for the first-layer MMA of tile 0 — comes after.
        if (elect_one_sync()) {
            const uint32_t idw = make_idesc_tf32(64, COUT, false, false);
            const uint32_t idn = smem_u32(sm + C::stage0);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int i = 0; i < COUT / 8; ++i) {
                    const uint32_t o = (uint32_t)(2 * i) * (COUT * 16);
                    mma_tf32(tmem + (uint32_t)(C::W_COL + h2 * 64), make_desc((h2 ? a_lo : a_hi) + o, COUT * 16, 128), make_desc(idn + o, COUT * 16, 128), idw, i > 0);
                }
        }
        __syncwarp();
        // Descriptors are loop-invariant up to the start-address field: bases are built once, a K step is an add of (bytes >> 4) to the low
        // word (smem addresses < 256 KB: the 14-bit field cannot carry).  The issuing thread is a single instruction stream that shares
        // its scheduler with three busy warps; at ~60 tensor cycles per instruction every ALU operation between two tcgen05.mma counts.
        auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
        const uint64_t dW1h = make_desc(w1h, COUT * 16, 128), dW1l = make_desc(w1l, COUT * 16, 128);
        uint64_t dEh[2], dEl[2], dEh2[2], dEl2[2], dHh[2], dHl[2];
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
            const uint32_t eh = smem_u32(sm + C::e0 + b2 * (2 * C::E_BYTES)), el = eh + C::E_BYTES;
            dEh[b2] = make_desc(eh, TE * 16, 128);
            dEl[b2] = make_desc(el, TE * 16, 128);
            dEh2[b2] = make_desc(eh + 4 * (TE * 16), ez - (eh + 4 * (TE * 16)), 128);      // K chunk 4 paired with the shared zero chunk
            dEl2[b2] = make_desc(el + 4 * (TE * 16), ez - (el + 4 * (TE * 16)), 128);
            const uint32_t hh = smem_u32(sm + C::stage0 + b2 * C::STAGE_BYTES);
            dHh[b2] = make_desc(hh, TE * 16, 128);
            dHl[b2] = make_desc(hh + C::TILE_BYTES, TE * 16, 128);
        }
        const uint64_t dGa = make_desc_sw(smem_u32(sm + C::ht0), 16, 512, 4), dGb = make_desc_sw(smem_u32(sm + C::ht0) + 8 * 512, 16, 512, 4);
        for (int t = 0; t <= ntiles; ++t) {
            if (t < ntiles) {
                const int eb = t & 1;
                const uint32_t ph1 = ((uint32_t)t >> 1) & 1u;
                WAITC(4, mbar_wait(&bar_efull[eb], ph1));
                WAITC(5, mbar_wait(&bar_d1empty[eb], ph1 ^ 1u));
                fence_after_sync();
                const uint64_t eh = eb ? dEh[1] : dEh[0], el = eb ? dEl[1] : dEl[0], eh2 = eb ? dEh2[1] : dEh2[0], el2 = eb ? dEl2[1] : dEl2[0];
                const uint32_t d1 = tmem + (uint32_t)(C::D1_COL + eb * 64);
                if (elect_one_sync()) {
#pragma unroll
                    for (int i = 0; i < C::KE / 8; ++i) {
                        const uint64_t dah = i < 2 ? adv(eh, 2 * i * (TE * 16)) : eh2, dal = i < 2 ? adv(el, 2 * i * (TE * 16)) : el2;
                        const uint64_t dbh = adv(dW1h, 2 * i * (COUT * 16)), dbl = adv(dW1l, 2 * i * (COUT * 16));
                        mma_tf32(d1, dah, dbh, idesc1, i > 0);
                        mma_tf32(d1, dal, dbh, idesc1, true);
                        mma_tf32(d1, dah, dbl, idesc1, true);
                    }
                    mma_commit(&bar_d1full[eb]);
                    mma_commit(&bar_eempty[eb]);
                }
                __syncwarp();
            }
            if (t >= 1) {
                const int u = t - 1, st = u & 1;
                const uint32_t ph = (uint32_t)(u >> 1) & 1u;
                const int seg = u / FLUSH;
                WAITC(6, mbar_wait(&bar_full[st], ph));
                WAITC(7, mbar_wait(&bar_tempty[st], ph ^ 1u));
                if (GRAM && u % FLUSH == 0) WAITC(8, mbar_wait(&bar_gempty[0], ((uint32_t)seg & 1u) ^ 1u));
                fence_after_sync();
                const uint64_t hh = st ? dHh[1] : dHh[0], hl = st ? dHl[1] : dHl[0];
                const uint32_t d = tmem + (uint32_t)(st * C::Z_COL), wb = tmem + (uint32_t)C::W_COL, dg = tmem + (uint32_t)C::G_COL0;
                const bool gacc = (u % FLUSH) > 0, glast = (u + 1) % FLUSH == 0 || u == ntiles - 1;
                if (elect_one_sync()) {
                    if (GRAM) {                         // first: the transposed tile has ONE buffer and is free again as soon as these complete
#pragma unroll
                        for (int s2 = 0; s2 < TE / 8; ++s2) {                              // 8 edges (K) per instruction
                            const uint32_t ko = (uint32_t)(s2 >> 1) * (GR * 64) + (uint32_t)(s2 & 1) * 32;
                            mma_tf32(dg, adv(dGa, ko), adv(dGb, ko), idesc_g, gacc || s2 > 0);
                        }
                        mma_commit(bar_htempty);
                        if (glast) mma_commit(&bar_gfull[0]);
                    }
#pragma unroll
                    for (int i = 0; i < COUT / 8; ++i) {                                   // A = W2 from TMEM: 8 columns (K) per instruction
                        const uint64_t dbh = adv(hh, 2 * i * (TE * 16)), dbl = adv(hl, 2 * i * (TE * 16));
                        mma_tf32_ts(d, wb + (uint32_t)(8 * i), dbh, idesc, i > 0);
                        mma_tf32_ts(d, wb + (uint32_t)(64 + 8 * i), dbh, idesc, true);
                        mma_tf32_ts(d, wb + (uint32_t)(8 * i), dbl, idesc, true);
                    }
                    mma_commit(&bar_tfull[st]);
                    mma_commit(&bar_empty[st]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: the two warps of TMEM quadrant q own channels 16q + lane (lanes 0..15) and split the points of a
        // tile between them; one point = 20 consecutive columns
        const int quad = warp & 3, half = warp >> 2;
        const int c = quad * 16 + (lane & 15);
        const int hp = lane >> 4;                       // an M = 64 accumulator keeps its rows on lanes 0..15 of the quadrant: the .16x32bx2 loads give
                                                        // the upper half-warp the NEXT point (20 columns further) of the same 16 channels
        const bool up = __ldg(gamma2 + c) > 0.f;
        float S1h = 0.f, S1l = 0.f, S2h = 0.f, S2l = 0.f;      // running sums as unevaluated fp32 pairs (two-sum)
        auto two_sum = [](float& h, float& l, float x) {
            const float s_ = h + x;
            const float bb = s_ - h;
            l += (h - (s_ - bb)) + (x - bb);
            h = s_;
        };
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            WAITC(9, mbar_wait(&bar_tfull[st], ph));
            fence_after_sync();
            const long long g0 = g_begin + (long long)t * TE;
            const int npts = (int)min((long long)C::PTS, (g_end - g0) / KNN);
            const long long p0 = g0 / KNN;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(st * C::Z_COL);
            const int pbeg = min(npts, half * (C::PTS / NH1)), pend = half == NH1 - 1 ? npts : min(npts, (half + 1) * (C::PTS / NH1));
            float ps[4] = {0.f, 0.f, 0.f, 0.f}, qs[4] = {0.f, 0.f, 0.f, 0.f};
            // BN2 + LeakyReLU is monotone per channel, increasing iff gamma2 > 0: only that extreme of the 20 pre-activations (and its
            // neighbour slot) is kept, found by a tournament (the left operand wins ties: the FIRST extreme, as in a left-to-right scan)
            auto reduce_point = [&](const uint32_t (&ra)[16], const uint32_t (&rb)[4], int pp, auto better) {
                float z[KNN];
#pragma unroll
                for (int i = 0; i < KNN; ++i) z[i] = __uint_as_float(i < 16 ? ra[i] : rb[i - 16]);
#pragma unroll
                for (int i = 0; i < KNN; ++i) { ps[i & 3] += z[i]; qs[i & 3] = fmaf(z[i], z[i], qs[i & 3]); }
                float m[10];
                int k[10];
#pragma unroll
                for (int j = 0; j < 10; ++j) { const bool r = better(z[2 * j + 1], z[2 * j]); m[j] = r ? z[2 * j + 1] : z[2 * j]; k[j] = r ? 2 * j + 1 : 2 * j; }
#pragma unroll
                for (int j = 0; j < 5; ++j) { const bool r = better(m[2 * j + 1], m[2 * j]); m[j] = r ? m[2 * j + 1] : m[2 * j]; k[j] = r ? k[2 * j + 1] : k[2 * j]; }
                { const bool r = better(m[1], m[0]); m[0] = r ? m[1] : m[0]; k[0] = r ? k[1] : k[0]; }
                { const bool r = better(m[3], m[2]); m[2] = r ? m[3] : m[2]; k[2] = r ? k[3] : k[2]; }
                { const bool r = better(m[2], m[0]); m[0] = r ? m[2] : m[0]; k[0] = r ? k[2] : k[0]; }
                { const bool r = better(m[4], m[0]); m[0] = r ? m[4] : m[0]; k[0] = r ? k[4] : k[0]; }
                if (pp + hp < pend) {
                    const size_t o = (size_t)(p0 + pp + hp) * COUT + c;
                    zsel[o] = m[0];
                    if (ARG) ksel[o] = (unsigned char)k[0];
                }
            };
            auto scan = [&](auto better) {              // a pair of points per load, two pairs per wait; columns past the last point hold z = 0 (h = 0)
#pragma unroll 1
                for (int pp = pbeg; pp < pend; pp += 4) {
                    uint32_t a16[16], a4[4], b16[16], b4[4];
                    const bool two = pp + 2 < pend;         // warp-uniform
                    tmem_ld16_halves_issue<KNN>(taddr + (uint32_t)(pp * KNN), a16);
                    tmem_ld4_halves_issue<KNN>(taddr + (uint32_t)(pp * KNN + 16), a4);
                    if (two) {
                        tmem_ld16_halves_issue<KNN>(taddr + (uint32_t)((pp + 2) * KNN), b16);
                        tmem_ld4_halves_issue<KNN>(taddr + (uint32_t)((pp + 2) * KNN + 16), b4);
                    }
                    tmem_ld_wait();
                    tmem_ld_pin20(a16, a4);
                    reduce_point(a16, a4, pp, better);
                    if (two) {
                        tmem_ld_pin20(b16, b4);
                        reduce_point(b16, b4, pp + 2, better);
                    }
                }
            };
            if (__all_sync(SGB_FULL_MASK, up)) scan([](float z, float b) { return z > b; });
            else if (__all_sync(SGB_FULL_MASK, !up)) scan([](float z, float b) { return z < b; });
            else scan([up](float z, float b) { return up ? z > b : z < b; });
            two_sum(S1h, S1l, (ps[0] + ps[1]) + (ps[2] + ps[3]));
            two_sum(S2h, S2l, (qs[0] + qs[1]) + (qs[2] + qs[3]));
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[st]);
            if (GRAM && ((t + 1) % FLUSH == 0 || t == ntiles - 1)) {
                // flush the finished Gram segment: thread = accumulator row (all 128 lanes hold data for M = 128)
                const int seg = t / FLUSH;
                WAITC(10, mbar_wait(&bar_gfull[0], (uint32_t)seg & 1u));
                fence_after_sync();
                float* dst = gslots + ((size_t)blockIdx.x * nflush + seg) * (128 * GN) + (size_t)(quad * 32 + lane) * GN;
                const uint32_t gaddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)C::G_COL0;
#pragma unroll 1
                for (int c0 = 16 * half; c0 < GN; c0 += 16 * NH1) {       // the quadrant's two warps take alternate 16-column groups
                    float v[16];
                    tmem_ld16(gaddr + (uint32_t)c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(dst + c0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_gempty[0]);
            }
        }
        double s1 = (double)S1h + (double)S1l, s2 = (double)S2h + (double)S2l;
        s1 += __shfl_xor_sync(SGB_FULL_MASK, s1, 16);
        s2 += __shfl_xor_sync(SGB_FULL_MASK, s2, 16);
        if (hp == 0) {
            part[((size_t)blockIdx.x * NH1 + half) * 128 + c] = s1;
            part[((size_t)blockIdx.x * NH1 + half) * 128 + 64 + c] = s2;
        }
    }
#ifdef SGB_ROLE_CLOCKS
    if (blockIdx.x == 3 && lane == 0) { const long long tot = clock64() - tstart; for (int i = 0; i < 32; ++i) if (wc[i]) printf("warp %2d wait %2d : %6.1f%%  (total %lld clk, %d tiles)\n", warp, i, 100.0 * wc[i] / tot, tot, ntiles); }
#endif
    fence_before_sync();
    __syncthreads();
    if (warp == MMA1W) tmem_dealloc(tmem, TMEM_COLS);
}

// mom2 [64*64 + 64] (sum h h^T, sum h) from the reduced Gram accumulator R [128][GN]: rows 0..63 = lo*(hi | 1), rows 64..127 = hi*(hi | 1),
// row of the transposed tile = hidden channel
__global__ void __launch_bounds__(64)
mom2_from_gram1_kernel(const double* __restrict__ R, double* __restrict__ mom2) {
    const int j = threadIdx.x;
    for (int i = 0; i < COUT; ++i) mom2[j * COUT + i] = R[(64 + j) * GN + i] + R[j * GN + i] + R[i * GN + j];     // hi_j.hi_i + lo_j.hi_i + hi_j.lo_i
    mom2[COUT * COUT + j] = R[(64 + j) * GN + 64] + R[j * GN + 64];
}

// BN2 statistics from the reduced sums (sum z [64], sum z^2 [64]); same stats layout as the SIMT path
__global__ void __launch_bounds__(64)
bn2_from_sums_kernel(const double* __restrict__ sums, double M, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ stats, float* __restrict__ var_out) {
    const int c = threadIdx.x;
    const double mean = sums[c] / M;
    double var = sums[64 + c] / M - mean * mean;
    if (var < 0) var = 0;
    const double invstd = 1.0 / sqrt(var + (double)sgb_bn::BN_EPS);
    stats[c] = (float)mean;
    stats[64 + c] = (float)invstd;
    stats[128 + c] = (float)((double)gamma[c] * invstd);
    stats[192 + c] = beta[c];
    if (var_out) var_out[c] = (float)var;
}

// out[p, c] = lrelu(BN2(z*)), z* = the extreme the forward kernel kept for the sign of the BN scale; argk = its neighbour slot
__global__ void __launch_bounds__(256)
ec2_apply_kernel(const float* __restrict__ zsel, const unsigned char* __restrict__ ksel,
                 const float* __restrict__ stats2, long long total, float* __restrict__ out, unsigned char* __restrict__ argk) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= total) return;
    const int c = (int)(i4 & 63);
    const float4 a = *reinterpret_cast<const float4*>(zsel + i4);
    const uchar4 kq = argk ? *reinterpret_cast<const uchar4*>(ksel + i4) : make_uchar4(0, 0, 0, 0);
    const float za[4] = {a.x, a.y, a.z, a.w};
    const unsigned char kv[4] = {kq.x, kq.y, kq.z, kq.w};
    float o[4];
    unsigned char ak[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float mean = stats2[c + q], scale = stats2[128 + c + q], beta = stats2[192 + c + q];
        o[q] = lrelu(fmaf(za[q] - mean, scale, beta));
        ak[q] = scale == 0.f ? (unsigned char)0 : kv[q];
    }
    *reinterpret_cast<float4*>(out + i4) = make_float4(o[0], o[1], o[2], o[3]);
    if (argk) *reinterpret_cast<uchar4*>(argk + i4) = make_uchar4(ak[0], ak[1], ak[2], ak[3]);
}

inline int tc_grid(int N, int TE) {
    const int tiles = sgb_div_up((long long)N * KNN, TE);
    return tiles < 148 ? (tiles < 1 ? 1 : tiles) : 148;
}
inline int tc_nflush(int N, int grid) {                 // Gram segments per CTA (the largest CTA range)
    const long long pts = (N + grid - 1) / grid;
    const int tiles = sgb_div_up(pts * KNN, Cfg1<true>::TE);
    return sgb_div_up(tiles > 0 ? tiles : 1, FLUSH);
}
}  // namespace sgb_ectc

// workspace: partial sums [2 * 148][128] f64, reduced sums [128] f64, reduced Gram [128*GN] f64, zsel [N,64] f32, ksel [N,64] u8,
// Gram slots [148][nflush][128*GN] f32
size_t sgb_ec2_tc_ws_bytes(int N) {
    using namespace sgb_ectc;
    const int nf = tc_nflush(N, 148) + 1;
    return (size_t)(2 * 148 + 1) * 128 * 8 + (size_t)128 * GN * 8 + (size_t)N * 64 * (4 + 1) +
           (size_t)148 * nf * 128 * GN * 4 + 1024;
}

// both layers of MLP3 on the tensor cores: stats2/var2 and out/argk as sgb_edgeconv_fwd produces them; mom2 (optional)
// = second moments of the hidden activations for the backward pass
int sgb_ec2_tc_forward(const float* x12 /*[N,12]: 48-byte padded rows*/, const int* knn, int N, const float* W1, const float* stats1, const float* W2,
                       const float* gamma2, const float* beta2, float* out, unsigned char* argk, float* stats2, float* var2,
                       double* mom2, void* ws, cudaStream_t st) {
    using namespace sgb_ectc;
    unsigned char* w8 = (unsigned char*)ws;
    double* part = (double*)w8;
    double* sums = part + 2 * 148 * 128;
    double* gred = sums + 128;
    float* zsel = (float*)(gred + 128 * GN);
    unsigned char* ksel = (unsigned char*)(zsel + (size_t)N * 64);
    float* gslots = (float*)(((uintptr_t)(ksel + (size_t)N * 64) + 255) & ~(uintptr_t)255);
    const bool gram = mom2 != nullptr;
    const int grid = tc_grid(N, gram ? Cfg1<true>::TE : Cfg1<false>::TE);
    const int nflush = gram ? tc_nflush(N, grid) : 0;
    if (gram) {
        const size_t smem = Cfg1<true>::total + 1024;
        SGB_CUDA(cudaMemsetAsync(gslots, 0, (size_t)grid * nflush * 128 * GN * 4, st));
        SGB_OPT_IN_SMEM((ec2_tc1_kernel<true, true>));
        { ec2_tc1_kernel<true, true><<<grid, THREADS1, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, gslots, nflush); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(gslots, grid * nflush, 128 * GN, gred, st);
        { mom2_from_gram1_kernel<<<1, 64, 0, st>>>(gred, mom2); SGB_COUNT_LAUNCH(); }
    } else if (argk) {
        const size_t smem = Cfg1<false>::total + 1024;
        SGB_OPT_IN_SMEM((ec2_tc1_kernel<true, false>));
        { ec2_tc1_kernel<true, false><<<grid, THREADS1, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    } else {
        const size_t smem = Cfg1<false>::total + 1024;
        SGB_OPT_IN_SMEM((ec2_tc1_kernel<false, false>));
        { ec2_tc1_kernel<false, false><<<grid, THREADS1, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    }
    sgb_bn::reduce_partials(part, NH1 * grid, 128, sums, st);
    { bn2_from_sums_kernel<<<1, 64, 0, st>>>(sums, (double)N * KNN, gamma2, beta2, stats2, var2); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)N * 64;
    { ec2_apply_kernel<<<sgb_div_up(total / 4, 256), 256, 0, st>>>(zsel, ksel, stats2, total, out, argk); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

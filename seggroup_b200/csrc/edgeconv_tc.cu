// a9 on the tensor cores: second EdgeConv layer of MLP3 (seggroup/model.py:121-138) as a warp-specialised tcgen05 kernel.
//
//   z[c, e] = sum_j W2[c, j] h[e, j],   h[e, :] = lrelu(BN1(W1 e_e)),   e = (x_j - x_i, x_i),  j in knn(i)
//
// The 64x64 contraction per edge (8.2 kFLOP x 3 M edges per scene) is the one GEMM-shaped piece of SegModel.  It runs as
// D[64 channels, TE edges] += W2[64, 64] * H[TE, 64]^T with kind::tf32 and the TF32 x 3 split of tc_common.cuh, so the
// features keep fp32-level accuracy (the merge decisions downstream are discrete).  Orientation: channels on the TMEM
// lanes, edges on the TMEM columns, so that one epilogue thread owns one channel and both reductions over edges are
// thread-local:
//   * BatchNorm-2 batch statistics  sum z, sum z^2  (fp32 per tile, fp64 across tiles, fixed order);
//   * max / min over the 20 neighbours of a point.  BN (per-channel affine) followed by LeakyReLU is monotone, so
//         max_k lrelu(BN2(z_k)) = lrelu(BN2(max_k z_k))  if gamma2 >= 0,   lrelu(BN2(min_k z_k))  otherwise,
//     which removes the separate statistics pass of the SIMT path: ONE pass over the edges produces the statistics and
//     the (max, min, arg) candidates; a light second kernel applies BN2 + LeakyReLU per point.
// GRAM variant (training): the backward pass needs the second moments of the hidden activations, sum_e h h^T and sum_e h
// (analytic BatchNorm backward, edgeconv_bwd.cu).  They are an edge contraction, i.e. the edges must lie along K.
// kind::tf32 only walks MN-major operands in one swizzle mode that no K-major layout shares (measured with a one-instruction descriptor probe during bring-up, round 1), so the
// producers store h a second time, transposed ([hidden row][edges], K-major SWIZZLE_64B), and a second accumulator
//   G[128, 80] += [H_lo^T ; H_hi^T] (K = edges) x [H_hi^T ; 1 ; 0]^T
// collects lo*hi (rows 0..63), hi*hi (rows 64..127) and the column of ones gives sum h; hi*lo follows by symmetry.
// G is flushed to global fp32 slots every FLUSH tiles and summed in fp64 in a fixed order (TMEM accumulates in fp32).
//
// Warp roles (416 threads, 1 CTA / SM, persistent over a contiguous range of points):
//   warps 0-3   epilogue: tcgen05.ld of their TMEM sub-partition (an M = 64 accumulator keeps rows 16q..16q+15 on lanes
//               32q..32q+15), statistics + max/min over 20 consecutive columns, stores of the per-point candidates;
//   warp  4     TMEM allocation + the single MMA-issuing thread + tcgen05.commit;
//   warps 5-12  producers.  (a) gather: thread g < TE copies the (48-byte padded) row x_j of edge g, thread p < TE / 20 the row x_i of
//               point p, into a shared ring RING - 1 tiles ahead with cp.async (neighbour indices by LDG two tiles before that):
//               the random row gathers cost ~1 us of latency against ~0.5 us per tile, so one tile of register look-ahead (round
//               1) left the producers waiting on the scoreboard; (b) first layer: a thread owns ONE
//               16-byte chunk of the hidden vector (4 channels) for the whole kernel, so its 72 first-layer weights (BN1
//               folded in) live in REGISTERS — round 1 re-read them from shared memory for every edge, 72 LDS.128 per thread
//               and tile, which was the largest consumer of a shared-memory-bound kernel (l1tex 93 %, tensor pipe 22 %) — and
//               walks the tile in blocks of 8 edges (lane & 7 = edge, lane >> 3 = which of the warp's 4 chunks): 5 broadcast
//               LDS.128 of the two rows (9 subtractions rebuild the edge vector), 36 packed FFMA2, LeakyReLU, hi/lo split, 16-byte stores into the canonical
//               no-swizzle K-major tile (+ the 4-byte transposed stores of the Gram variant), fence.proxy.async, mbarrier arrive.
// Pipelines: shared-memory tiles full/empty (2 stages), the gather ring (producer-internal: cp.async groups + one named barrier
// per tile) and TMEM accumulators full/empty (2 buffers).
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

#ifndef SGB_ABL
#define SGB_ABL 0      // role-ablation timing experiments (tools/ablate.sh): results are WRONG for any value but 0
#endif

namespace sgb_ectc {
using namespace sgb_tc;
using sgb_ec::CIN;
using sgb_ec::COUT;
using sgb_ec::KNN;
using sgb_bn::lrelu;

constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;           // warp 4
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;     // 416
constexpr int PROD_THREADS = PROD_WARPS * 32; // 256
// raw gather ring: per tile the 48-byte rows x_j of its TE edges, the rows x_i of its TE / 20 points and a validity word per edge,
// filled by cp.async several tiles ahead (no registers held across the latency of the random row gathers)
constexpr int W2_BYTES = COUT * COUT * 4;     // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int FLUSH = 32;                     // tiles per Gram segment
constexpr int GN = 80;                        // Gram accumulator columns: 64 hidden + ones + 15 zero rows
constexpr int GR = 144;                       // rows of the transposed tile: 64 lo + 64 hi + 16 extra

template <bool GRAM> struct Cfg {
    static constexpr int TE = GRAM ? 80 : 160;                     // edges per tile (TMEM columns per z accumulator)
    static constexpr int PTS = TE / KNN;                           // points per tile
    static constexpr int BLOCKS = TE / 8;                          // 8-edge blocks per tile; a group of 4 producer warps takes every 2nd one
    static constexpr int TILE_BYTES = TE * COUT * 4;               // one K-major H tile (hi or lo)
    static constexpr int HT_BYTES = GRAM ? GR * TE * 4 : 0;        // transposed tile (lo, hi, extra rows)
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES + HT_BYTES;  // multiple of 1024 in both variants
    // offsets into the dynamic shared memory block (1024-byte aligned base)
    static constexpr int w2_hi = 0;
    static constexpr int w2_lo = w2_hi + W2_BYTES;
    static constexpr int stage0 = w2_lo + W2_BYTES;
    static constexpr int RING = GRAM ? 4 : 3;                      // tiles in flight in the gather ring
    static constexpr int RAW_BYTES = TE * 48 + PTS * 48 + TE * 8;  // x_j rows, x_i rows, validity words, neighbour indices of a LATER tile
    static constexpr int raw0 = stage0 + 2 * STAGE_BYTES;
    static constexpr int bars = raw0 + RING * RAW_BYTES;           // 12 mbarriers
    static constexpr int tmem_slot = bars + 12 * 8;
    static constexpr int wc_tab = tmem_slot + 16;                  // centre-half first-layer weights [16 chunks][9] float4 (BN1 folded in)
    static constexpr int ctr_buf = wc_tab + 16 * 9 * 16;           // per producer warp: centre rows [PTS][4 parts] float4 of the current tile
    static constexpr int total = ctr_buf + PROD_WARPS * PTS * 4 * 16;
    static constexpr int Z_COL = 128 + (GRAM ? 0 : 128);           // TMEM column stride of the two z accumulators
    static constexpr int G_COL0 = 256;                             // TMEM columns of the two Gram accumulators (256, 384)
};
static_assert(Cfg<true>::STAGE_BYTES % 1024 == 0 && Cfg<false>::STAGE_BYTES % 1024 == 0, "stage alignment");
static_assert((2 * Cfg<true>::TILE_BYTES) % 1024 == 0, "transposed tile alignment");
static_assert(Cfg<true>::total + 1024 <= 227 * 1024 && Cfg<false>::total + 1024 <= 227 * 1024, "shared memory budget");

// row of the transposed tile that holds hidden channel (part p, local index i): chosen so that the four parts of a warp hit
// distinct (row parity, swizzle class) pairs -> conflict-free 4-byte stores
__host__ __device__ constexpr int ht_row(int p, int i) { return 8 * (i >> 1) + ((p & 1) + 4 * (p >> 1)) + 2 * (i & 1); }

// byte offset of (row r, edge e) in the transposed K-major SWIZZLE_64B tile (atoms of 8 rows x 16 edges, 512 B)
__device__ __forceinline__ uint32_t ht_off(int r, int e) {
    return (uint32_t)((e >> 4) * (GR * 64) + (r >> 3) * 512 + (r & 7) * 64 + ((((e & 15) >> 2) ^ ((r >> 1) & 3)) << 4) + (e & 3) * 4);
}

template <bool ARG, bool GRAM>
__global__ void __launch_bounds__(THREADS, 1)      // 13 warps: one scheduler hosts 4 of them -> 16384 / (4 * 32) = 128 registers per thread at most
ec2_tc_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
              const float* __restrict__ stats1, const float* __restrict__ W2, const float* __restrict__ gamma2,
              float* __restrict__ zsel, unsigned char* __restrict__ ksel, double* __restrict__ part /*[grid][128]*/,
              float* __restrict__ gslots /*[grid][nflush][128*GN]*/, int nflush) {
    using C = Cfg<GRAM>;
    constexpr int TE = C::TE;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // keeps the shared address space
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + C::bars);          // [2] producers -> MMA
    uint64_t* bar_empty = bar_full + 2;                                      // [2] MMA (commit) -> producers
    uint64_t* bar_tfull = bar_full + 4;                                      // [2] MMA (commit) -> epilogue
    uint64_t* bar_tempty = bar_full + 6;                                     // [2] epilogue -> MMA
    uint64_t* bar_gfull = bar_full + 8;                                      // [2] MMA (commit) -> epilogue: Gram segment complete
    uint64_t* bar_gempty = bar_full + 10;                                    // [2] epilogue -> MMA: Gram accumulator flushed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // contiguous, balanced range of points for this CTA
    const int per = N / gridDim.x, rem = N % gridDim.x;
    const int p_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int p_end = p_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
    const long long g_begin = (long long)p_begin * KNN, g_end = (long long)p_end * KNN;
    const int ntiles = (int)((g_end - g_begin + TE - 1) / TE);

    // ---- one-time setup
    for (int i = tid; i < COUT * COUT; i += THREADS) {              // W2 [c][j] -> K-major canonical tile (rows = c, K = j)
        const int c = i / COUT, j = i % COUT;
        const float w = __ldg(W2 + i);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, j, COUT);
        *reinterpret_cast<float*>(sm + C::w2_hi + off) = hi;
        *reinterpret_cast<float*>(sm + C::w2_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < 16 * 9; i += THREADS) {                   // centre half of the first layer: columns 9..17 of W1, BN1 scale folded in
        const int ch = i / 9, q = i % 9;
        float w[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) w[s4] = __ldg(W1 + (4 * ch + s4) * CIN + 9 + q) * stats1[128 + 4 * ch + s4];
        *reinterpret_cast<float4*>(sm + C::wc_tab + i * 16) = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (GRAM) {                                                     // extra rows of the transposed tiles: ones, then zeros
        for (int s = 0; s < 2; ++s) {
            unsigned char* ht = sm + C::stage0 + s * C::STAGE_BYTES + 2 * C::TILE_BYTES;
            for (int i = tid; i < 16 * TE; i += THREADS) {
                const int r = 128 + i / TE, e = i % TE;
                *reinterpret_cast<float*>(ht + ht_off(r, e)) = (r == 128) ? 1.f : 0.f;
            }
        }
    }
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], EPI_WARPS); mbar_init(&bar_tempty[1], EPI_WARPS);
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
        // ================= producers
        const int pw = warp - (MMA_WARP + 1);           // 0..7
        const int ptid = pw * 32 + lane;                // 0..255: gather role = edge row `ptid` of the tile (if < TE)
        const int grp = pw >> 2;                        // which half of the 8-edge blocks this warp computes
        const int part = lane >> 3;                     // hidden channels 16 part .. 16 part + 15 (the Gram row mapping of ht_row)
        const int c4 = part * 4 + (pw & 3);             // my 16-byte chunk of the hidden vector: channels 4 c4 .. 4 c4 + 3, for the whole kernel
        const int eb = lane & 7;                        // my edge inside an 8-edge block
        // first layer with the BatchNorm-1 affine folded in:  v1 = (scale1 W1) e + (beta1 - scale1 mean1).  The edge vector is
        // (x_j - x_i, x_i): the centre half  bias + Wc x_i  is the same for the 20 edges of a point and is evaluated ONCE per point and
        // tile (each warp for its own 4 chunks, below); only the 9 difference columns are per-edge work and only their weights
        // (36 registers instead of 72) stay in registers for the whole kernel.
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
        const float4* wc = reinterpret_cast<const float4*>(sm + C::wc_tab) + c4 * 9;
        unsigned char* ctr = sm + C::ctr_buf + pw * (C::PTS * 4 * 16);
        uint32_t ht_rb[4], ht_rx[4];                    // my 4 rows of the transposed tile: byte offset of the row, swizzle term
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int r = ht_row(part, (pw & 3) * 4 + s4);
            ht_rb[s4] = (uint32_t)((r >> 3) * 512 + (r & 7) * 64);
            ht_rx[s4] = (uint32_t)(((r >> 1) & 3) << 4);
        }
        // gather pipeline: neighbour indices by LDG two iterations before they are needed, rows by cp.async RING - 1 tiles ahead
        constexpr int RING = C::RING;
        const bool gatherer = ptid < TE;                // fetches x_j of edge row `ptid`
        const bool pgatherer = ptid < C::PTS;           // fetches x_i of point `ptid` of the tile
        // Neighbour indices travel through shared memory as well (4-byte cp.async, RING - 1 tiles before the rows that need them):
        // held in rotating registers, the first register move after the LDG waits for it, i.e. one tile of look-ahead at best.
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
#pragma unroll 1
        for (int tt = 0; tt < RING - 1; ++tt) issue_rows(tt, edge_in_range(tt) ? __ldg(knn + (g_begin + (long long)tt * TE + ptid)) : -1);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            asm volatile("cp.async.wait_group %0;" :: "n"(RING - 2) : "memory");     // my copies for tile t (and the indices of tile t + RING - 1) have landed
            asm volatile("bar.sync 1, %0;" :: "n"(PROD_THREADS) : "memory");         // everybody's have, and everybody is done with tile t - 1
            issue_rows(t + RING - 1, staged_index(t + RING - 1));                    // refill the slot tile t - 1 used
            const unsigned char* raw = sm + C::raw0 + (t % RING) * C::RAW_BYTES;
            if (eb < C::PTS) {                           // lane (part, eb): centre row of point eb of the tile for my chunk c4
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
            mbar_wait(&bar_empty[st], ph ^ 1u);
            unsigned char* dst_hi = sm + C::stage0 + st * C::STAGE_BYTES;
            unsigned char* dst_lo = dst_hi + C::TILE_BYTES;
            unsigned char* dst_t = dst_lo + C::TILE_BYTES;
            // two 8-edge blocks in flight per thread: the shared-memory loads of the next block are issued before the arithmetic of the
            // current one (ncu source view of the one-block loop: 22 % of the producers' samples waited on LDS, 24 % on dependent FMAs)
            struct Blk { float4 b0, b1, b2, a0, a1, a2, cc; };
            const uint32_t raw_s = smem_u32(raw), ctr_s = smem_u32(ctr);
            auto load_blk = [&](int blk, Blk& B) {
                const int er = blk * 8 + eb;            // edge row in the tile
                const int pt = (er * 205) >> 12;        // er / KNN, exact for er < 1039
                const uint32_t rj = raw_s + (uint32_t)(er * 48);                                          // 48-byte stride: conflict-free
                const uint32_t ri = raw_s + (uint32_t)(TE * 48 + pt * 48);                                // the edge's own point (broadcast)
                B.b0 = lds128_ordered(rj); B.b1 = lds128_ordered(rj + 16); B.b2 = lds128_ordered(rj + 32);
                B.a0 = lds128_ordered(ri); B.a1 = lds128_ordered(ri + 16); B.a2 = lds128_ordered(ri + 32);
                B.cc = lds128_ordered(ctr_s + (uint32_t)((pt * 4 + part) * 16));
            };
            auto finish_blk = [&](int blk, const Blk& B) {
                const int er = blk * 8 + eb;
                const float ev[CD] = {B.b0.x - B.a0.x, B.b0.y - B.a0.y, B.b0.z - B.a0.z, B.b0.w - B.a0.w, B.b1.x - B.a1.x, B.b1.y - B.a1.y,
                                      B.b1.z - B.a1.z, B.b1.w - B.a1.w, B.b2.x - B.a2.x};
                float2 y01 = make_float2(B.cc.x, B.cc.y), y23 = make_float2(B.cc.z, B.cc.w);
#pragma unroll
                for (int q = 0; q < ((SGB_ABL & 1) ? 1 : CD); ++q) {
                    const float2 ee = make_float2(ev[q], ev[q]);
                    ffma2(y01, w01[q], ee);
                    ffma2(y23, w23[q], ee);
                }
                const float vm = B.b2.w;                 // 1 for a gathered row, 0 for a zero-filled one (edge outside the CTA's range)
                const float4 y = make_float4(lrelu(y01.x) * vm, lrelu(y01.y) * vm, lrelu(y23.x) * vm, lrelu(y23.y) * vm);
                const float4 hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                const float4 lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
                const uint32_t off = (uint32_t)c4 * (TE * 16) + (uint32_t)(er >> 3) * 128 + (uint32_t)(er & 7) * 16;
                *reinterpret_cast<float4*>(dst_hi + off) = hi;
                *reinterpret_cast<float4*>(dst_lo + off) = lo;
                if (GRAM && !(SGB_ABL & 16)) {          // transposed copy: rows 0..63 lo, 64..127 hi
                    const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
                    const uint32_t e1 = (uint32_t)((er >> 4) * (GR * 64) + (er & 3) * 4), eq = (uint32_t)(((er & 15) >> 2) << 4);
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {      // ht_off(r, er) with the row-dependent terms hoisted (ht_rb / ht_rx); row 64 + r = + 8 atoms
                        const uint32_t o = ht_rb[s4] + e1 + (eq ^ ht_rx[s4]);
                        *reinterpret_cast<float*>(dst_t + o) = lv[s4];
                        *reinterpret_cast<float*>(dst_t + o + 8 * 512) = hv[s4];
                    }
                }
            };
            Blk A, Bq;
            int blk = grp;
            load_blk(blk, A);
#pragma unroll 1
            for (; blk + 2 < C::BLOCKS; blk += 4) {
                load_blk(blk + 2, Bq);
                finish_blk(blk, A);
                if (blk + 4 < C::BLOCKS) load_blk(blk + 4, A);
                finish_blk(blk + 2, Bq);
            }
            if (blk < C::BLOCKS) finish_blk(blk, A);    // odd number of blocks per warp (TE = 80)
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[st]);
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer
        const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
        const uint32_t idesc_g = make_idesc_tf32(128, GN, false, false);
        const uint32_t a_hi = smem_u32(sm + C::w2_hi), a_lo = smem_u32(sm + C::w2_lo);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            const int seg = t / FLUSH, gb = seg & 1;
            mbar_wait(&bar_full[st], ph);
            mbar_wait(&bar_tempty[st], ph ^ 1u);
            if (GRAM && t % FLUSH == 0) mbar_wait(&bar_gempty[gb], ((uint32_t)(seg >> 1) & 1u) ^ 1u);
            fence_after_sync();
            if (elect_one_sync()) {
                const uint32_t b_hi = smem_u32(sm + C::stage0 + st * C::STAGE_BYTES), b_lo = b_hi + C::TILE_BYTES;
                const uint32_t d = tmem + (uint32_t)(st * C::Z_COL);
#pragma unroll
                for (int i = 0; i < COUT / 8; ++i) {
                    const uint32_t ao = (uint32_t)(2 * i) * (COUT * 16), bo = (uint32_t)(2 * i) * (TE * 16);
                    const uint64_t dah = make_desc(a_hi + ao, COUT * 16, 128), dal = make_desc(a_lo + ao, COUT * 16, 128);
                    const uint64_t dbh = make_desc(b_hi + bo, TE * 16, 128), dbl = make_desc(b_lo + bo, TE * 16, 128);
                    mma_tf32(d, dah, dbh, idesc, i > 0);
                    if (!(SGB_ABL & 4)) {
                        mma_tf32(d, dal, dbh, idesc, true);
                        mma_tf32(d, dah, dbl, idesc, true);
                    }
                }
                mma_commit(&bar_tfull[st]);
                if (GRAM) {
                    const uint32_t ht = b_lo + C::TILE_BYTES;
                    const uint32_t dg = tmem + (uint32_t)(C::G_COL0 + gb * 128);
#pragma unroll
                    for (int s = 0; s < ((SGB_ABL & 8) ? 1 : TE / 8); ++s) {              // 8 edges (K) per instruction
                        const uint32_t ko = (uint32_t)(s >> 1) * (GR * 64) + (uint32_t)(s & 1) * 32;
                        const uint64_t da = make_desc_sw(ht + ko, 16, 512, 4);
                        const uint64_t db = make_desc_sw(ht + ko + 8 * 512, 16, 512, 4);      // rows 64.. : hi, ones, zeros
                        mma_tf32(dg, da, db, idesc_g, (t % FLUSH) > 0 || s > 0);
                    }
                    if ((t + 1) % FLUSH == 0 || t == ntiles - 1) mma_commit(&bar_gfull[gb]);
                }
                mma_commit(&bar_empty[st]);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: warp q owns channels 16q + lane (lanes 0..15); one point = 20 consecutive columns
        const int c = warp * 16 + (lane & 15);
        const bool owner = lane < 16;
        const bool up = __ldg(gamma2 + c) > 0.f;
        // running sums over the CTA's tiles as unevaluated fp32 pairs (two-sum): FP64 adds cost ~100 cycles per warp on this part
        // (ncu source view: 10 % of the epilogue's samples sat on the two DADDs per tile)
        float S1h = 0.f, S1l = 0.f, S2h = 0.f, S2l = 0.f;
        auto two_sum = [](float& h, float& l, float x) {
            const float s_ = h + x;
            const float bb = s_ - h;
            l += (h - (s_ - bb)) + (x - bb);
            h = s_;
        };
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_tfull[st], ph);
            fence_after_sync();
            const long long g0 = g_begin + (long long)t * TE;
            const int npts = (int)min((long long)C::PTS, (g_end - g0) / KNN);
            const long long p0 = g0 / KNN;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * C::Z_COL);
            float ps[4] = {0.f, 0.f, 0.f, 0.f}, qs[4] = {0.f, 0.f, 0.f, 0.f};      // four independent chains each for sum z and sum z^2
            // BN2 + LeakyReLU is monotone per channel, increasing iff gamma2 > 0: only that extreme of the 20 pre-activations (and its
            // neighbour slot) is kept.  The extreme is found by a tournament (depth 5 instead of a chain of 19 compare / select pairs; the
            // left operand wins ties, so the FIRST extreme is kept exactly as in a left-to-right scan); two points per iteration share
            // one tcgen05.wait::ld.  Warp-uniform fast paths for the usual cases (all scales positive / all negative).
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
                if (owner) {
                    const size_t o = (size_t)(p0 + pp) * COUT + c;
                    zsel[o] = m[0];
                    if (ARG) ksel[o] = (unsigned char)k[0];
                }
            };
            auto scan = [&](auto better) {
                const int nloop = (SGB_ABL & 2) ? 1 : npts;
#pragma unroll 1
                for (int pp = 0; pp < nloop; pp += 2) {
                    uint32_t a16[16], a4[4], b16[16], b4[4];
                    const bool two = pp + 1 < nloop;         // warp-uniform
                    tmem_ld16_issue(taddr + (uint32_t)(pp * KNN), a16);
                    tmem_ld4_issue(taddr + (uint32_t)(pp * KNN + 16), a4);
                    if (two) {
                        tmem_ld16_issue(taddr + (uint32_t)((pp + 1) * KNN), b16);
                        tmem_ld4_issue(taddr + (uint32_t)((pp + 1) * KNN + 16), b4);
                    }
                    tmem_ld_wait();
                    tmem_ld_pin20(a16, a4);
                    reduce_point(a16, a4, pp, better);
                    if (two) {
                        tmem_ld_pin20(b16, b4);
                        reduce_point(b16, b4, pp + 1, better);
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
                const int seg = t / FLUSH, gb = seg & 1;
                mbar_wait(&bar_gfull[gb], (uint32_t)(seg >> 1) & 1u);
                fence_after_sync();
                float* dst = gslots + ((size_t)blockIdx.x * nflush + seg) * (128 * GN) + (size_t)(warp * 32 + lane) * GN;
                const uint32_t gaddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(C::G_COL0 + gb * 128);
#pragma unroll 1
                for (int c0 = 0; c0 < GN; c0 += 16) {
                    float v[16];
                    tmem_ld16(gaddr + (uint32_t)c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(dst + c0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_gempty[gb]);
            }
        }
        if (owner) {
            part[(size_t)blockIdx.x * 128 + c] = (double)S1h + (double)S1l;
            part[(size_t)blockIdx.x * 128 + 64 + c] = (double)S2h + (double)S2l;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// =================================================================================================================================
// Round 3: BOTH layers on the tensor cores.
//
// The kernel above left the first layer (18 -> 64, BN1 folded in) on the CUDA cores: 16 producer threads per edge, each re-reading
// the edge's two raw rows and spending ~110 instructions on 4 hidden channels (~1,700 per edge), which made the producers the critical
// role of a kernel whose tensor pipe sat at 37-39 %.  Here the first layer is one more small GEMM per tile,
//     D1[128 edge rows, 64 channels] = E[128, 24] * W1s[64, 24]^T          (kind::tf32 x 3, K = 9 differences + 9 centre + 1 + 5 zero)
// with E = (x_j - x_i, x_i, valid, 0...) assembled by ONE thread per edge (hi / lo split, six 16-byte stores) and
// W1s = (scale1 W1 | beta1 - scale1 mean1 | 0): the constant-1 column carries the folded BatchNorm-1 bias, an all-zero row (edge
// outside the CTA's range) gives h = lrelu(0) = 0.  The edges lie on the TMEM LANES of D1, so a producer thread then owns one edge
// row: tcgen05.ld of 32 of its 64 pre-activations (two warps per TMEM quadrant, one per channel half), LeakyReLU, hi / lo split and
// 16-byte stores into the K-major H tile (consecutive lanes = consecutive rows: conflict-free) — ~250 instructions per edge and
// channel half instead of ~1,700.  In the Gram variant the transposed copy is 32 consecutive edges of one channel row per store.
// Roles and pipelines otherwise as above; per tile the producers run  [E(t)] -> [H(t-1)]  so that the first-layer MMA of tile t
// overlaps the second half of the step, and the MMA warp issues  MMA1(t), Gram(t-1), MMA2(t-1).
// Tiles: 120 edges (6 points; rows 120..127 of the M = 128 operand are whatever follows in shared memory — their D1 rows are never
// read) without the Gram, 80 edges (4 points) with it (the transposed tile is single-buffered: the Gram MMAs are issued first).
template <bool GRAM> struct Cfg1 {
    static constexpr int TE = GRAM ? 80 : 120;
    static constexpr int PTS = TE / KNN;
    static constexpr int KE = 24;                                  // K of the first layer
    static constexpr int TILE_BYTES = TE * COUT * 4;               // one K-major H tile (hi or lo)
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
    static constexpr int HT_BYTES = GRAM ? GR * TE * 4 : 0;        // transposed tile (lo, hi, ones / zero rows), one buffer
    static constexpr int E_BYTES = TE * KE * 4;                    // one K-major E tile (hi or lo), one buffer
    static constexpr int W1_BYTES = COUT * KE * 4;
    static constexpr int w2_hi = 0;
    static constexpr int w2_lo = w2_hi + W2_BYTES;
    static constexpr int w1_hi = w2_lo + W2_BYTES;
    static constexpr int w1_lo = w1_hi + W1_BYTES;
    static constexpr int stage0 = w1_lo + W1_BYTES;
    static constexpr int ht0 = stage0 + 2 * STAGE_BYTES;
    static constexpr int e_hi = ht0 + HT_BYTES;
    static constexpr int e_lo = e_hi + E_BYTES;
    static constexpr int RING = GRAM ? 4 : 3;
    static constexpr int RAW_BYTES = TE * 48 + PTS * 48 + TE * 8;
    static constexpr int raw0 = e_lo + E_BYTES;                    // the M = 128 read of the last E chunk runs (128 - TE) * 16 bytes into the ring
    static constexpr int bars = raw0 + RING * RAW_BYTES;           // 17 mbarriers
    static constexpr int tmem_slot = bars + 20 * 8;
    static constexpr int total = tmem_slot + 16;
    static constexpr int Z_COL = GRAM ? 80 : 128;                  // TMEM: z accumulators at 0 and Z_COL
    static constexpr int D1_COL = GRAM ? 160 : 256;                //       first-layer accumulator (64 columns)
    static constexpr int G_COL0 = 256;                             //       Gram accumulators at 256 and 384
};
static_assert(Cfg1<true>::ht0 % 1024 == 0 && Cfg1<true>::stage0 % 1024 == 0, "swizzled tile alignment");
static_assert(Cfg1<true>::bars % 8 == 0 && Cfg1<false>::bars % 8 == 0, "barrier alignment");
static_assert(Cfg1<true>::total + 1024 <= 227 * 1024 && Cfg1<false>::total + 1024 <= 227 * 1024, "shared memory budget");
static_assert((128 - Cfg1<true>::TE) * 16 <= Cfg1<true>::RING * Cfg1<true>::RAW_BYTES, "E overrun stays inside the block");

template <bool ARG, bool GRAM>
__global__ void __launch_bounds__(THREADS, 1)
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
    uint64_t* bar_efull = bar_full + 12;                                     // producers -> MMA: E tile written
    uint64_t* bar_eempty = bar_full + 13;                                    // MMA (commit after the first-layer MMAs) -> producers
    uint64_t* bar_d1full = bar_full + 14;                                    // MMA (same commit) -> producers: D1 complete
    uint64_t* bar_d1empty = bar_full + 15;                                   // producers -> MMA: D1 read into registers
    uint64_t* bar_htempty = bar_full + 16;                                   // MMA (commit after the Gram MMAs) -> producers: transposed tile free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int per = N / gridDim.x, rem = N % gridDim.x;
    const int p_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int p_end = p_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
    const long long g_begin = (long long)p_begin * KNN, g_end = (long long)p_end * KNN;
    const int ntiles = (int)((g_end - g_begin + TE - 1) / TE);

    // ---- one-time setup
    for (int i = tid; i < COUT * COUT; i += THREADS) {              // W2 [c][j] -> K-major canonical tile (rows = c, K = j)
        const int c = i / COUT, j = i % COUT;
        const float w = __ldg(W2 + i);
        const float hi = tf32_hi(w);
        const uint32_t off = tile_off(c, j, COUT);
        *reinterpret_cast<float*>(sm + C::w2_hi + off) = hi;
        *reinterpret_cast<float*>(sm + C::w2_lo + off) = tf32_hi(w - hi);
    }
    for (int i = tid; i < COUT * C::KE; i += THREADS) {             // W1s [c][k]: BN1 scale folded in, column 18 = folded bias
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
    for (int i = tid; i < TE; i += THREADS) {                       // chunk 5 of the E tiles (K = 20..23): zero, never rewritten
        *reinterpret_cast<float4*>(sm + C::e_hi + 5 * (TE * 16) + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sm + C::e_lo + 5 * (TE * 16) + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (GRAM) {                                                     // extra rows of the transposed tile: ones, then zeros
        unsigned char* ht = sm + C::ht0;
        for (int i = tid; i < 16 * TE; i += THREADS) {
            const int r = 128 + i / TE, e = i % TE;
            *reinterpret_cast<float*>(ht + ht_off(r, e)) = (r == 128) ? 1.f : 0.f;
        }
    }
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], EPI_WARPS); mbar_init(&bar_tempty[1], EPI_WARPS);
        mbar_init(&bar_gfull[0], 1); mbar_init(&bar_gfull[1], 1);
        mbar_init(&bar_gempty[0], EPI_WARPS); mbar_init(&bar_gempty[1], EPI_WARPS);
        mbar_init(bar_efull, PROD_WARPS); mbar_init(bar_eempty, 1);
        mbar_init(bar_d1full, 1); mbar_init(bar_d1empty, PROD_WARPS);
        mbar_init(bar_htempty, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp > MMA_WARP) {
        // ================= producers
        const int pw = warp - (MMA_WARP + 1);           // 0..7
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
                asm volatile("cp.async.wait_group %0;" :: "n"(RING - 2) : "memory");     // my copies for tile t (and the indices of tile t + RING - 1) have landed
                asm volatile("bar.sync 1, %0;" :: "n"(PROD_THREADS) : "memory");         // everybody's have, and everybody is done with tile t - 1
                issue_rows(t + RING - 1, staged_index(t + RING - 1));                    // refill the slot tile t - 1 used
                mbar_wait(bar_eempty, ((uint32_t)t & 1u) ^ 1u);
                if (gatherer) {
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
                        *reinterpret_cast<float4*>(sm + C::e_hi + q * (TE * 16) + ptid * 16) = hi;
                        *reinterpret_cast<float4*>(sm + C::e_lo + q * (TE * 16) + ptid * 16) = lo;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_efull);
            }
            if (t >= 1) {
                // ---- H(u): hidden activations of tile u = t - 1 from the first-layer accumulator
                const int u = t - 1, su = u & 1;
                mbar_wait(bar_d1full, (uint32_t)u & 1u);
                fence_after_sync();
                uint32_t va[16], vb[16];
                tmem_ld16_issue(d1addr, va);
                tmem_ld16_issue(d1addr + 16u, vb);
                tmem_ld_wait();
                tmem_ld_pin16(va);
                tmem_ld_pin16(vb);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_d1empty);
                mbar_wait(&bar_empty[su], (((uint32_t)(u >> 1)) & 1u) ^ 1u);
                unsigned char* dst_hi = sm + C::stage0 + su * C::STAGE_BYTES;
                unsigned char* dst_lo = dst_hi + C::TILE_BYTES;
                const bool rowv = hrow < TE;
                if (GRAM) mbar_wait(bar_htempty, ((uint32_t)u & 1u) ^ 1u);      // the Gram MMAs of tile u - 1 were issued before its second layer
                // transposed copy of the Gram variant: rows 0..63 lo, 64..127 hi (row = channel), 32 consecutive edges of one row per store
                unsigned char* dst_t = sm + C::ht0 + (uint32_t)chh * (4 * 512);
                const uint32_t e1 = (uint32_t)((hrow >> 4) * (GR * 64) + (hrow & 3) * 4), eq = (uint32_t)(((hrow & 15) >> 2) << 4);
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
                        if (GRAM) {
                            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                            for (int s4 = 0; s4 < 4; ++s4) {
                                const int ci = 4 * i4 + s4;                  // channel 32 chh + ci = row of the transposed tile
                                const uint32_t o = (uint32_t)((ci >> 3) * 512 + (ci & 7) * 64) + e1 + (eq ^ (uint32_t)(((ci >> 1) & 3) << 4));
                                *reinterpret_cast<float*>(dst_t + o) = lv[s4];
                                *reinterpret_cast<float*>(dst_t + o + 8 * 512) = hv[s4];
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
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer
        const uint32_t idesc1 = make_idesc_tf32(128, COUT, false, false);
        const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
        const uint32_t idesc_g = make_idesc_tf32(128, GN, false, false);
        const uint32_t a_hi = smem_u32(sm + C::w2_hi), a_lo = smem_u32(sm + C::w2_lo);
        const uint32_t w1h = smem_u32(sm + C::w1_hi), w1l = smem_u32(sm + C::w1_lo);
        const uint32_t eh = smem_u32(sm + C::e_hi), el = smem_u32(sm + C::e_lo);
        for (int t = 0; t <= ntiles; ++t) {
            if (t < ntiles) {
                mbar_wait(bar_efull, (uint32_t)t & 1u);
                mbar_wait(bar_d1empty, ((uint32_t)t & 1u) ^ 1u);
                fence_after_sync();
                if (elect_one_sync()) {
                    const uint32_t d1 = tmem + (uint32_t)C::D1_COL;
#pragma unroll
                    for (int i = 0; i < C::KE / 8; ++i) {
                        const uint32_t ao = (uint32_t)(2 * i) * (TE * 16), bo = (uint32_t)(2 * i) * (COUT * 16);
                        const uint64_t dah = make_desc(eh + ao, TE * 16, 128), dal = make_desc(el + ao, TE * 16, 128);
                        const uint64_t dbh = make_desc(w1h + bo, COUT * 16, 128), dbl = make_desc(w1l + bo, COUT * 16, 128);
                        mma_tf32(d1, dah, dbh, idesc1, i > 0);
                        mma_tf32(d1, dal, dbh, idesc1, true);
                        mma_tf32(d1, dah, dbl, idesc1, true);
                    }
                    mma_commit(bar_d1full);
                    mma_commit(bar_eempty);
                }
                __syncwarp();
            }
            if (t >= 1) {
                const int u = t - 1, st = u & 1;
                const uint32_t ph = (uint32_t)(u >> 1) & 1u;
                const int seg = u / FLUSH, gb = seg & 1;
                mbar_wait(&bar_full[st], ph);
                mbar_wait(&bar_tempty[st], ph ^ 1u);
                if (GRAM && u % FLUSH == 0) mbar_wait(&bar_gempty[gb], ((uint32_t)(seg >> 1) & 1u) ^ 1u);
                fence_after_sync();
                if (elect_one_sync()) {
                    const uint32_t b_hi = smem_u32(sm + C::stage0 + st * C::STAGE_BYTES), b_lo = b_hi + C::TILE_BYTES;
                    if (GRAM) {                         // first: the transposed tile has ONE buffer and is free again as soon as these complete
                        const uint32_t ht = smem_u32(sm + C::ht0);
                        const uint32_t dg = tmem + (uint32_t)(C::G_COL0 + gb * 128);
#pragma unroll
                        for (int s = 0; s < TE / 8; ++s) {                               // 8 edges (K) per instruction
                            const uint32_t ko = (uint32_t)(s >> 1) * (GR * 64) + (uint32_t)(s & 1) * 32;
                            const uint64_t da = make_desc_sw(ht + ko, 16, 512, 4);
                            const uint64_t db = make_desc_sw(ht + ko + 8 * 512, 16, 512, 4);      // rows 64.. : hi, ones, zeros
                            mma_tf32(dg, da, db, idesc_g, (u % FLUSH) > 0 || s > 0);
                        }
                        mma_commit(bar_htempty);
                        if ((u + 1) % FLUSH == 0 || u == ntiles - 1) mma_commit(&bar_gfull[gb]);
                    }
                    const uint32_t d = tmem + (uint32_t)(st * C::Z_COL);
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
                    mma_commit(&bar_empty[st]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: warp q owns channels 16q + lane (lanes 0..15); one point = 20 consecutive columns
        const int c = warp * 16 + (lane & 15);
        const bool owner = lane < 16;
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
            mbar_wait(&bar_tfull[st], ph);
            fence_after_sync();
            const long long g0 = g_begin + (long long)t * TE;
            const int npts = (int)min((long long)C::PTS, (g_end - g0) / KNN);
            const long long p0 = g0 / KNN;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * C::Z_COL);
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
                if (owner) {
                    const size_t o = (size_t)(p0 + pp) * COUT + c;
                    zsel[o] = m[0];
                    if (ARG) ksel[o] = (unsigned char)k[0];
                }
            };
            auto scan = [&](auto better) {
#pragma unroll 1
                for (int pp = 0; pp < npts; pp += 2) {
                    uint32_t a16[16], a4[4], b16[16], b4[4];
                    const bool two = pp + 1 < npts;         // warp-uniform
                    tmem_ld16_issue(taddr + (uint32_t)(pp * KNN), a16);
                    tmem_ld4_issue(taddr + (uint32_t)(pp * KNN + 16), a4);
                    if (two) {
                        tmem_ld16_issue(taddr + (uint32_t)((pp + 1) * KNN), b16);
                        tmem_ld4_issue(taddr + (uint32_t)((pp + 1) * KNN + 16), b4);
                    }
                    tmem_ld_wait();
                    tmem_ld_pin20(a16, a4);
                    reduce_point(a16, a4, pp, better);
                    if (two) {
                        tmem_ld_pin20(b16, b4);
                        reduce_point(b16, b4, pp + 1, better);
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
                const int seg = t / FLUSH, gb = seg & 1;
                mbar_wait(&bar_gfull[gb], (uint32_t)(seg >> 1) & 1u);
                fence_after_sync();
                float* dst = gslots + ((size_t)blockIdx.x * nflush + seg) * (128 * GN) + (size_t)(warp * 32 + lane) * GN;
                const uint32_t gaddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(C::G_COL0 + gb * 128);
#pragma unroll 1
                for (int c0 = 0; c0 < GN; c0 += 16) {
                    float v[16];
                    tmem_ld16(gaddr + (uint32_t)c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(dst + c0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_gempty[gb]);
            }
        }
        if (owner) {
            part[(size_t)blockIdx.x * 128 + c] = (double)S1h + (double)S1l;
            part[(size_t)blockIdx.x * 128 + 64 + c] = (double)S2h + (double)S2l;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// mom2 from the reduced Gram accumulator of ec2_tc1_kernel (row of the transposed tile = hidden channel)
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

// mom2 [64*64 + 64] (sum h h^T, sum h) from the reduced Gram accumulator R [128][GN]:
// rows 0..63 = lo*(hi | 1), rows 64..127 = hi*(hi | 1), both in transposed-tile row order -> un-permute
__global__ void __launch_bounds__(64)
mom2_from_gram_kernel(const double* __restrict__ R, double* __restrict__ mom2) {
    __shared__ int s_row[COUT];                         // hidden channel -> row of the transposed tile
    const int j = threadIdx.x;
    s_row[j] = ht_row(j >> 4, j & 15);
    __syncthreads();
    const int rj = s_row[j];
    for (int i = 0; i < COUT; ++i) {
        const int ri = s_row[i];
        // hi_j.hi_i + lo_j.hi_i + hi_j.lo_i
        mom2[j * COUT + i] = R[(64 + rj) * GN + ri] + R[rj * GN + ri] + R[ri * GN + rj];
    }
    mom2[COUT * COUT + j] = R[(64 + rj) * GN + 64] + R[rj * GN + 64];
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
    const int tiles = sgb_div_up(pts * KNN, Cfg<true>::TE);
    return sgb_div_up(tiles > 0 ? tiles : 1, FLUSH);
}
}  // namespace sgb_ectc

// workspace: partial sums [148][128] f64, reduced sums [128] f64, reduced Gram [128*GN] f64, zsel [N,64] f32, ksel [N,64] u8,
// Gram slots [148][nflush][128*GN] f32
size_t sgb_ec2_tc_ws_bytes(int N) {
    using namespace sgb_ectc;
    const int nf = tc_nflush(N, 148) + 1;
    return (size_t)(148 + 1) * 128 * 8 + (size_t)128 * GN * 8 + (size_t)N * 64 * (4 + 1) +
           (size_t)148 * nf * 128 * GN * 4 + 1024;
}

// second layer of MLP3 on the tensor cores: stats2/var2 and out/argk as sgb_edgeconv_fwd produces them; mom2 (optional)
// = second moments of the hidden activations for the backward pass
int sgb_ec2_tc_forward(const float* x12 /*[N,12]: 48-byte padded rows*/, const int* knn, int N, const float* W1, const float* stats1, const float* W2,
                       const float* gamma2, const float* beta2, float* out, unsigned char* argk, float* stats2, float* var2,
                       double* mom2, void* ws, cudaStream_t st) {
    using namespace sgb_ectc;
    unsigned char* w8 = (unsigned char*)ws;
    double* part = (double*)w8;
    double* sums = part + 148 * 128;
    double* gred = sums + 128;
    float* zsel = (float*)(gred + 128 * GN);
    unsigned char* ksel = (unsigned char*)(zsel + (size_t)N * 64);
    float* gslots = (float*)(((uintptr_t)(ksel + (size_t)N * 64) + 255) & ~(uintptr_t)255);
    const bool gram = mom2 != nullptr;
    static const bool old_path = getenv("SGB_EC2_OLD") != nullptr;      // bring-up switch: the round-2 kernel (first layer on the CUDA cores)
    int grid;
    if (!old_path) {
        grid = tc_grid(N, gram ? Cfg1<true>::TE : Cfg1<false>::TE);
        const int nflush = gram ? tc_nflush(N, grid) : 0;
        if (gram) {
            const size_t smem = Cfg1<true>::total + 1024;
            SGB_CUDA(cudaMemsetAsync(gslots, 0, (size_t)grid * nflush * 128 * GN * 4, st));
            SGB_OPT_IN_SMEM((ec2_tc1_kernel<true, true>));
            { ec2_tc1_kernel<true, true><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, gslots, nflush); SGB_COUNT_LAUNCH(); }
            sgb_bn::reduce_partials(gslots, grid * nflush, 128 * GN, gred, st);
            { mom2_from_gram1_kernel<<<1, 64, 0, st>>>(gred, mom2); SGB_COUNT_LAUNCH(); }
        } else if (argk) {
            const size_t smem = Cfg1<false>::total + 1024;
            SGB_OPT_IN_SMEM((ec2_tc1_kernel<true, false>));
            { ec2_tc1_kernel<true, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
        } else {
            const size_t smem = Cfg1<false>::total + 1024;
            SGB_OPT_IN_SMEM((ec2_tc1_kernel<false, false>));
            { ec2_tc1_kernel<false, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
        }
    } else {
    grid = tc_grid(N, gram ? Cfg<true>::TE : Cfg<false>::TE);
    const int nflush = gram ? tc_nflush(N, grid) : 0;
    if (gram) {
        const size_t smem = Cfg<true>::total + 1024;
        SGB_CUDA(cudaMemsetAsync(gslots, 0, (size_t)grid * nflush * 128 * GN * 4, st));
        SGB_OPT_IN_SMEM(ec2_tc_kernel<true, true>);
        { ec2_tc_kernel<true, true><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, gslots, nflush); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(gslots, grid * nflush, 128 * GN, gred, st);
        { mom2_from_gram_kernel<<<1, 64, 0, st>>>(gred, mom2); SGB_COUNT_LAUNCH(); }
    } else if (argk) {
        const size_t smem = Cfg<false>::total + 1024;
        SGB_OPT_IN_SMEM(ec2_tc_kernel<true, false>);
        { ec2_tc_kernel<true, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    } else {
        const size_t smem = Cfg<false>::total + 1024;
        SGB_OPT_IN_SMEM(ec2_tc_kernel<false, false>);
        { ec2_tc_kernel<false, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, gamma2, zsel, ksel, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    }
    }
    sgb_bn::reduce_partials(part, grid, 128, sums, st);
    { bn2_from_sums_kernel<<<1, 64, 0, st>>>(sums, (double)N * KNN, gamma2, beta2, stats2, var2); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)N * 64;
    { ec2_apply_kernel<<<sgb_div_up(total / 4, 256), 256, 0, st>>>(zsel, ksel, stats2, total, out, argk); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

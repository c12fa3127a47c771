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
// Warp roles (480 threads, 1 CTA / SM, persistent over a contiguous range of points):
//   warps 0-3   epilogue: tcgen05.ld of their TMEM sub-partition (an M = 64 accumulator keeps rows 16q..16q+15 on lanes
//               32q..32q+15), statistics + max/min over 20 consecutive columns, stores of the per-point candidates;
//   warp  4     TMEM allocation + the single MMA-issuing thread + tcgen05.commit;
//   warps 5-14  producers: gather the (48-byte padded) rows one tile ahead into registers, first layer + BN1 (folded
//               into the weights) + LeakyReLU on the CUDA cores with packed FFMA2, hi/lo split, 16-byte stores into the
//               canonical no-swizzle K-major tile, fence.proxy.async, mbarrier arrive.
// Pipelines: shared-memory tiles full/empty (2 stages) and TMEM accumulators full/empty (2 buffers), all mbarriers.
#include "common.cuh"
#include "bn_moments.cuh"
#include "edgeconv_common.cuh"
#include "tc_common.cuh"

namespace sgb_ectc {
using namespace sgb_tc;
using sgb_ec::CIN;
using sgb_ec::COUT;
using sgb_ec::KNN;
using sgb_bn::lrelu;

constexpr int EPI_WARPS = 4, PROD_WARPS = 10;
constexpr int MMA_WARP = EPI_WARPS;           // warp 4
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;     // 480
constexpr int PROD_THREADS = PROD_WARPS * 32; // 320
constexpr int W2_BYTES = COUT * COUT * 4;     // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int FLUSH = 32;                     // tiles per Gram segment
constexpr int GN = 80;                        // Gram accumulator columns: 64 hidden + ones + 15 zero rows
constexpr int GR = 144;                       // rows of the transposed tile: 64 lo + 64 hi + 16 extra
constexpr int W1T_STRIDE = 80;                // floats per q-row of the staged first-layer weights (4 parts x (16 + 4 pad))

template <bool GRAM> struct Cfg {
    static constexpr int TE = GRAM ? 80 : 160;                     // edges per tile (TMEM columns per z accumulator)
    static constexpr int PTS = TE / KNN;                           // points per tile
    static constexpr int PPE = PROD_THREADS / TE;                  // producer threads per edge (4 : 2)
    static constexpr int CPT = 16 / PPE;                           // 16-byte chunks (4 hidden channels) per producer thread
    static constexpr int TILE_BYTES = TE * COUT * 4;               // one K-major H tile (hi or lo)
    static constexpr int HT_BYTES = GRAM ? GR * TE * 4 : 0;        // transposed tile (lo, hi, extra rows)
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES + HT_BYTES;  // multiple of 1024 in both variants
    // offsets into the dynamic shared memory block (1024-byte aligned base)
    static constexpr int w2_hi = 0;
    static constexpr int w2_lo = w2_hi + W2_BYTES;
    static constexpr int stage0 = w2_lo + W2_BYTES;
    static constexpr int w1t = stage0 + 2 * STAGE_BYTES;           // [18][W1T_STRIDE] floats: BN1 scale folded in
    static constexpr int b1 = w1t + CIN * W1T_STRIDE * 4;          // [64]: beta1 - scale1 * mean1
    static constexpr int bars = b1 + COUT * 4;                     // 12 mbarriers
    static constexpr int tmem_slot = bars + 12 * 8;
    static constexpr int total = tmem_slot + 16;
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
__global__ void __launch_bounds__(THREADS, 1)
ec2_tc_kernel(const float* __restrict__ x12, const int* __restrict__ knn, int N, const float* __restrict__ W1,
              const float* __restrict__ stats1, const float* __restrict__ W2,
              float* __restrict__ zmax, float* __restrict__ zmin, unsigned short* __restrict__ kk, double* __restrict__ part /*[grid][128]*/,
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
    float* s_w1t = reinterpret_cast<float*>(sm + C::w1t);
    float* s_b1 = reinterpret_cast<float*>(sm + C::b1);
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
    // first layer with the BatchNorm-1 affine folded in:  v1 = (scale1 W1) e + (beta1 - scale1 mean1);
    // staged [q][channel] with a 16-byte pad after every 16 channels so that the producer parts read distinct banks
    for (int i = tid; i < COUT * CIN; i += THREADS) {
        const int c = i / CIN, q = i % CIN;
        s_w1t[q * W1T_STRIDE + (c >> 4) * 20 + (c & 15)] = __ldg(W1 + i) * stats1[128 + c];
    }
    for (int i = tid; i < COUT; i += THREADS) s_b1[i] = fmaf(-stats1[128 + i], stats1[i], stats1[192 + i]);
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
        // ================= producers: thread = (edge row, part of the hidden channels); a warp holds 32/PPE consecutive edges
        constexpr int PPE = C::PPE, CPT = C::CPT;
        const int pw = warp - (MMA_WARP + 1);
        const int er = pw * (32 / PPE) + lane / PPE;    // edge row in the tile
        const int part_id = lane % PPE;                 // 16-byte chunks part_id*CPT .. part_id*CPT + CPT - 1 of the hidden vector
        // software pipeline over tiles: neighbour index two tiles ahead, gathered rows one tile ahead
        float en[CIN];                                  // edge vector of the NEXT tile (in flight during this tile's math)
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
            const bool valid = v_next;
#pragma unroll
            for (int q = 0; q < CIN; ++q) ee[q] = make_float2(en[q], en[q]);
            issue_rows(t + 1, j_next);                  // loads land while this tile is computed
            j_next = issue_index(t + 2);
            mbar_wait(&bar_empty[st], ph ^ 1u);
            unsigned char* dst_hi = sm + C::stage0 + st * C::STAGE_BYTES;
            unsigned char* dst_lo = dst_hi + C::TILE_BYTES;
            unsigned char* dst_t = dst_lo + C::TILE_BYTES;
#pragma unroll 2
            for (int i4 = 0; i4 < CPT; ++i4) {
                const int c4 = part_id * CPT + i4;      // 16-byte chunk = hidden channels 4 c4 .. 4 c4 + 3
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
                    y = make_float4(lrelu(y01.x), lrelu(y01.y), lrelu(y23.x), lrelu(y23.y));
                }
                const float4 hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                const float4 lo = make_float4(tf32_hi(y.x - hi.x), tf32_hi(y.y - hi.y), tf32_hi(y.z - hi.z), tf32_hi(y.w - hi.w));
                const uint32_t off = (uint32_t)c4 * (TE * 16) + (uint32_t)(er >> 3) * 128 + (uint32_t)(er & 7) * 16;
                *reinterpret_cast<float4*>(dst_hi + off) = hi;
                *reinterpret_cast<float4*>(dst_lo + off) = lo;
                if (GRAM) {                             // transposed copy: rows 0..63 lo, 64..127 hi (GRAM: part_id = c4 / 4)
                    const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int r = ht_row(part_id, i4 * 4 + s);
                        *reinterpret_cast<float*>(dst_t + ht_off(r, er)) = lv[s];
                        *reinterpret_cast<float*>(dst_t + ht_off(64 + r, er)) = hv[s];
                    }
                }
            }
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
            if (lane == 0) {
                const uint32_t b_hi = smem_u32(sm + C::stage0 + st * C::STAGE_BYTES), b_lo = b_hi + C::TILE_BYTES;
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
                if (GRAM) {
                    const uint32_t ht = b_lo + C::TILE_BYTES;
                    const uint32_t dg = tmem + (uint32_t)(C::G_COL0 + gb * 128);
#pragma unroll
                    for (int s = 0; s < TE / 8; ++s) {              // 8 edges (K) per instruction
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
        double S1 = 0.0, S2 = 0.0;
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_tfull[st], ph);
            fence_after_sync();
            const long long g0 = g_begin + (long long)t * TE;
            const int npts = (int)min((long long)C::PTS, (g_end - g0) / KNN);
            const long long p0 = g0 / KNN;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * C::Z_COL);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int pp = 0; pp < npts; ++pp) {
                float v[16], u[4];
                tmem_ld16(taddr + (uint32_t)(pp * KNN), v);
                tmem_ld4(taddr + (uint32_t)(pp * KNN + 16), u);
                float mx = v[0], mn = v[0];
                int kx = 0, kn = 0;
                s1 += v[0]; s2 = fmaf(v[0], v[0], s2);
#pragma unroll
                for (int i = 1; i < KNN; ++i) {
                    const float z = i < 16 ? v[i] : u[i - 16];
                    s1 += z; s2 = fmaf(z, z, s2);
                    if (ARG) {
                        if (z > mx) { mx = z; kx = i; }
                        if (z < mn) { mn = z; kn = i; }
                    } else {
                        mx = fmaxf(mx, z); mn = fminf(mn, z);
                    }
                }
                if (owner) {
                    const size_t o = (size_t)(p0 + pp) * COUT + c;
                    zmax[o] = mx; zmin[o] = mn;
                    if (ARG) kk[o] = (unsigned short)(kx | (kn << 8));
                }
            }
            S1 += (double)s1; S2 += (double)s2;
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
            part[(size_t)blockIdx.x * 128 + c] = S1;
            part[(size_t)blockIdx.x * 128 + 64 + c] = S2;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
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

// out[p, c] = lrelu(BN2(z*)), z* = max or min candidate by the sign of the BN scale; argk = its neighbour slot
__global__ void __launch_bounds__(256)
ec2_apply_kernel(const float* __restrict__ zmax, const float* __restrict__ zmin, const unsigned short* __restrict__ kk,
                 const float* __restrict__ stats2, long long total, float* __restrict__ out, unsigned char* __restrict__ argk) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= total) return;
    const int c = (int)(i4 & 63);
    const float4 a = *reinterpret_cast<const float4*>(zmax + i4);
    const float4 b = *reinterpret_cast<const float4*>(zmin + i4);
    const ushort4 kq = argk ? *reinterpret_cast<const ushort4*>(kk + i4) : make_ushort4(0, 0, 0, 0);
    const float za[4] = {a.x, a.y, a.z, a.w}, zb[4] = {b.x, b.y, b.z, b.w};
    const unsigned short kv[4] = {kq.x, kq.y, kq.z, kq.w};
    float o[4];
    unsigned char ak[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float mean = stats2[c + q], scale = stats2[128 + c + q], beta = stats2[192 + c + q];
        const bool up = scale > 0.f;
        const float z = up ? za[q] : zb[q];
        o[q] = lrelu(fmaf(z - mean, scale, beta));
        ak[q] = scale == 0.f ? (unsigned char)0 : (unsigned char)(up ? (kv[q] & 0xff) : (kv[q] >> 8));
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

// workspace: partial sums [148][128] f64, reduced sums [128] f64, reduced Gram [128*GN] f64, zmax, zmin [N,64] f32,
// kk [N,64] u16, Gram slots [148][nflush][128*GN] f32
size_t sgb_ec2_tc_ws_bytes(int N) {
    using namespace sgb_ectc;
    const int nf = tc_nflush(N, 148) + 1;
    return (size_t)(148 + 1) * 128 * 8 + (size_t)128 * GN * 8 + (size_t)N * 64 * (4 + 4 + 2) +
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
    float* zmax = (float*)(gred + 128 * GN);
    float* zmin = zmax + (size_t)N * 64;
    unsigned short* kk = (unsigned short*)(zmin + (size_t)N * 64);
    float* gslots = (float*)(((uintptr_t)(kk + (size_t)N * 64) + 255) & ~(uintptr_t)255);
    const bool gram = mom2 != nullptr;
    const int grid = tc_grid(N, gram ? Cfg<true>::TE : Cfg<false>::TE);
    const int nflush = gram ? tc_nflush(N, grid) : 0;
    if (gram) {
        const size_t smem = Cfg<true>::total + 1024;
        SGB_CUDA(cudaMemsetAsync(gslots, 0, (size_t)grid * nflush * 128 * GN * 4, st));
        SGB_OPT_IN_SMEM(ec2_tc_kernel<true, true>);
        { ec2_tc_kernel<true, true><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, zmax, zmin, kk, part, gslots, nflush); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(gslots, grid * nflush, 128 * GN, gred, st);
        { mom2_from_gram_kernel<<<1, 64, 0, st>>>(gred, mom2); SGB_COUNT_LAUNCH(); }
    } else if (argk) {
        const size_t smem = Cfg<false>::total + 1024;
        SGB_OPT_IN_SMEM(ec2_tc_kernel<true, false>);
        { ec2_tc_kernel<true, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, zmax, zmin, kk, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    } else {
        const size_t smem = Cfg<false>::total + 1024;
        SGB_OPT_IN_SMEM(ec2_tc_kernel<false, false>);
        { ec2_tc_kernel<false, false><<<grid, THREADS, smem, st>>>(x12, knn, N, W1, stats1, W2, zmax, zmin, kk, part, nullptr, 0); SGB_COUNT_LAUNCH(); }
    }
    sgb_bn::reduce_partials(part, grid, 128, sums, st);
    { bn2_from_sums_kernel<<<1, 64, 0, st>>>(sums, (double)N * KNN, gamma2, beta2, stats2, var2); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)N * 64;
    { ec2_apply_kernel<<<sgb_div_up(total / 4, 256), 256, 0, st>>>(zmax, zmin, kk, stats2, total, out, argk); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// a20 backward on the tensor cores: gradients of the rigid KPConv (kpconv/kernels/convolution_ops.py:240-247, the chain
// tf.matmul(all_weights, neighborhood_features) -> tf.matmul(weighted_features, K_values) -> reduce_sum) with respect to the
// features and K_values.  SIMT version: kpconv_bwd_kernel in kpconv.cu (12x the forward time: 7.3 ms against 0.61 ms at
// 119k x 64 x 64, profiles/r02t_kernels.json), which stays the path for the deformable operator and for shapes outside the limits below.
//
//   out[i, :] = sum_k WF[i, k, :] . K_values[k],      WF[i, k, c] = sum_j h_k(y_j - x_i) f_j[c]
//   dK[k, c, o]  = sum_i WF[i, k, c] g[i, o]                                            (1)  contraction over the QUERIES
//   GW[i, k, c]  = sum_o g[i, o] K_values[k, c, o]                                      (2)  contraction over Cout
//   dfeat[j, c] += sum_k h_k(y_j - x_i) GW[i, k, c]     for every neighbour j of query i     scatter through the neighbour lists
//
// Two kernels, both built from the forward kernel's producer (kpconv_tc.cu: neighbour ids and relative coordinates of a
// warp's queries in registers, per kernel point one ballot of the non-zero influences, only those feature rows touched) and
// both contracting as TF32 x 3 (tc_common.cuh) with fp32 accumulation in TMEM:
//
//  * kpconv_bwd_w_tc_kernel  (1): persistent CTAs.  A tile is 64 queries; the producers rebuild WF exactly as the forward
//    does but store it TRANSPOSED, rows = 128 consecutive (kernel point, channel) pairs of a "stage", K = the 64 queries
//    (K-major SWIZZLE_128B; a lane owns 4 rows and writes the 4 queries of its warp as ONE 16-byte chunk per row ->
//    conflict-free), next to the transposed gradient tile g^T [Cout, 64 queries].  D[128, Cout] += WF^T g accumulates in TMEM
//    over ALL tiles of the CTA: ceil(K Cin / 128) stages x Cout columns.  When that exceeds the 512 TMEM columns the stages
//    are split into passes that run as separate CTA groups of the same launch (a pass only evaluates its own kernel
//    points, so no gather is repeated).  The per-CTA partial sums are written once and summed in a fixed order
//    (kpconv_bwd_reduce): dK is deterministic.
//  * kpconv_bwd_f_tc_kernel  (2) + scatter: one CTA per 128 queries.  The gradient tile g [128, Cout] (hi / lo) is the A
//    operand for the whole tile; per chunk (kernel point, block of CW input channels) the matching K_values block
//    [CW, Cout] (pre-split, pre-swizzled image, ONE cp.async.bulk per chunk) is the B operand and D[128 queries, CW] lands
//    in one of two TMEM buffers.  The 16 consumer warps move it through shared memory into "lanes own channels" order and
//    scatter  h_k * GW  to the neighbours' gradient rows with one coalesced red.global.add per (query, neighbour with a
//    non-zero influence, chunk) - the mirror image of the forward's gather.  As with any scatter-add the summation order
//    across queries is not fixed (documented in DESIGN.md); dK is.
// Limits: Cin in {32, 64, 128}, Cout a multiple of 32 up to 128, K <= 32, rows of at most 64 neighbours.
#include "common.cuh"
#include "tc_common.cuh"

namespace sgb_kpbt {
using namespace sgb_tc;

constexpr int PROD_WARPS = 16;
constexpr int THREADS = (PROD_WARPS + 1) * 32;        // 544
constexpr int MAXK = 32;
constexpr uint32_t NB_MASK = 0x03ffffffu;             // neighbour id in the low 26 bits, closest kernel point above
enum { INFL_LINEAR = 0, INFL_CONSTANT = 1, INFL_GAUSSIAN = 2 };

struct Args {
    const float* q; const float* s; const int* idx; const float* feat; const float* kpts; const float* g;
    const unsigned char* bimg; float* gfeat; float* dk_part;
    int n, n0, W, Cin, Cout, K;
    float extent; int influence; int closest;
    int n_tiles, cpp, spp, ns;                          // dK kernel: tiles of 64 queries, CTAs per pass, stages per pass, stages in total
    int nx;                                             // feature kernel: 1 or 2 transpose buffers
};

// neighbour ids (+ closest kernel point) and coordinates relative to the query, for the QPW queries of a warp: lane j <-> neighbours j, j + 32
template <int QPW, int NBL>
__device__ __forceinline__ void load_neighbours(const Args& a, int qbase, const float* s_kp, int lane,
                                                int (&nbp)[QPW][NBL], float (&rx)[QPW][NBL], float (&ry)[QPW][NBL], float (&rz)[QPW][NBL]) {
#pragma unroll
    for (int t = 0; t < QPW; ++t) {
        const int qi = qbase + t;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (qi < a.n) { qx = __ldg(a.q + (size_t)qi * 3); qy = __ldg(a.q + (size_t)qi * 3 + 1); qz = __ldg(a.q + (size_t)qi * 3 + 2); }
#pragma unroll
        for (int h = 0; h < NBL; ++h) {
            const int j = h * 32 + lane;
            int nb = -1;
            if (qi < a.n && j < a.W) {
                const int v = __ldg(a.idx + (size_t)qi * a.W + j);
                if (v >= 0 && v < a.n0) nb = v;
            }
            float x = 0.f, y = 0.f, z = 0.f;
            if (nb >= 0) {
                x = __ldg(a.s + (size_t)nb * 3) - qx; y = __ldg(a.s + (size_t)nb * 3 + 1) - qy; z = __ldg(a.s + (size_t)nb * 3 + 2) - qz;
                if (a.closest) {
                    float best = INFINITY; int bk = 0;
                    for (int k = 0; k < a.K; ++k) {
                        const float dx = x - s_kp[k * 3], dy = y - s_kp[k * 3 + 1], dz = z - s_kp[k * 3 + 2];
                        const float sq = dx * dx + dy * dy + dz * dz;
                        if (sq < best) { best = sq; bk = k; }
                    }
                    nb |= bk << 26;
                }
            }
            nbp[t][h] = nb; rx[t][h] = x; ry[t][h] = y; rz[t][h] = z;
        }
    }
}

// influence of kernel point k on one neighbour (convolution_ops.py:194-229), same expression as the forward kernel
__device__ __forceinline__ float influence_of(const Args& a, int me, float x, float y, float z, float kx, float ky, float kz, int k,
                                              float inv_extent, float inv_gauss) {
    if (me < 0) return 0.f;
    const float dx = x - kx, dy = y - ky, dz = z - kz;
    const float sq = dx * dx + dy * dy + dz * dz;
    float w;
    if (a.influence == INFL_LINEAR) w = fmaxf(1.f - sqrtf(sq) * inv_extent, 0.f);
    else if (a.influence == INFL_CONSTANT) w = 1.f;
    else w = expf(-sq * inv_gauss);
    if (a.closest && (me >> 26) != k) w = 0.f;
    return w;
}

// ================================================================================================ (1) dK
// CPL = Cin / 32 channels per lane; a stage holds KPS = 4 / CPL kernel points = 128 rows; row (idx * 32 + lane) of a stage is
// kernel point idx / CPL, channel lane * CPL + idx % CPL (a lane's channels are consecutive -> one vector load per gathered row)
constexpr int WQ = 64;                                // queries per tile = K extent of the MMAs
constexpr int WQPW = WQ / PROD_WARPS;                 // 4 queries per warp = one 16-byte K chunk
constexpr int WA_TILE = 128 * WQ * 4;                 // 32 KB: one of (hi, lo)
constexpr int WA_STAGE = 2 * WA_TILE;

template <int CPL, int NBL>
__global__ void __launch_bounds__(THREADS, 1)
kpconv_bwd_w_tc_kernel(Args a) {
    constexpr int KPS = 4 / CPL;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int G_TILE = a.Cout * WQ * 4;
    unsigned char* sA = sm;                                       // 2 stages x (hi, lo)
    unsigned char* sG = sA + 2 * WA_STAGE;                        // (hi, lo)
    float* s_kp = reinterpret_cast<float*>(sG + 2 * G_TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kp + MAXK * 4);
    uint64_t* full_a = bars;            // [2] producers -> MMA, 16 arrivals
    uint64_t* empty_a = bars + 2;       // [2] MMA commit -> producers
    uint64_t* full_g = bars + 4;        // producers -> MMA, 16 arrivals
    uint64_t* empty_g = bars + 5;       // MMA commit (last stage of a tile) -> producers
    uint64_t* done = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pass = blockIdx.x / a.cpp, j = blockIdx.x % a.cpp;
    const int st0 = pass * a.spp, nst = min(a.spp, a.ns - st0);   // this CTA's stages
    const int my_tiles = (a.n_tiles - j + a.cpp - 1) / a.cpp;     // >= 1 (host: cpp <= n_tiles)

    for (int i = tid; i < a.K * 3; i += THREADS) s_kp[i] = __ldg(a.kpts + i);
    if (tid == 0) {
        mbar_init(&full_a[0], PROD_WARPS); mbar_init(&full_a[1], PROD_WARPS);
        mbar_init(&empty_a[0], 1); mbar_init(&empty_a[1], 1);
        mbar_init(full_g, PROD_WARPS); mbar_init(empty_g, 1); mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == PROD_WARPS) tmem_alloc(tmem_slot, 512u);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == PROD_WARPS) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_tf32(128, a.Cout, false, false);
            const uint32_t g_hi = smem_u32(sG), g_lo = g_hi + (uint32_t)G_TILE;
            int it = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                mbar_wait(full_g, (uint32_t)(ti & 1));
                for (int sl = 0; sl < nst; ++sl, ++it) {
                    const int b = it & 1;
                    mbar_wait(&full_a[b], (uint32_t)((it >> 1) & 1));
                    fence_after_sync();
                    const uint32_t a_hi = smem_u32(sA + b * WA_STAGE), a_lo = a_hi + WA_TILE;
                    const uint32_t d = tmem + (uint32_t)(sl * a.Cout);
#pragma unroll
                    for (int i = 0; i < WQ / 8; ++i) {
                        const uint32_t ao = (uint32_t)((i >> 2) * (128 * 128) + (i & 3) * 32);
                        const uint32_t bo = (uint32_t)((i >> 2) * (a.Cout * 128) + (i & 3) * 32);
                        const uint64_t dah = make_desc_sw(a_hi + ao, 16, 1024, 2), dal = make_desc_sw(a_lo + ao, 16, 1024, 2);
                        const uint64_t dbh = make_desc_sw(g_hi + bo, 16, 1024, 2), dbl = make_desc_sw(g_lo + bo, 16, 1024, 2);
                        mma_tf32(d, dah, dbh, idesc, ti > 0 || i > 0);
                        mma_tf32(d, dal, dbh, idesc, true);
                        mma_tf32(d, dah, dbl, idesc, true);
                    }
                    mma_commit(&empty_a[b]);
                }
                mma_commit(empty_g);
            }
            mma_commit(done);
        }
    } else {
        // ------------------------------------------------------------------ producers
        const float inv_extent = 1.f / a.extent;
        const float sigma = a.extent * 0.3f;
        const float inv_gauss = 1.f / (2.f * sigma * sigma + 1e-9f);
        int it = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int q0 = (j + ti * a.cpp) * WQ;
            const int qb = q0 + warp * WQPW;
            int nbp[WQPW][NBL];
            float rx[WQPW][NBL], ry[WQPW][NBL], rz[WQPW][NBL];
            load_neighbours<WQPW, NBL>(a, qb, s_kp, lane, nbp, rx, ry, rz);
            // g^T: rows = output channel, K = the tile's queries; this warp writes the 16-byte chunk of its 4 queries
            if (ti > 0) mbar_wait(empty_g, (uint32_t)((ti - 1) & 1));
            for (int o = lane; o < a.Cout; o += 32) {
                float v[WQPW];
#pragma unroll
                for (int t = 0; t < WQPW; ++t) v[t] = (qb + t < a.n) ? __ldg(a.g + (size_t)(qb + t) * a.Cout + o) : 0.f;
                const float h0 = tf32_hi(v[0]), h1 = tf32_hi(v[1]), h2 = tf32_hi(v[2]), h3 = tf32_hi(v[3]);
                const uint32_t off = sw128_off(o, warp * WQPW, a.Cout);
                *reinterpret_cast<float4*>(sG + off) = make_float4(h0, h1, h2, h3);
                *reinterpret_cast<float4*>(sG + G_TILE + off) = make_float4(tf32_hi(v[0] - h0), tf32_hi(v[1] - h1), tf32_hi(v[2] - h2), tf32_hi(v[3] - h3));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_g);

            for (int sl = 0; sl < nst; ++sl, ++it) {
                const int b = it & 1;
                if (it >= 2) mbar_wait(&empty_a[b], (uint32_t)(((it >> 1) - 1) & 1));
                float acc[WQPW][4];
#pragma unroll
                for (int t = 0; t < WQPW; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
#pragma unroll
                for (int kl = 0; kl < KPS; ++kl) {
                    const int k = (st0 + sl) * KPS + kl;
                    if (k >= a.K) continue;                                   // warp-uniform: the last stage may be partly empty
                    const float kx = s_kp[k * 3], ky = s_kp[k * 3 + 1], kz = s_kp[k * 3 + 2];
                    const float* fcol = a.feat + lane * CPL;
#pragma unroll
                    for (int t = 0; t < WQPW; ++t) {
#pragma unroll
                        for (int h = 0; h < NBL; ++h) {
                            const int me = nbp[t][h];
                            const float w = influence_of(a, me, rx[t][h], ry[t][h], rz[t][h], kx, ky, kz, k, inv_extent, inv_gauss);
                            unsigned m = __ballot_sync(SGB_FULL_MASK, w != 0.f);
                            while (m) {                               // warp-uniform; 4 feature rows in flight
                                const int l0 = __ffs(m) - 1;
                                float wl[4]; int nl[4];
#pragma unroll
                                for (int gq = 0; gq < 4; ++gq) {
                                    const int l = m ? __ffs(m) - 1 : l0;
                                    const bool live = m != 0;
                                    m &= m - 1;
                                    wl[gq] = __shfl_sync(SGB_FULL_MASK, w, l);
                                    nl[gq] = __shfl_sync(SGB_FULL_MASK, me, l) & NB_MASK;
                                    if (!live) wl[gq] = 0.f;
                                }
                                float f[4][CPL];
#pragma unroll
                                for (int gq = 0; gq < 4; ++gq) {
                                    const float* src = fcol + (size_t)nl[gq] * a.Cin;
                                    if (CPL == 4) {
                                        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                                        f[gq][0] = v.x; f[gq][1 % CPL] = v.y; f[gq][2 % CPL] = v.z; f[gq][3 % CPL] = v.w;
                                    } else if (CPL == 2) {
                                        const float2 v = __ldg(reinterpret_cast<const float2*>(src));
                                        f[gq][0] = v.x; f[gq][1 % CPL] = v.y;
                                    } else {
                                        f[gq][0] = __ldg(src);
                                    }
                                }
#pragma unroll
                                for (int gq = 0; gq < 4; ++gq)
#pragma unroll
                                    for (int v = 0; v < CPL; ++v) acc[t][kl * CPL + v] = fmaf(wl[gq], f[gq][v], acc[t][kl * CPL + v]);
                            }
                        }
                    }
                }
                unsigned char* Ahi = sA + b * WA_STAGE;
#pragma unroll
                for (int idx = 0; idx < 4; ++idx) {
                    const float h0 = tf32_hi(acc[0][idx]), h1 = tf32_hi(acc[1][idx]), h2 = tf32_hi(acc[2][idx]), h3 = tf32_hi(acc[3][idx]);
                    const uint32_t off = sw128_off(idx * 32 + lane, warp * WQPW, 128);
                    *reinterpret_cast<float4*>(Ahi + off) = make_float4(h0, h1, h2, h3);
                    *reinterpret_cast<float4*>(Ahi + WA_TILE + off) =
                        make_float4(tf32_hi(acc[0][idx] - h0), tf32_hi(acc[1][idx] - h1), tf32_hi(acc[2][idx] - h2), tf32_hi(acc[3][idx] - h3));
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_a[b]);
            }
        }

        // ------------------------------------------------------------------ epilogue: this CTA's partial dK rows
        mbar_wait(done, 0);
        fence_after_sync();
        const int quarter = warp & 3;                                          // TMEM lanes 32 * quarter + lane = row idx * 32 + lane, idx = quarter
        const int kl = quarter / CPL, v = quarter % CPL, ch = lane * CPL + v;
        float* part = a.dk_part + (size_t)j * a.K * a.Cin * a.Cout;
        for (int sl = 0; sl < nst; ++sl) {
            const int k = (st0 + sl) * KPS + kl;
            for (int gq = warp >> 2; gq < (a.Cout >> 4); gq += 4) {
                float d[16];
                tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * a.Cout + gq * 16), d);
                if (k < a.K) {
                    float* dst = part + ((size_t)k * a.Cin + ch) * a.Cout + gq * 16;
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        *reinterpret_cast<float4*>(dst + x * 4) = make_float4(d[x * 4], d[x * 4 + 1], d[x * 4 + 2], d[x * 4 + 3]);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == PROD_WARPS) tmem_dealloc(tmem, 512u);
}

// ================================================================================================ (2) dfeat
constexpr int FQ = 128;                               // queries per tile = TMEM lanes
constexpr int FQPW = FQ / PROD_WARPS;                 // 8 queries per warp

// K_values [K][Cin][Cout] -> per chunk (kernel point k, channel block cb) the shared-memory image of the B operand of (2):
// rows = input channel within the block, columns (K of the MMA) = output channel, hi tile followed by lo tile
__global__ void prep_kernel(const float* __restrict__ kval, int K, int Cin, int Cout, int CW, unsigned char* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Cin * Cout) return;
    const int o = i % Cout, cin = (i / Cout) % Cin, k = i / (Cout * Cin);
    const int ncb = Cin / CW, cb = cin / CW, row = cin % CW;
    const size_t tile = (size_t)CW * Cout * 4;
    unsigned char* base = img + (size_t)(k * ncb + cb) * 2 * tile;
    const uint32_t off = sw128_off(row, o, CW);
    const float v = __ldg(kval + i), hi = tf32_hi(v);
    *reinterpret_cast<float*>(base + off) = hi;
    *reinterpret_cast<float*>(base + tile + off) = tf32_hi(v - hi);
}

template <int CPL, int NBL>
__global__ void __launch_bounds__(THREADS, 1)
kpconv_bwd_f_tc_kernel(Args a) {
    constexpr int CW = 32 * CPL;                      // input channels per chunk = N of the MMAs
    constexpr int XS = CW + 4;                        // row stride of the transpose buffers (floats)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int G_TILE = FQ * a.Cout * 4, B_TILE = CW * a.Cout * 4, B_STAGE = 2 * B_TILE;
    unsigned char* sG = sm;                                                   // (hi, lo) [128 queries][Cout]
    unsigned char* sB = sG + 2 * G_TILE;                                      // 2 stages x (hi, lo) [CW][Cout]
    float* sX = reinterpret_cast<float*>(sB + 2 * B_STAGE);                   // nx x [128][XS]
    float* s_kp = sX + (size_t)a.nx * FQ * XS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kp + MAXK * 4);
    uint64_t* full_g = bars;            // consumers -> MMA: gradient tile staged, 16 arrivals
    uint64_t* full_b = bars + 1;        // [2] TMA bytes
    uint64_t* dfull = bars + 3;         // [2] MMA commit -> consumers (and the TMA issuer: the B stage is free again)
    uint64_t* dempty = bars + 5;        // [2] consumers -> MMA: TMEM buffer drained, 16 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FQ;
    const int ncb = a.Cin / CW, NC = a.K * ncb;
    constexpr uint32_t TCOLS = 2 * CW;                                        // 64 or 128 TMEM columns

    for (int i = tid; i < a.K * 3; i += THREADS) s_kp[i] = __ldg(a.kpts + i);
    if (tid == 0) {
        mbar_init(full_g, PROD_WARPS);
        mbar_init(&full_b[0], 1); mbar_init(&full_b[1], 1);
        mbar_init(&dfull[0], 1); mbar_init(&dfull[1], 1);
        mbar_init(&dempty[0], PROD_WARPS); mbar_init(&dempty[1], PROD_WARPS);
        mbar_fence_init();
    }
    if (warp == PROD_WARPS) tmem_alloc(tmem_slot, TCOLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == PROD_WARPS) {
        // ------------------------------------------------------------------ TMA + MMA issuer
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_tf32(FQ, CW, false, false);
            for (int c = 0; c < 2 && c < NC; ++c) {
                mbar_expect_tx(&full_b[c], (uint32_t)B_STAGE);
                bulk_g2s(sB + c * B_STAGE, a.bimg + (size_t)c * B_STAGE, (uint32_t)B_STAGE, &full_b[c]);
            }
            mbar_wait(full_g, 0);
            const uint32_t g_hi = smem_u32(sG), g_lo = g_hi + (uint32_t)G_TILE;
            for (int c = 0; c < NC; ++c) {
                const int s = c & 1;
                const uint32_t par = (uint32_t)((c >> 1) & 1);
                mbar_wait(&full_b[s], par);
                if (c >= 2) mbar_wait(&dempty[s], (uint32_t)(((c >> 1) - 1) & 1));
                fence_after_sync();
                const uint32_t b_hi = smem_u32(sB + s * B_STAGE), b_lo = b_hi + (uint32_t)B_TILE;
                const uint32_t d = tmem + (uint32_t)(s * CW);
                for (int i = 0; i < (a.Cout >> 3); ++i) {
                    const uint32_t ao = (uint32_t)((i >> 2) * (FQ * 128) + (i & 3) * 32);
                    const uint32_t bo = (uint32_t)((i >> 2) * (CW * 128) + (i & 3) * 32);
                    const uint64_t dah = make_desc_sw(g_hi + ao, 16, 1024, 2), dal = make_desc_sw(g_lo + ao, 16, 1024, 2);
                    const uint64_t dbh = make_desc_sw(b_hi + bo, 16, 1024, 2), dbl = make_desc_sw(b_lo + bo, 16, 1024, 2);
                    mma_tf32(d, dah, dbh, idesc, i > 0);
                    mma_tf32(d, dal, dbh, idesc, true);
                    mma_tf32(d, dah, dbl, idesc, true);
                }
                mma_commit(&dfull[s]);
                if (c + 2 < NC) {                       // refill this B stage as soon as its MMAs have retired
                    mbar_wait(&dfull[s], par);
                    mbar_expect_tx(&full_b[s], (uint32_t)B_STAGE);
                    bulk_g2s(sB + s * B_STAGE, a.bimg + (size_t)(c + 2) * B_STAGE, (uint32_t)B_STAGE, &full_b[s]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        // gradient tile: this warp's 8 rows, hi / lo split, K-major SWIZZLE_128B (K = output channel)
#pragma unroll 1
        for (int t = 0; t < FQPW; ++t) {
            const int r = warp * FQPW + t, qi = q0 + r;
            for (int c4 = lane; c4 < (a.Cout >> 2); c4 += 32) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qi < a.n) v = __ldg(reinterpret_cast<const float4*>(a.g + (size_t)qi * a.Cout) + c4);
                const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const uint32_t off = sw128_off(r, c4 * 4, FQ);
                *reinterpret_cast<float4*>(sG + off) = hi;
                *reinterpret_cast<float4*>(sG + G_TILE + off) = make_float4(tf32_hi(v.x - hi.x), tf32_hi(v.y - hi.y), tf32_hi(v.z - hi.z), tf32_hi(v.w - hi.w));
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_g);

        int nbp[FQPW][NBL];
        float rx[FQPW][NBL], ry[FQPW][NBL], rz[FQPW][NBL];
        load_neighbours<FQPW, NBL>(a, q0 + warp * FQPW, s_kp, lane, nbp, rx, ry, rz);
        const float inv_extent = 1.f / a.extent;
        const float sigma = a.extent * 0.3f;
        const float inv_gauss = 1.f / (2.f * sigma * sigma + 1e-9f);
        const int quarter = warp & 3, cgp = warp >> 2;
        constexpr int CPW = CW / 4;                                           // columns a warp moves out of TMEM: 16 or 8

        for (int c = 0; c < NC; ++c) {
            const int s = c & 1;
            const int k = c / ncb, cb = c - k * ncb;
            float* X = sX + (size_t)(a.nx == 2 ? s : 0) * FQ * XS;
            mbar_wait(&dfull[s], (uint32_t)((c >> 1) & 1));
            fence_after_sync();
            {
                float* xr = X + (size_t)(quarter * 32 + lane) * XS + cgp * CPW;
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * CW + cgp * CPW);
                if (CPW == 16) {
                    float d[16];
                    tmem_ld16(taddr, d);
#pragma unroll
                    for (int x = 0; x < 4; ++x) *reinterpret_cast<float4*>(xr + x * 4) = make_float4(d[x * 4], d[x * 4 + 1], d[x * 4 + 2], d[x * 4 + 3]);
                } else {
                    float d[8];
                    tmem_ld8(taddr, d);
#pragma unroll
                    for (int x = 0; x < 2; ++x) *reinterpret_cast<float4*>(xr + x * 4) = make_float4(d[x * 4], d[x * 4 + 1], d[x * 4 + 2], d[x * 4 + 3]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[s]);
            asm volatile("bar.sync 1, 512;" ::: "memory");                    // the 16 consumer warps: X is complete

            const float kx = s_kp[k * 3], ky = s_kp[k * 3 + 1], kz = s_kp[k * 3 + 2];
            float* gcol = a.gfeat + cb * CW + lane * CPL;
#pragma unroll
            for (int t = 0; t < FQPW; ++t) {
                const float* xr = X + (size_t)(warp * FQPW + t) * XS + lane * CPL;
                float x[CPL];
                if (CPL == 2) { const float2 v = *reinterpret_cast<const float2*>(xr); x[0] = v.x; x[CPL - 1] = v.y; }
                else x[0] = xr[0];
#pragma unroll
                for (int h = 0; h < NBL; ++h) {
                    const int me = nbp[t][h];
                    const float w = influence_of(a, me, rx[t][h], ry[t][h], rz[t][h], kx, ky, kz, k, inv_extent, inv_gauss);
                    unsigned m = __ballot_sync(SGB_FULL_MASK, w != 0.f);
                    while (m) {                                               // warp-uniform
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const float wl = __shfl_sync(SGB_FULL_MASK, w, l);
                        const int nl = __shfl_sync(SGB_FULL_MASK, me, l) & NB_MASK;
                        float* dst = gcol + (size_t)nl * a.Cin;
                        if (CPL == 2) atomicAdd(reinterpret_cast<float2*>(dst), make_float2(wl * x[0], wl * x[CPL - 1]));
                        else atomicAdd(dst, wl * x[0]);
                    }
                }
            }
            if (a.nx == 1) asm volatile("bar.sync 1, 512;" ::: "memory");     // single buffer: nobody may refill X before all have read it
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == PROD_WARPS) tmem_dealloc(tmem, TCOLS);
}

struct Plan { int cpl_w, cpl_f, nbl, ns, spp, passes, cpp, nx; size_t smem_w, smem_f; };
inline bool plan(int n, int W, int Cin, int Cout, int K, int n0, Plan& p) {
    if (K < 1 || K > MAXK || W < 0 || W > 64 || !(Cin == 32 || Cin == 64 || Cin == 128) || Cout < 32 || (Cout & 31) || Cout > 128) return false;
    if (n0 > (int)NB_MASK || n <= 0) return false;
    p.nbl = W <= 32 ? 1 : 2;
    // (1) dK
    p.cpl_w = Cin / 32;
    const int kps = 4 / p.cpl_w;
    p.ns = (K + kps - 1) / kps;
    p.spp = 512 / Cout;
    if (p.spp > p.ns) p.spp = p.ns;
    p.passes = (p.ns + p.spp - 1) / p.spp;
    const int tiles = (n + WQ - 1) / WQ;
    p.cpp = 148 / p.passes;
    if (p.cpp < 1) p.cpp = 1;
    if (p.cpp > tiles) p.cpp = tiles;
    p.smem_w = (size_t)2 * WA_STAGE + (size_t)2 * Cout * WQ * 4 + MAXK * 16 + 8 * 8 + 16 + 1024;
    // (2) dfeat: 64-channel chunks when Cin allows it and the tiles fit, else 32
    p.cpl_f = 0;
    for (int cpl = ((Cin & 63) == 0 ? 2 : 1); cpl >= 1 && !p.cpl_f; --cpl) {
        const int CW = 32 * cpl;
        for (int nx = 2; nx >= 1; --nx) {
            const size_t sm = (size_t)2 * FQ * Cout * 4 + (size_t)2 * 2 * CW * Cout * 4 + (size_t)nx * FQ * (CW + 4) * 4 + MAXK * 16 + 8 * 8 + 16 + 1024;
            if (sm <= 227 * 1024) { p.cpl_f = cpl; p.nx = nx; p.smem_f = sm; break; }
        }
    }
    return p.cpl_f != 0 && p.smem_w <= 227 * 1024;
}
}  // namespace sgb_kpbt

namespace {
__global__ void kpbt_reduce(const float* __restrict__ part, int nparts, long long total, float* __restrict__ gk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * total + i];      // fixed order: deterministic
    gk[i] = s;
}
}  // namespace

/* 1 when sgb_kpconv_bwd_tc takes this shape (otherwise the caller uses sgb_kpconv_bwd) */
extern "C" int sgb_kpconv_bwd_tc_supported(int n, int W, int Cin, int Cout, int K, int n0) {
    sgb_kpbt::Plan p;
    return sgb_kpbt::plan(n, W, Cin, Cout, K, n0, p) ? 1 : 0;
}

extern "C" size_t sgb_kpconv_bwd_tc_ws_bytes(int n, int Cin, int Cout, int K) {
    const size_t kcc = (size_t)(K > 0 ? K : 0) * (size_t)(Cin > 0 ? Cin : 0) * (size_t)(Cout > 0 ? Cout : 0);
    return kcc * 8 + 256 + (size_t)148 * kcc * 4;       // K_values image (hi, lo) + per-CTA partial dK
}

// g [n,Cout] -> gfeat [n0,Cin] (+=, caller zero-fills), gK [K,Cin,Cout] (overwritten)
extern "C" int sgb_kpconv_bwd_tc(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                                 const float* features, const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout,
                                 int K, float KP_extent, int influence, int closest, float* gfeat, float* gK,
                                 void* ws, size_t ws_bytes, void* stream) {
    using namespace sgb_kpbt;
    if (n <= 0 || n0 <= 0 || W < 0 || !(KP_extent > 0.f) || influence < 0 || influence > 2) return SGB_ERR_INVALID;
    if (!g || !query_points || !support_points || (!neighbors && W > 0) || !features || !K_points || !K_values || !gfeat || !gK || !ws) return SGB_ERR_INVALID;
    Plan p;
    if (!plan(n, W, Cin, Cout, K, n0, p)) return SGB_ERR_UNSUPPORTED;
    if (((uintptr_t)features & 15) || ((uintptr_t)g & 15) || ((uintptr_t)gfeat & 15)) return SGB_ERR_UNSUPPORTED;
    if (ws_bytes < sgb_kpconv_bwd_tc_ws_bytes(n, Cin, Cout, K)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t kcc = (size_t)K * Cin * Cout;
    unsigned char* img = (unsigned char*)(((uintptr_t)ws + 127) & ~(uintptr_t)127);
    float* part = (float*)(img + kcc * 8);
    const int tiles_w = sgb_div_up(n, WQ);
    Args a{query_points, support_points, neighbors, features, K_points, g, img, gfeat, part, n, n0, W, Cin, Cout, K,
           KP_extent, influence, closest, tiles_w, p.cpp, p.spp, p.ns, p.nx};
    // ---- (1) dK
#define SGB_KPBW_LAUNCH(CPL, NBL)                                                                                  \
    do {                                                                                                           \
        SGB_OPT_IN_SMEM(kpconv_bwd_w_tc_kernel<CPL, NBL>);                                                         \
        kpconv_bwd_w_tc_kernel<CPL, NBL><<<p.passes * p.cpp, THREADS, p.smem_w, st>>>(a); SGB_COUNT_LAUNCH();     \
    } while (0)
    if (p.cpl_w == 1)      { if (p.nbl == 1) SGB_KPBW_LAUNCH(1, 1); else SGB_KPBW_LAUNCH(1, 2); }
    else if (p.cpl_w == 2) { if (p.nbl == 1) SGB_KPBW_LAUNCH(2, 1); else SGB_KPBW_LAUNCH(2, 2); }
    else                   { if (p.nbl == 1) SGB_KPBW_LAUNCH(4, 1); else SGB_KPBW_LAUNCH(4, 2); }
#undef SGB_KPBW_LAUNCH
    { kpbt_reduce<<<sgb_div_up((long long)kcc, 256), 256, 0, st>>>(part, p.cpp, (long long)kcc, gK); SGB_COUNT_LAUNCH(); }
    // ---- (2) dfeat
    { prep_kernel<<<sgb_div_up((long long)kcc, 256), 256, 0, st>>>(K_values, K, Cin, Cout, 32 * p.cpl_f, img); SGB_COUNT_LAUNCH(); }
    const int grid_f = sgb_div_up(n, FQ);
#define SGB_KPBF_LAUNCH(CPL, NBL)                                                                                  \
    do {                                                                                                           \
        SGB_OPT_IN_SMEM(kpconv_bwd_f_tc_kernel<CPL, NBL>);                                                         \
        kpconv_bwd_f_tc_kernel<CPL, NBL><<<grid_f, THREADS, p.smem_f, st>>>(a); SGB_COUNT_LAUNCH();               \
    } while (0)
    if (p.cpl_f == 2) { if (p.nbl == 1) SGB_KPBF_LAUNCH(2, 1); else SGB_KPBF_LAUNCH(2, 2); }
    else              { if (p.nbl == 1) SGB_KPBF_LAUNCH(1, 1); else SGB_KPBF_LAUNCH(1, 2); }
#undef SGB_KPBF_LAUNCH
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

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
// Warp roles (480 threads, 1 CTA / SM, persistent over a contiguous range of points):
//   warps 0-3   epilogue: tcgen05.ld of their TMEM sub-partition (an M = 64 accumulator keeps rows 16q..16q+15 on lanes
//               32q..32q+15), statistics + max/min, coalesced stores of the per-point candidates;
//   warp  4     TMEM allocation + the single MMA-issuing thread (24 tcgen05.mma per 160-edge tile) + tcgen05.commit;
//   warps 5-14  producers: gather the edge vectors one tile ahead (register prefetch), first layer + BN1 + LeakyReLU on the
//               CUDA cores with packed FFMA2,
//               hi/lo split, 16-byte stores into the canonical no-swizzle K-major tile (conflict free: a warp writes 32
//               consecutive edge rows of one 16-byte chunk), fence.proxy.async, mbarrier arrive.
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

constexpr int TE = 160;                       // edges per tile = 8 points x 20 neighbours (TMEM columns per accumulator)
constexpr int PTS = TE / KNN;                 // 8
constexpr int EPI_WARPS = 4, PROD_WARPS = 10;
constexpr int MMA_WARP = EPI_WARPS;           // warp 4
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;     // 480
constexpr int TILE_BYTES = TE * COUT * 4;     // one H tile (hi or lo): 40 KB
constexpr int W2_BYTES = COUT * COUT * 4;     // 16 KB
constexpr int TMEM_COLS = 512;                // 2 accumulators x 160 columns (power of two)

struct Smem {
    // offsets into the dynamic shared memory block (128-byte aligned base)
    static constexpr int w2_hi = 0;
    static constexpr int w2_lo = w2_hi + W2_BYTES;
    static constexpr int h = w2_lo + W2_BYTES;                     // [stage][hi, lo]
    static constexpr int w1t = h + 4 * TILE_BYTES;                 // [18][64] floats: BN1 scale folded in
    static constexpr int b1 = w1t + CIN * COUT * 4;                // [64]: beta1 - scale1 * mean1
    static constexpr int bars = b1 + COUT * 4;                     // 8 mbarriers
    static constexpr int tmem_slot = bars + 8 * 8;
    static constexpr int total = tmem_slot + 16;
};

template <bool ARG>
__global__ void __launch_bounds__(THREADS, 1)
ec2_tc_kernel(const float* __restrict__ x9, const int* __restrict__ knn, int N, const float* __restrict__ W1,
              const float* __restrict__ stats1, const float* __restrict__ W2,
              float* __restrict__ zmax, float* __restrict__ zmin, unsigned short* __restrict__ kk, double* __restrict__ part /*[grid][128]*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);     // keeps the shared address space (LDS/STS)
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + Smem::bars);       // [2] producers -> MMA
    uint64_t* bar_empty = bar_full + 2;                                      // [2] MMA (commit) -> producers
    uint64_t* bar_tfull = bar_full + 4;                                      // [2] MMA (commit) -> epilogue
    uint64_t* bar_tempty = bar_full + 6;                                     // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Smem::tmem_slot);
    float* s_w1t = reinterpret_cast<float*>(sm + Smem::w1t);
    float* s_b1 = reinterpret_cast<float*>(sm + Smem::b1);
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
        *reinterpret_cast<float*>(sm + Smem::w2_hi + off) = hi;
        *reinterpret_cast<float*>(sm + Smem::w2_lo + off) = tf32_hi(w - hi);
    }
    // first layer with the BatchNorm-1 affine folded in:  v1 = (scale1 W1) e + (beta1 - scale1 mean1)
    for (int i = tid; i < COUT * CIN; i += THREADS) s_w1t[(i % CIN) * COUT + i / CIN] = __ldg(W1 + i) * stats1[128 + i / CIN];
    for (int i = tid; i < COUT; i += THREADS) s_b1[i] = fmaf(-stats1[128 + i], stats1[i], stats1[192 + i]);
    if (tid == 0) {
        mbar_init(&bar_full[0], PROD_WARPS); mbar_init(&bar_full[1], PROD_WARPS);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
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
        // ================= producers: thread = (edge row, half of the hidden channels)
        const int pt = tid - (MMA_WARP + 1) * 32;       // 0..319
        const int er = pt % TE;                         // edge row in the tile
        const int hsel = pt / TE;                       // which 32 hidden channels
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
                const float* xi = x9 + (size_t)(g / KNN) * 9;
                const float* xj = x9 + (size_t)j * 9;
#pragma unroll
                for (int q = 0; q < 9; ++q) { const float a = __ldg(xi + q); en[q] = __ldg(xj + q) - a; en[9 + q] = a; }
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
            unsigned char* dst_hi = sm + Smem::h + (st * 2) * TILE_BYTES;
            unsigned char* dst_lo = dst_hi + TILE_BYTES;
#pragma unroll 2
            for (int c4 = hsel * 8; c4 < hsel * 8 + 8; ++c4) {
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float4 b = *reinterpret_cast<const float4*>(s_b1 + c4 * 4);
                    float2 y01 = make_float2(b.x, b.y), y23 = make_float2(b.z, b.w);
#pragma unroll
                    for (int q = 0; q < CIN; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(s_w1t + q * COUT + c4 * 4);
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
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[st]);
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer
        const uint32_t idesc = make_idesc_tf32(64, TE, false, false);
        const uint32_t a_hi = smem_u32(sm + Smem::w2_hi), a_lo = smem_u32(sm + Smem::w2_lo);
        for (int t = 0; t < ntiles; ++t) {
            const int st = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&bar_full[st], ph);
            mbar_wait(&bar_tempty[st], ph ^ 1u);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t b_hi = smem_u32(sm + Smem::h + (st * 2) * TILE_BYTES), b_lo = b_hi + TILE_BYTES;
                const uint32_t d = tmem + (uint32_t)(st * 256);
#pragma unroll
                for (int i = 0; i < COUT / 8; ++i) {
                    const uint32_t ao = (uint32_t)(2 * i) * (COUT * 16), bo = (uint32_t)(2 * i) * (TE * 16);
                    const uint64_t dah = make_desc(a_hi + ao, COUT * 16, 128), dal = make_desc(a_lo + ao, COUT * 16, 128);
                    const uint64_t dbh = make_desc(b_hi + bo, TE * 16, 128), dbl = make_desc(b_lo + bo, TE * 16, 128);
                    mma_tf32(d, dah, dbh, idesc, i > 0);
                    mma_tf32(d, dal, dbh, idesc, true);
                    mma_tf32(d, dah, dbl, idesc, true);
                }
                mma_commit(&bar_empty[st]);
                mma_commit(&bar_tfull[st]);
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
            const int npts = (int)min((long long)PTS, (g_end - g0) / KNN);
            const long long p0 = g0 / KNN;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * 256);
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

inline int tc_grid(int N) {
    const int tiles = sgb_div_up((long long)N * KNN, TE);
    return tiles < 148 ? (tiles < 1 ? 1 : tiles) : 148;
}
}  // namespace sgb_ectc

// workspace: zmax, zmin [N,64] f32, kk [N,64] u16, partial sums [148][128] f64, reduced sums [128] f64
size_t sgb_ec2_tc_ws_bytes(int N) {
    return (size_t)N * 64 * (4 + 4 + 2) + (size_t)(148 + 1) * 128 * 8 + 256;
}

// second layer of MLP3 on the tensor cores: stats2/var2 and out/argk as sgb_edgeconv_fwd produces them
int sgb_ec2_tc_forward(const float* x9, const int* knn, int N, const float* W1, const float* stats1, const float* W2,
                       const float* gamma2, const float* beta2, float* out, unsigned char* argk, float* stats2, float* var2,
                       void* ws, cudaStream_t st) {
    using namespace sgb_ectc;
    unsigned char* w8 = (unsigned char*)ws;
    double* part = (double*)w8;
    double* sums = part + 148 * 128;
    float* zmax = (float*)(sums + 128);
    float* zmin = zmax + (size_t)N * 64;
    unsigned short* kk = (unsigned short*)(zmin + (size_t)N * 64);
    const int grid = tc_grid(N);
    const size_t smem = Smem::total + 128;
    if (argk) {
        SGB_CUDA(cudaFuncSetAttribute(ec2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { ec2_tc_kernel<true><<<grid, THREADS, smem, st>>>(x9, knn, N, W1, stats1, W2, zmax, zmin, kk, part); SGB_COUNT_LAUNCH(); }
    } else {
        SGB_CUDA(cudaFuncSetAttribute(ec2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { ec2_tc_kernel<false><<<grid, THREADS, smem, st>>>(x9, knn, N, W1, stats1, W2, zmax, zmin, kk, part); SGB_COUNT_LAUNCH(); }
    }
    sgb_bn::reduce_partials(part, grid, 128, sums, st);
    { bn2_from_sums_kernel<<<1, 64, 0, st>>>(sums, (double)N * KNN, gamma2, beta2, stats2, var2); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)N * 64;
    { ec2_apply_kernel<<<sgb_div_up(total / 4, 256), 256, 0, st>>>(zmax, zmin, kk, stats2, total, out, argk); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// a20 on the tensor cores: rigid KPConv forward with the K x Cin x Cout weight contraction on tcgen05.
// replaces kpconv/kernels/convolution_ops.py:161-249 `KPConv_ops` (same semantics as kpconv.cu, which stays the path for
// shapes this kernel does not take: Cin not a multiple of 32, Cout > 256, rows wider than 64 neighbours).
//
//   out[i, :] = sum_k ( sum_j h_k(y_j - x_i) f_j ) . K_values[k]          (convolution_ops.py:240-247)
//
// is a GEMM  out[n, Cout] = WF[n, K*Cin] x K_values[K*Cin, Cout]  whose left operand never exists in HBM.  One CTA owns a
// tile of 128 queries (TMEM lanes) and all Cout <= 256 columns (TMEM columns); the K*Cin reduction is walked kernel point
// by kernel point in chunks of CW = 64 (Cout <= 64) or 32 columns:
//   producer warps 0-15 (8 queries each): the neighbour ids and the neighbour coordinates relative to the query stay in
//       REGISTERS for the whole tile (lane j <-> neighbours j, j + 32); per chunk a lane evaluates the influence of the
//       chunk's kernel point on its neighbours, the warp ballots the non-zero ones and streams only those feature rows
//       (lanes own channels, 4 rows in flight), so with the 'linear' influence a row is gathered about once per tile, not
//       K times.  The weighted sum is split hi/lo (TF32 x 3, tc_common.cuh) and stored as one 128-byte line per row into a
//       K-major SWIZZLE_128B operand tile (conflict-free), fence.proxy.async, mbarrier arrive.
//   warp 16, one thread: streams the matching K_values chunk (pre-split, pre-swizzled image built by a prep kernel) with
//       ONE cp.async.bulk (TMA) per chunk into a 2-stage ring, issues CW/8 x 3 tcgen05.mma (kind::tf32, M = 128, N = Cout)
//       and commits them to the stage's `empty` barrier.
//   epilogue: all 16 producer warps read their TMEM lane quarter with tcgen05.ld and store rows of `out`.
// Algorithmic HBM bytes: 4 n W + 12 (n + n0) + 4 n0 Cin + 4 n Cout + 8 K Cin Cout; gathered rows come from L2.
#include "common.cuh"
#include "tc_common.cuh"

namespace sgb_kptc {
using namespace sgb_tc;

constexpr int PROD_WARPS = 16;
constexpr int THREADS = (PROD_WARPS + 1) * 32;        // 544
constexpr int TQ = 128;                               // queries per tile = TMEM lanes
constexpr int QPW = TQ / PROD_WARPS;                  // 8 queries per producer warp
constexpr int MAXK = 32;
constexpr uint32_t NB_MASK = 0x03ffffffu;             // neighbour id in the low 26 bits, closest kernel point above
enum { INFL_LINEAR = 0, INFL_CONSTANT = 1, INFL_GAUSSIAN = 2 };

struct Args {
    const float* q; const float* s; const int* idx; const float* feat; const float* kpts; const unsigned char* bimg; float* out;
    int n, n0, W, Cin, Cout, K, tmem_cols;
    float extent; int influence; int closest;
};

// K_values [K][Cin][Cout] -> per chunk (kernel point k, channel block cb) the shared-memory image of the B operand:
// rows = output channel, columns = input channel within the block, hi tile followed by lo tile
__global__ void prep_kernel(const float* __restrict__ kval, int K, int Cin, int Cout, int CW, unsigned char* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Cin * Cout) return;
    const int o = i % Cout, cin = (i / Cout) % Cin, k = i / (Cout * Cin);
    const int ncb = Cin / CW, cb = cin / CW, col = cin % CW;
    const size_t tile = (size_t)Cout * CW * 4;
    unsigned char* base = img + (size_t)(k * ncb + cb) * 2 * tile;
    const uint32_t off = sw128_off(o, col, Cout);
    const float v = __ldg(kval + i), hi = tf32_hi(v);
    *reinterpret_cast<float*>(base + off) = hi;
    *reinterpret_cast<float*>(base + tile + off) = tf32_hi(v - hi);
}

template <int CPL, int NBL>
__global__ void __launch_bounds__(THREADS, 1)
kpconv_tc_fwd_kernel(Args a) {
    constexpr int CW = 32 * CPL;
    constexpr int A_TILE = TQ * CW * 4, A_STAGE = 2 * A_TILE;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int B_TILE = a.Cout * CW * 4, B_STAGE = 2 * B_TILE;
    unsigned char* sA = sm;
    unsigned char* sB = sA + 2 * A_STAGE;
    float* s_kp = reinterpret_cast<float*>(sB + 2 * B_STAGE);              // [MAXK][3] (+ pad)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kp + MAXK * 4);
    uint64_t* full_a = bars;            // [2] producers -> MMA, 16 arrivals
    uint64_t* full_b = bars + 2;        // [2] TMA bytes
    uint64_t* empty = bars + 4;         // [2] MMA commit -> producers + TMA issuer
    uint64_t* done = bars + 6;          // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TQ;
    const int ncb = a.Cin / CW, NC = a.K * ncb;

    for (int i = tid; i < a.K * 3; i += THREADS) s_kp[i] = __ldg(a.kpts + i);
    if (tid == 0) {
        mbar_init(&full_a[0], PROD_WARPS); mbar_init(&full_a[1], PROD_WARPS);
        mbar_init(&full_b[0], 1); mbar_init(&full_b[1], 1);
        mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == PROD_WARPS) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == PROD_WARPS) {
        // ------------------------------------------------------------------ TMA + MMA issuer
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_tf32(TQ, a.Cout, false, false);
            for (int c = 0; c < 2 && c < NC; ++c) {
                mbar_expect_tx(&full_b[c], (uint32_t)B_STAGE);
                bulk_g2s(sB + c * B_STAGE, a.bimg + (size_t)c * B_STAGE, (uint32_t)B_STAGE, &full_b[c]);
            }
            for (int c = 0; c < NC; ++c) {
                const int s = c & 1;
                const uint32_t par = (uint32_t)((c >> 1) & 1);
                mbar_wait(&full_b[s], par);
                mbar_wait(&full_a[s], par);
                fence_after_sync();
                const uint32_t a_hi = smem_u32(sA + s * A_STAGE), a_lo = a_hi + A_TILE;
                const uint32_t b_hi = smem_u32(sB + s * B_STAGE), b_lo = b_hi + (uint32_t)B_TILE;
#pragma unroll
                for (int i = 0; i < CW / 8; ++i) {
                    const uint32_t ao = (uint32_t)((i >> 2) * (TQ * 128) + (i & 3) * 32);
                    const uint32_t bo = (uint32_t)((i >> 2) * (a.Cout * 128) + (i & 3) * 32);
                    const uint64_t dah = make_desc_sw(a_hi + ao, 16, 1024, 2), dal = make_desc_sw(a_lo + ao, 16, 1024, 2);
                    const uint64_t dbh = make_desc_sw(b_hi + bo, 16, 1024, 2), dbl = make_desc_sw(b_lo + bo, 16, 1024, 2);
                    mma_tf32(tmem, dah, dbh, idesc, c > 0 || i > 0);
                    mma_tf32(tmem, dal, dbh, idesc, true);
                    mma_tf32(tmem, dah, dbl, idesc, true);
                }
                mma_commit(&empty[s]);
                if (c + 2 < NC) {                       // refill this B stage as soon as its MMAs have retired
                    mbar_wait(&empty[s], par);
                    mbar_expect_tx(&full_b[s], (uint32_t)B_STAGE);
                    bulk_g2s(sB + s * B_STAGE, a.bimg + (size_t)(c + 2) * B_STAGE, (uint32_t)B_STAGE, &full_b[s]);
                }
            }
            mma_commit(done);
        }
    } else {
        // ------------------------------------------------------------------ producers
        int nbp[QPW][NBL];
        float rx[QPW][NBL], ry[QPW][NBL], rz[QPW][NBL];
#pragma unroll
        for (int t = 0; t < QPW; ++t) {
            const int qi = q0 + warp * QPW + t;
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
        const float inv_extent = 1.f / a.extent;
        const float sigma = a.extent * 0.3f;
        const float inv_gauss = 1.f / (2.f * sigma * sigma + 1e-9f);

        for (int c = 0; c < NC; ++c) {
            const int s = c & 1, u = c >> 1;
            const int k = c / ncb, cb = c - k * ncb;
            if (u >= 1) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
            const float kx = s_kp[k * 3], ky = s_kp[k * 3 + 1], kz = s_kp[k * 3 + 2];
            const float* fcol = a.feat + cb * CW + lane * CPL;
            unsigned char* Ahi = sA + s * A_STAGE;
#pragma unroll
            for (int t = 0; t < QPW; ++t) {
                float acc[CPL];
#pragma unroll
                for (int v = 0; v < CPL; ++v) acc[v] = 0.f;
#pragma unroll
                for (int h = 0; h < NBL; ++h) {
                    float w = 0.f;
                    const int me = nbp[t][h];
                    if (me >= 0) {
                        const float dx = rx[t][h] - kx, dy = ry[t][h] - ky, dz = rz[t][h] - kz;
                        const float sq = dx * dx + dy * dy + dz * dz;
                        if (a.influence == INFL_LINEAR) w = fmaxf(1.f - sqrtf(sq) * inv_extent, 0.f);
                        else if (a.influence == INFL_CONSTANT) w = 1.f;
                        else w = expf(-sq * inv_gauss);
                        if (a.closest && (me >> 26) != k) w = 0.f;
                    }
                    unsigned m = __ballot_sync(SGB_FULL_MASK, w != 0.f);
                    while (m) {                               // warp-uniform; 4 feature rows in flight
                        const int l0 = __ffs(m) - 1;
                        float wl[4]; int nl[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int l = m ? __ffs(m) - 1 : l0;
                            const bool live = m != 0;
                            m &= m - 1;
                            wl[g] = __shfl_sync(SGB_FULL_MASK, w, l);
                            nl[g] = __shfl_sync(SGB_FULL_MASK, me, l) & NB_MASK;
                            if (!live) wl[g] = 0.f;
                        }
                        float f[4][CPL];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (CPL == 2) {
                                const float2 v = __ldg(reinterpret_cast<const float2*>(fcol + (size_t)nl[g] * a.Cin));
                                f[g][0] = v.x; f[g][CPL - 1] = v.y;
                            } else {
                                f[g][0] = __ldg(fcol + (size_t)nl[g] * a.Cin);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int v = 0; v < CPL; ++v) acc[v] = fmaf(wl[g], f[g][v], acc[v]);
                    }
                }
                const int r = warp * QPW + t;
                const uint32_t off = sw128_off(r, lane * CPL, TQ);
                if (CPL == 2) {
                    const float h0 = tf32_hi(acc[0]), h1 = tf32_hi(acc[CPL - 1]);
                    *reinterpret_cast<float2*>(Ahi + off) = make_float2(h0, h1);
                    *reinterpret_cast<float2*>(Ahi + A_TILE + off) = make_float2(tf32_hi(acc[0] - h0), tf32_hi(acc[CPL - 1] - h1));
                } else {
                    const float h0 = tf32_hi(acc[0]);
                    *reinterpret_cast<float*>(Ahi + off) = h0;
                    *reinterpret_cast<float*>(Ahi + A_TILE + off) = tf32_hi(acc[0] - h0);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[s]);
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(done, 0);
        fence_after_sync();
        const int quarter = warp & 3, row = quarter * 32 + lane, qi = q0 + row;
        for (int g = warp >> 2; g < (a.Cout >> 4); g += 4) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 16), v);
            if (qi < a.n) {
                float* dst = a.out + (size_t)qi * a.Cout + g * 16;
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    *reinterpret_cast<float4*>(dst + x * 4) = make_float4(v[x * 4], v[x * 4 + 1], v[x * 4 + 2], v[x * 4 + 3]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == PROD_WARPS) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

struct Plan { int cpl, nbl, tmem_cols; size_t smem; };
inline bool plan(int W, int Cin, int Cout, int K, int n0, Plan& p) {
    if (K < 1 || K > MAXK || W < 0 || W > 64 || Cin < 32 || (Cin & 31) || Cout < 16 || (Cout & 15) || Cout > 256) return false;
    if (n0 > (int)NB_MASK) return false;
    p.cpl = ((Cin & 63) == 0 && Cout <= 64) ? 2 : 1;
    p.nbl = W <= 32 ? 1 : 2;
    p.tmem_cols = 32;
    while (p.tmem_cols < Cout) p.tmem_cols <<= 1;
    const int CW = 32 * p.cpl;
    p.smem = (size_t)2 * (2 * TQ * CW * 4) + (size_t)2 * (2 * Cout * CW * 4) + MAXK * 16 + 8 * 8 + 16 + 1024;
    return p.smem <= 227 * 1024;
}
}  // namespace sgb_kptc

/* 1 when sgb_kpconv_fwd_tc takes this shape (otherwise the caller uses sgb_kpconv_fwd) */
extern "C" int sgb_kpconv_tc_supported(int W, int Cin, int Cout, int K, int n0) {
    sgb_kptc::Plan p;
    return sgb_kptc::plan(W, Cin, Cout, K, n0, p) ? 1 : 0;
}

extern "C" size_t sgb_kpconv_tc_ws_bytes(int Cin, int Cout, int K) {
    return (size_t)(K > 0 ? K : 0) * (size_t)(Cin > 0 ? Cin : 0) * (size_t)(Cout > 0 ? Cout : 0) * 8 + 256;
}

extern "C" int sgb_kpconv_fwd_tc(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                                 const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout, int K,
                                 float KP_extent, int influence, int closest, float* out, void* ws, size_t ws_bytes, void* stream) {
    using namespace sgb_kptc;
    if (n < 0 || n0 <= 0 || W < 0 || !(KP_extent > 0.f) || influence < 0 || influence > 2) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!query_points || !support_points || (!neighbors && W > 0) || !features || !K_points || !K_values || !out || !ws) return SGB_ERR_INVALID;
    Plan p;
    if (!plan(W, Cin, Cout, K, n0, p)) return SGB_ERR_UNSUPPORTED;
    if (((uintptr_t)features & 15) || ((uintptr_t)out & 15)) return SGB_ERR_UNSUPPORTED;
    if (ws_bytes < sgb_kpconv_tc_ws_bytes(Cin, Cout, K)) return SGB_ERR_WORKSPACE;
    unsigned char* img = (unsigned char*)(((uintptr_t)ws + 127) & ~(uintptr_t)127);
    cudaStream_t st = (cudaStream_t)stream;
    const int total = K * Cin * Cout;
    { prep_kernel<<<sgb_div_up(total, 256), 256, 0, st>>>(K_values, K, Cin, Cout, 32 * p.cpl, img); SGB_COUNT_LAUNCH(); }
    Args a{query_points, support_points, neighbors, features, K_points, img, out, n, n0, W, Cin, Cout, K, p.tmem_cols,
           KP_extent, influence, closest};
    const int grid = sgb_div_up(n, TQ);
#define SGB_KPTC_LAUNCH(CPL, NBL)                                                                     \
    do {                                                                                              \
        SGB_OPT_IN_SMEM(kpconv_tc_fwd_kernel<CPL, NBL>);                                              \
        kpconv_tc_fwd_kernel<CPL, NBL><<<grid, THREADS, p.smem, st>>>(a); SGB_COUNT_LAUNCH();         \
    } while (0)
    if (p.cpl == 2) { if (p.nbl == 1) SGB_KPTC_LAUNCH(2, 1); else SGB_KPTC_LAUNCH(2, 2); }
    else            { if (p.nbl == 1) SGB_KPTC_LAUNCH(1, 1); else SGB_KPTC_LAUNCH(1, 2); }
#undef SGB_KPTC_LAUNCH
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

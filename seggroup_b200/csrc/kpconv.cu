// a20: fused rigid KPConv (kernel-point convolution), fp32 SIMT path.
// replaces kpconv/kernels/convolution_ops.py:161-249 `KPConv_ops` (a chain of tf.gather / tf.matmul that
// materialises [n, W, K, 3] differences, [n, K, W] weights and [n, K, Cin] weighted features in HBM).
//
// One CTA owns a tile of T queries and never writes an intermediate to HBM:
//   phase 1 (gather + influence): per query, lanes first evaluate the K kernel-point influences of 32 neighbours
//            at a time (one neighbour per lane: centre on the query, distance to every kernel point, influence
//            function, optional closest-kernel-point mask), park them in shared memory, then all lanes stream the
//            neighbours' feature rows (one coalesced row per neighbour, lanes own channels) and accumulate
//            wf[k][c] += w_k * f[c] in registers, skipping zero influences (warp-uniform branch);
//            the [T, K*Cin] tile of weighted features lands in shared memory;
//   phase 2 (contraction): out[T, Cout] = wf[T, K*Cin] x K_values[K*Cin, Cout], register-tiled, K_values streamed
//            through shared memory in row chunks (read once per CTA, L2-resident across CTAs).
// Shadow neighbours (index >= n_support; the reference appends a point at 1e6 and a zero feature row) contribute
// exactly zero in every mode and are skipped.
// Algorithmic HBM bytes: 4 n W (indices) + 12 (n + n0) + 4 n0 Cin + 4 n Cout + 4 K Cin Cout.
#include "common.cuh"

namespace {
constexpr int KP_THREADS = 256;
constexpr int KP_WARPS = 8;
constexpr int KP_MAXK = 32;          // kernel points
enum { INFL_LINEAR = 0, INFL_CONSTANT = 1, INFL_GAUSSIAN = 2 };

struct KpArgs {
    const float* q; const float* s; const int* idx; const float* feat; const float* kpts; const float* kval; float* out;
    int n, n0, W, Cin, Cout, K, T, rows_b;
    float extent; int influence; int closest;
};

// K influences of one neighbour (convolution_ops.py:194-229) into w[0..K)
__device__ __forceinline__ void lane_influences(const KpArgs& a, int nb, float qx, float qy, float qz, const float* s_kp, float* w,
                                                float inv_extent, float inv_gauss) {
    const float rx = __ldg(a.s + (size_t)nb * 3) - qx, ry = __ldg(a.s + (size_t)nb * 3 + 1) - qy, rz = __ldg(a.s + (size_t)nb * 3 + 2) - qz;
    float best = INFINITY; int bk = 0;
    for (int k = 0; k < a.K; ++k) {
        const float dx = rx - s_kp[k * 3], dy = ry - s_kp[k * 3 + 1], dz = rz - s_kp[k * 3 + 2];
        const float sq = dx * dx + dy * dy + dz * dz;
        float v;
        if (a.influence == INFL_LINEAR) v = fmaxf(1.f - sqrtf(sq) * inv_extent, 0.f);
        else if (a.influence == INFL_CONSTANT) v = 1.f;
        else v = expf(-sq * inv_gauss);
        w[k] = v;
        if (sq < best) { best = sq; bk = k; }
    }
    if (a.closest) for (int k = 0; k < a.K; ++k) if (k != bk) w[k] = 0.f;
}

// CPL = channels per lane handled in one pass over the neighbours (Cin is covered in ceil(Cin / (32*CPL)) passes)
template <int CPL>
__device__ __forceinline__ void weighted_features_tile(const KpArgs& a, int q0, float* s_A, float* my_w, const float* s_kp) {
    const int KC = a.K * a.Cin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_extent = 1.f / a.extent;
    const float sigma = a.extent * 0.3f;
    const float inv_gauss = 1.f / (2.f * sigma * sigma + 1e-9f);
    // ---------------- phase 1
    for (int t = warp; t < a.T; t += KP_WARPS) {
        const int qi = q0 + t;
        float* Arow = s_A + (size_t)t * KC;
        if (qi >= a.n) { for (int i = lane; i < KC; i += 32) Arow[i] = 0.f; continue; }
        const float qx = __ldg(a.q + (size_t)qi * 3), qy = __ldg(a.q + (size_t)qi * 3 + 1), qz = __ldg(a.q + (size_t)qi * 3 + 2);
        for (int cbase = 0; cbase < a.Cin; cbase += 32 * CPL) {
            float wf[17][CPL];
#pragma unroll
            for (int k = 0; k < 17; ++k)
#pragma unroll
                for (int c = 0; c < CPL; ++c) wf[k][c] = 0.f;
            for (int j0 = 0; j0 < a.W; j0 += 32) {
                // lanes: one neighbour each -> K influences into shared memory
                const int j = j0 + lane;
                int nb = a.n0;
                if (j < a.W) nb = __ldg(a.idx + (size_t)qi * a.W + j);
                const bool valid = nb >= 0 && nb < a.n0;
                __syncwarp();
                if (valid) lane_influences(a, nb, qx, qy, qz, s_kp, my_w + lane * KP_MAXK, inv_extent, inv_gauss);
                const unsigned vmask = __ballot_sync(SGB_FULL_MASK, valid);
                __syncwarp();
                // all lanes: stream the valid neighbours' feature rows
                unsigned m = vmask;
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int nbl = __shfl_sync(SGB_FULL_MASK, nb, l);
                    float f[CPL];
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const int ch = cbase + c * 32 + lane;
                        f[c] = ch < a.Cin ? __ldg(a.feat + (size_t)nbl * a.Cin + ch) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 17; ++k) {
                        if (k < a.K) {
                            const float w = my_w[l * KP_MAXK + k];
                            if (w != 0.f) {
#pragma unroll
                                for (int c = 0; c < CPL; ++c) wf[k][c] = fmaf(w, f[c], wf[k][c]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                if (k < a.K) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const int ch = cbase + c * 32 + lane;
                        if (ch < a.Cin) Arow[k * a.Cin + ch] = wf[k][c];
                    }
                }
            }
        }
    }
}

template <int CPL>
__global__ void __launch_bounds__(KP_THREADS)
kpconv_fwd_kernel(KpArgs a) {
    extern __shared__ __align__(16) float kp_smem[];
    const int KC = a.K * a.Cin;
    float* s_A = kp_smem;                                   // [T][KC]
    float* s_B = s_A + (((size_t)a.T * KC + 3) & ~(size_t)3);   // [rows_b][Cout], 16 B aligned
    float* s_w = s_B + (size_t)a.rows_b * a.Cout;           // [KP_WARPS][32][KP_MAXK]
    float* s_kp = s_w + KP_WARPS * 32 * KP_MAXK;            // [KP_MAXK][3]
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.K * 3; i += KP_THREADS) s_kp[i] = __ldg(a.kpts + i);
    __syncthreads();
    const int q0 = blockIdx.x * a.T;
    weighted_features_tile<CPL>(a, q0, s_A, s_w + warp * 32 * KP_MAXK, s_kp);
    __syncthreads();

    // ---------------- phase 2: out[T][Cout] = A[T][KC] x B[KC][Cout]
    const int cols4 = a.Cout >> 2;
    const int nrg = KP_THREADS / cols4;                     // row groups
    const int cg = threadIdx.x % cols4, rg = threadIdx.x / cols4;
    const bool worker = rg < nrg;
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
    for (int r0 = 0; r0 < KC; r0 += a.rows_b) {
        const int nr = min(a.rows_b, KC - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * cols4; i += KP_THREADS)
            reinterpret_cast<float4*>(s_B)[i] = __ldg(reinterpret_cast<const float4*>(a.kval + (size_t)r0 * a.Cout) + i);
        __syncthreads();
        if (worker) {
            for (int r = 0; r < nr; ++r) {
                const float4 b = reinterpret_cast<const float4*>(s_B)[r * cols4 + cg];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int t = rg + i * nrg;
                    if (t < a.T) {
                        const float av = s_A[(size_t)t * KC + r0 + r];
                        acc[i][0] = fmaf(av, b.x, acc[i][0]); acc[i][1] = fmaf(av, b.y, acc[i][1]);
                        acc[i][2] = fmaf(av, b.z, acc[i][2]); acc[i][3] = fmaf(av, b.w, acc[i][3]);
                    }
                }
            }
        }
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = rg + i * nrg;
            if (t < a.T && q0 + t < a.n)
                *reinterpret_cast<float4*>(a.out + (size_t)(q0 + t) * a.Cout + cg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward (gradients w.r.t. features and K_values; kernel points and coordinates are not trained in rigid KPConv)
//   dK[kc][o]   = sum_i wf_i[kc] * g_i[o]
//   dfeat[j][c] += sum_k w_ik(j) * (g_i . K_values[k][c][:])        for every neighbour j of query i
// Persistent CTAs: each owns a private [K*Cin, Cout] accumulator in global scratch (L2-resident) that it updates
// tile after tile without atomics; `kpconv_bwd_reduce` then sums the per-CTA accumulators in a fixed order, so dK is
// deterministic.  dfeat is a scatter through the neighbour lists: warp-coalesced red.global.add.f32 (summation order
// across queries is not fixed, as with any scatter-add; documented in DESIGN.md).
// ------------------------------------------------------------------------------------------------
struct KpBwdArgs { const float* g; float* gfeat; float* dk_part; int n_tiles; };

template <int CPL>
__global__ void __launch_bounds__(KP_THREADS)
kpconv_bwd_kernel(KpArgs a, KpBwdArgs b) {
    extern __shared__ __align__(16) float kp_smem[];
    const int KC = a.K * a.Cin;
    float* s_A = kp_smem;                                        // [T][KC]   wf, later gw
    float* s_B = s_A + (((size_t)a.T * KC + 3) & ~(size_t)3);    // [rows_b][Cout]
    float* s_w = s_B + (size_t)a.rows_b * a.Cout;                // [KP_WARPS][32][KP_MAXK]
    float* s_kp = s_w + KP_WARPS * 32 * KP_MAXK;                 // [KP_MAXK][3]
    float* s_G = s_kp + KP_MAXK * 3 + 1;                         // [T][Cout + 4] (padded rows: fewer bank conflicts)
    const int GS = a.Cout + 4;
    s_G = (float*)(((uintptr_t)s_G + 15) & ~(uintptr_t)15);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.K * 3; i += KP_THREADS) s_kp[i] = __ldg(a.kpts + i);
    float* my_w = s_w + warp * 32 * KP_MAXK;
    float* dk = b.dk_part + (size_t)blockIdx.x * KC * a.Cout;
    const int cols4 = a.Cout >> 2;
    const float inv_extent = 1.f / a.extent;
    const float sigma = a.extent * 0.3f;
    const float inv_gauss = 1.f / (2.f * sigma * sigma + 1e-9f);
    bool first = true;
    for (int tile = blockIdx.x; tile < b.n_tiles; tile += gridDim.x) {
        const int q0 = tile * a.T;
        __syncthreads();
        for (int i = threadIdx.x; i < a.T * a.Cout; i += KP_THREADS) {
            const int t = i / a.Cout;
            s_G[t * GS + (i % a.Cout)] = (q0 + t < a.n) ? __ldg(b.g + (size_t)(q0 + t) * a.Cout + (i % a.Cout)) : 0.f;
        }
        weighted_features_tile<CPL>(a, q0, s_A, my_w, s_kp);
        __syncthreads();
        // ---- dK accumulator: dk[r][o] (+)= sum_t A[t][r] G[t][o]; thread owns float4 columns of consecutive rows
        for (int e = threadIdx.x; e < KC * cols4; e += KP_THREADS) {
            const int r = e / cols4, cg = e % cols4;
            float4 acc = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dk + (size_t)r * a.Cout + cg * 4);
            for (int t = 0; t < a.T; ++t) {
                const float av = s_A[(size_t)t * KC + r];
                const float4 gv = *reinterpret_cast<const float4*>(s_G + t * GS + cg * 4);
                acc.x = fmaf(av, gv.x, acc.x); acc.y = fmaf(av, gv.y, acc.y); acc.z = fmaf(av, gv.z, acc.z); acc.w = fmaf(av, gv.w, acc.w);
            }
            *reinterpret_cast<float4*>(dk + (size_t)r * a.Cout + cg * 4) = acc;
        }
        first = false;
        __syncthreads();
        // ---- gw[t][r] = sum_o G[t][o] K_values[r][o]  (overwrites A), K_values streamed through s_B
        for (int r0 = 0; r0 < KC; r0 += a.rows_b) {
            const int nr = min(a.rows_b, KC - r0);
            __syncthreads();
            for (int i = threadIdx.x; i < nr * cols4; i += KP_THREADS)
                reinterpret_cast<float4*>(s_B)[i] = __ldg(reinterpret_cast<const float4*>(a.kval + (size_t)r0 * a.Cout) + i);
            __syncthreads();
            for (int e = threadIdx.x; e < nr * a.T; e += KP_THREADS) {
                const int t = e % a.T, r = e / a.T;
                float acc = 0.f;
                for (int c4 = 0; c4 < cols4; ++c4) {
                    const float4 bv = reinterpret_cast<const float4*>(s_B)[r * cols4 + c4];
                    const float4 gv = *reinterpret_cast<const float4*>(s_G + t * GS + c4 * 4);
                    acc = fmaf(bv.x, gv.x, acc); acc = fmaf(bv.y, gv.y, acc); acc = fmaf(bv.z, gv.z, acc); acc = fmaf(bv.w, gv.w, acc);
                }
                s_A[(size_t)t * KC + r0 + r] = acc;
            }
        }
        __syncthreads();
        // ---- scatter to the neighbours' feature gradients
        for (int t = warp; t < a.T; t += KP_WARPS) {
            const int qi = q0 + t;
            if (qi >= a.n) continue;
            const float* gw = s_A + (size_t)t * KC;
            const float qx = __ldg(a.q + (size_t)qi * 3), qy = __ldg(a.q + (size_t)qi * 3 + 1), qz = __ldg(a.q + (size_t)qi * 3 + 2);
            for (int j0 = 0; j0 < a.W; j0 += 32) {
                const int j = j0 + lane;
                int nb = a.n0;
                if (j < a.W) nb = __ldg(a.idx + (size_t)qi * a.W + j);
                const bool valid = nb >= 0 && nb < a.n0;
                __syncwarp();
                if (valid) lane_influences(a, nb, qx, qy, qz, s_kp, my_w + lane * KP_MAXK, inv_extent, inv_gauss);
                unsigned m = __ballot_sync(SGB_FULL_MASK, valid);
                __syncwarp();
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int nbl = __shfl_sync(SGB_FULL_MASK, nb, l);
                    for (int ch = lane; ch < a.Cin; ch += 32) {
                        float acc = 0.f;
                        for (int k = 0; k < a.K; ++k) {
                            const float w = my_w[l * KP_MAXK + k];
                            if (w != 0.f) acc = fmaf(w, gw[k * a.Cin + ch], acc);
                        }
                        if (acc != 0.f) atomicAdd(b.gfeat + (size_t)nbl * a.Cin + ch, acc);
                    }
                }
            }
        }
    }
    if (first) {                        // this CTA had no tile: its accumulator must still read as zero
        for (int e = threadIdx.x; e < KC * a.Cout; e += KP_THREADS) dk[e] = 0.f;
    }
}

__global__ void kpconv_bwd_reduce(const float* __restrict__ part, int nparts, long long total, float* __restrict__ gk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * total + i];
    gk[i] = s;
}

struct KpPlan { int T, rows_b, cpl; size_t smem; };
inline bool kp_plan(int K, int Cin, int Cout, KpPlan& p) {
    if (K < 1 || K > 17 || Cin < 1 || Cout < 4 || (Cout & 3) || Cout > 1024) return false;
    const int cols4 = Cout >> 2;
    if (cols4 > KP_THREADS) return false;
    const int nrg = KP_THREADS / cols4;
    p.rows_b = Cout <= 128 ? 32 : 8;
    p.cpl = Cin <= 32 ? 1 : (Cin <= 64 ? 2 : 4);
    const size_t fixed = ((size_t)p.rows_b * Cout + KP_WARPS * 32 * KP_MAXK + KP_MAXK * 3) * sizeof(float);
    for (int T = 32; T >= 1; T >>= 1) {
        if (T > 8 * nrg) continue;                       // at most 8 rows per thread in phase 2
        const size_t sm = fixed + ((((size_t)T * K * Cin + 3) & ~(size_t)3)) * sizeof(float);
        if (sm <= 200 * 1024) { p.T = T; p.smem = sm; return true; }
    }
    return false;
}
}  // namespace

// influence: 0 linear, 1 constant, 2 gaussian; closest: 0 = 'sum', 1 = 'closest' aggregation (convolution_ops.py:208-231)
extern "C" int sgb_kpconv_fwd(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                              const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout, int K,
                              float KP_extent, int influence, int closest, float* out, void* stream) {
    if (n < 0 || n0 <= 0 || W < 0 || !(KP_extent > 0.f) || influence < 0 || influence > 2) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!query_points || !support_points || (!neighbors && W > 0) || !features || !K_points || !K_values || !out) return SGB_ERR_INVALID;
    KpPlan p;
    if (!kp_plan(K, Cin, Cout, p)) return SGB_ERR_UNSUPPORTED;
    KpArgs a{query_points, support_points, neighbors, features, K_points, K_values, out, n, n0, W, Cin, Cout, K, p.T, p.rows_b,
             KP_extent, influence, closest};
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sgb_div_up(n, p.T);
    if (p.cpl == 1) {
        SGB_OPT_IN_SMEM(kpconv_fwd_kernel<1>);
        { kpconv_fwd_kernel<1><<<grid, KP_THREADS, p.smem, st>>>(a); SGB_COUNT_LAUNCH(); }
    } else if (p.cpl == 2) {
        SGB_OPT_IN_SMEM(kpconv_fwd_kernel<2>);
        { kpconv_fwd_kernel<2><<<grid, KP_THREADS, p.smem, st>>>(a); SGB_COUNT_LAUNCH(); }
    } else {
        SGB_OPT_IN_SMEM(kpconv_fwd_kernel<4>);
        { kpconv_fwd_kernel<4><<<grid, KP_THREADS, p.smem, st>>>(a); SGB_COUNT_LAUNCH(); }
    }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

inline int kp_bwd_grid(int n, int T) {
    const int tiles = sgb_div_up(n, T);
    return tiles < 148 ? tiles : 148;
}

extern "C" size_t sgb_kpconv_bwd_ws_bytes(int n, int Cin, int Cout, int K) {
    return (size_t)148 * K * Cin * Cout * sizeof(float) + 256;
}

// g [n,Cout] -> gfeat [n0,Cin] (+=, caller zero-fills), gK [K,Cin,Cout] (overwritten)
extern "C" int sgb_kpconv_bwd(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                              const float* features, const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout,
                              int K, float KP_extent, int influence, int closest, float* gfeat, float* gK,
                              void* ws, size_t ws_bytes, void* stream) {
    if (n <= 0 || n0 <= 0 || W < 0 || !(KP_extent > 0.f) || influence < 0 || influence > 2) return SGB_ERR_INVALID;
    if (!g || !query_points || !support_points || (!neighbors && W > 0) || !features || !K_points || !K_values || !gfeat || !gK || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_kpconv_bwd_ws_bytes(n, Cin, Cout, K)) return SGB_ERR_WORKSPACE;
    KpPlan p;
    if (!kp_plan(K, Cin, Cout, p)) return SGB_ERR_UNSUPPORTED;
    // extra shared memory for the gradient tile; shrink the tile if needed
    size_t smem = p.smem + ((size_t)p.T * (Cout + 4) + 8) * sizeof(float);
    while (smem > 200 * 1024 && p.T > 1) {
        p.T >>= 1;
        smem = ((size_t)p.rows_b * Cout + KP_WARPS * 32 * KP_MAXK + KP_MAXK * 3 + ((((size_t)p.T * K * Cin + 3) & ~(size_t)3)) + (size_t)p.T * (Cout + 4) + 8) * sizeof(float);
    }
    if (smem > 200 * 1024) return SGB_ERR_UNSUPPORTED;
    KpArgs a{query_points, support_points, neighbors, features, K_points, K_values, nullptr, n, n0, W, Cin, Cout, K, p.T, p.rows_b,
             KP_extent, influence, closest};
    const int tiles = sgb_div_up(n, p.T);
    const int grid = kp_bwd_grid(n, p.T);
    KpBwdArgs b{g, gfeat, (float*)ws, tiles};
    cudaStream_t st = (cudaStream_t)stream;
    if (p.cpl == 1) {
        SGB_OPT_IN_SMEM(kpconv_bwd_kernel<1>);
        { kpconv_bwd_kernel<1><<<grid, KP_THREADS, smem, st>>>(a, b); SGB_COUNT_LAUNCH(); }
    } else if (p.cpl == 2) {
        SGB_OPT_IN_SMEM(kpconv_bwd_kernel<2>);
        { kpconv_bwd_kernel<2><<<grid, KP_THREADS, smem, st>>>(a, b); SGB_COUNT_LAUNCH(); }
    } else {
        SGB_OPT_IN_SMEM(kpconv_bwd_kernel<4>);
        { kpconv_bwd_kernel<4><<<grid, KP_THREADS, smem, st>>>(a, b); SGB_COUNT_LAUNCH(); }
    }
    const long long total = (long long)K * Cin * Cout;
    { kpconv_bwd_reduce<<<sgb_div_up(total, 256), 256, 0, st>>>((const float*)ws, grid, total, gK); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

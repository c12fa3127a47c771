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
    // deformable KPConv (convolution_ops.py:371-493): per-query kernel-point offsets [n,K,3] (NULL = rigid) and modulations [n,K] (or NULL)
    const float* offsets; const float* mods;
};
constexpr int KP_DSTRIDE = KP_MAXK * 4;     // per-warp staging of one query's deformed kernel points [K][3] and modulations [K]

// K influences of one neighbour (convolution_ops.py:194-229) into w[0..K).  Deformable variant (a.offsets != NULL, :415-471): the kernel
// points are the query's deformed ones, 'constant' influence is the indicator sq < extent^2, a neighbour that is within KP_extent of
// no kernel point is dropped altogether (the reference compacts such neighbours away, :426-445), and the influences are scaled by the
// query's modulations (:485-486: equivalent to scaling the weighted features per kernel point).
__device__ __forceinline__ void lane_influences(const KpArgs& a, int nb, float qx, float qy, float qz, const float* s_kp, float* w,
                                                float inv_extent, float inv_gauss, const float* s_mod = nullptr) {
    const float rx = __ldg(a.s + (size_t)nb * 3) - qx, ry = __ldg(a.s + (size_t)nb * 3 + 1) - qy, rz = __ldg(a.s + (size_t)nb * 3 + 2) - qz;
    const bool deform = a.offsets != nullptr;
    const float ext2 = a.extent * a.extent;
    float best = INFINITY; int bk = 0;
    bool in_range = false;
    for (int k = 0; k < a.K; ++k) {
        const float dx = rx - s_kp[k * 3], dy = ry - s_kp[k * 3 + 1], dz = rz - s_kp[k * 3 + 2];
        const float sq = dx * dx + dy * dy + dz * dz;
        float v;
        if (a.influence == INFL_LINEAR) v = fmaxf(1.f - sqrtf(sq) * inv_extent, 0.f);
        else if (a.influence == INFL_CONSTANT) v = deform ? (sq < ext2 ? 1.f : 0.f) : 1.f;
        else v = expf(-sq * inv_gauss);
        w[k] = v;
        in_range |= sq < ext2;
        if (sq < best) { best = sq; bk = k; }
    }
    if (a.closest) for (int k = 0; k < a.K; ++k) if (k != bk) w[k] = 0.f;
    if (deform) {
        if (!in_range) for (int k = 0; k < a.K; ++k) w[k] = 0.f;
        else if (s_mod) for (int k = 0; k < a.K; ++k) w[k] *= s_mod[k];
    }
}
// stage the kernel points of query qi for the calling warp: rigid -> the shared table, deformable -> K_points + offsets[qi] (+ modulations)
__device__ __forceinline__ const float* stage_query_kp(const KpArgs& a, int qi, const float* s_kp, float* s_d, const float*& s_mod) {
    s_mod = nullptr;
    if (!a.offsets) return s_kp;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    for (int i = lane; i < a.K * 3; i += 32) s_d[i] = s_kp[i] + __ldg(a.offsets + (size_t)qi * a.K * 3 + i);
    if (a.mods) {
        for (int i = lane; i < a.K; i += 32) s_d[KP_MAXK * 3 + i] = __ldg(a.mods + (size_t)qi * a.K + i);
        s_mod = s_d + KP_MAXK * 3;
    }
    __syncwarp();
    return s_d;
}

// CPL = channels per lane handled in one pass over the neighbours (Cin is covered in ceil(Cin / (32*CPL)) passes)
template <int CPL>
__device__ __forceinline__ void weighted_features_tile(const KpArgs& a, int q0, float* s_A, float* my_w, const float* s_kp, float* s_d) {
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
        const float* s_mod;
        const float* kp = stage_query_kp(a, qi, s_kp, s_d, s_mod);
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
                if (valid) lane_influences(a, nb, qx, qy, qz, kp, my_w + lane * KP_MAXK, inv_extent, inv_gauss, s_mod);
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
    __shared__ float s_dq[KP_WARPS * KP_DSTRIDE];           // deformed kernel points / modulations of the query a warp works on
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.K * 3; i += KP_THREADS) s_kp[i] = __ldg(a.kpts + i);
    __syncthreads();
    const int q0 = blockIdx.x * a.T;
    weighted_features_tile<CPL>(a, q0, s_A, s_w + warp * 32 * KP_MAXK, s_kp, s_dq + warp * KP_DSTRIDE);
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
struct KpBwdArgs { const float* g; float* gfeat; float* dk_part; int n_tiles; float* goff /*[n,K,3]*/; float* gmod /*[n,K]*/; };

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
    __shared__ float s_dq[KP_WARPS * KP_DSTRIDE];
    __shared__ float s_dmod[32 * KP_MAXK];                       // [T][K]: sum_c wf_mod * gw per (query, kernel point)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.K * 3; i += KP_THREADS) s_kp[i] = __ldg(a.kpts + i);
    float* my_w = s_w + warp * 32 * KP_MAXK;
    float* my_dq = s_dq + warp * KP_DSTRIDE;
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
        if (b.gmod) for (int i = threadIdx.x; i < 32 * KP_MAXK; i += KP_THREADS) s_dmod[i] = 0.f;
        weighted_features_tile<CPL>(a, q0, s_A, my_w, s_kp, my_dq);
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
                if (b.gmod) atomicAdd(&s_dmod[t * KP_MAXK + (r0 + r) / a.Cin], s_A[(size_t)t * KC + r0 + r] * acc);   // wf_mod * gw
                s_A[(size_t)t * KC + r0 + r] = acc;
            }
        }
        __syncthreads();
        // d out / d modulation[k] = sum_c wf[k][c] gw[k][c] with the UNmodulated wf = wf_mod / mod  (mod = 2 sigmoid(.) > 0)
        if (b.gmod) {
            for (int i = threadIdx.x; i < a.T * a.K; i += KP_THREADS) {
                const int t = i / a.K, k = i % a.K;
                if (q0 + t < a.n) {
                    const float m = __ldg(a.mods + (size_t)(q0 + t) * a.K + k);
                    b.gmod[(size_t)(q0 + t) * a.K + k] = m != 0.f ? s_dmod[t * KP_MAXK + k] / m : 0.f;
                }
            }
        }
        // ---- scatter to the neighbours' feature gradients (+ the offset gradients of the deformable operator)
        for (int t = warp; t < a.T; t += KP_WARPS) {
            const int qi = q0 + t;
            if (qi >= a.n) continue;
            const float* gw = s_A + (size_t)t * KC;
            const float qx = __ldg(a.q + (size_t)qi * 3), qy = __ldg(a.q + (size_t)qi * 3 + 1), qz = __ldg(a.q + (size_t)qi * 3 + 2);
            const float* s_mod;
            const float* kp = stage_query_kp(a, qi, s_kp, my_dq, s_mod);
            float go0 = 0.f, go1 = 0.f, go2 = 0.f;                   // lane k: d out / d offset[qi][k][:]
            for (int j0 = 0; j0 < a.W; j0 += 32) {
                const int j = j0 + lane;
                int nb = a.n0;
                if (j < a.W) nb = __ldg(a.idx + (size_t)qi * a.W + j);
                const bool valid = nb >= 0 && nb < a.n0;
                __syncwarp();
                if (valid) lane_influences(a, nb, qx, qy, qz, kp, my_w + lane * KP_MAXK, inv_extent, inv_gauss, s_mod);
                unsigned m = __ballot_sync(SGB_FULL_MASK, valid);
                __syncwarp();
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int nbl = __shfl_sync(SGB_FULL_MASK, nb, l);
                    if (b.goff && a.influence != INFL_CONSTANT) {
                        // d out / d w'_k(j) = sum_c f_j[c] gw[k][c] (w' = modulated influence); chain through the influence function
                        float mine = 0.f;
                        for (int k = 0; k < a.K; ++k) {
                            if (my_w[l * KP_MAXK + k] == 0.f) continue;          // clamped / masked / dropped: no gradient (warp uniform)
                            float part = 0.f;
                            for (int ch = lane; ch < a.Cin; ch += 32) part = fmaf(__ldg(a.feat + (size_t)nbl * a.Cin + ch), gw[k * a.Cin + ch], part);
                            part = sgb_warp_sum(part);
                            if (lane == k) mine = part;
                        }
                        if (lane < a.K && my_w[l * KP_MAXK + lane] != 0.f) {
                            const float rx = __ldg(a.s + (size_t)nbl * 3) - qx, ry = __ldg(a.s + (size_t)nbl * 3 + 1) - qy, rz = __ldg(a.s + (size_t)nbl * 3 + 2) - qz;
                            const float dx = rx - kp[lane * 3], dy = ry - kp[lane * 3 + 1], dz = rz - kp[lane * 3 + 2];
                            const float sq = dx * dx + dy * dy + dz * dz;
                            const float md = s_mod ? s_mod[lane] : 1.f;
                            float coef;                                            // d w'_k / d kp_k = coef * (r - kp_k)
                            if (a.influence == INFL_LINEAR) coef = sq > 0.f ? md * inv_extent * rsqrtf(sq) : 0.f;      // w = 1 - d / extent
                            else coef = my_w[l * KP_MAXK + lane] * 2.f * inv_gauss;                                    // w' = md exp(-sq * inv_gauss)
                            go0 = fmaf(mine * coef, dx, go0); go1 = fmaf(mine * coef, dy, go1); go2 = fmaf(mine * coef, dz, go2);
                        }
                    }
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
            if (b.goff && lane < a.K) {
                float* o = b.goff + ((size_t)qi * a.K + lane) * 3;
                o[0] = go0; o[1] = go1; o[2] = go2;
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

// 'fitting' regulariser of the deformable layers (kpconv/models/KPFCNN_model.py:242-266): for every (query, kernel point) the neighbour
// closest to the DEFORMED kernel point (first minimum over the neighbour row; shadow neighbours count as a point at 1000).  One warp
// per query, lanes over neighbours, K sequential.
__global__ void deform_closest_kernel(const float* __restrict__ q, const float* __restrict__ s, const int* __restrict__ idx,
                                      const float* __restrict__ kpts, const float* __restrict__ offsets, int n, int n0, int W, int K,
                                      int* __restrict__ arg) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= n) return;
    const float qx = __ldg(q + (size_t)qi * 3), qy = __ldg(q + (size_t)qi * 3 + 1), qz = __ldg(q + (size_t)qi * 3 + 2);
    for (int k = 0; k < K; ++k) {
        const float kx = __ldg(kpts + k * 3) + __ldg(offsets + ((size_t)qi * K + k) * 3);
        const float ky = __ldg(kpts + k * 3 + 1) + __ldg(offsets + ((size_t)qi * K + k) * 3 + 1);
        const float kz = __ldg(kpts + k * 3 + 2) + __ldg(offsets + ((size_t)qi * K + k) * 3 + 2);
        float best = INFINITY; int bj = 0x7fffffff;
        for (int j = lane; j < W; j += 32) {
            int nb = __ldg(idx + (size_t)qi * W + j);
            const bool valid = nb >= 0 && nb < n0;
            const float px = valid ? __ldg(s + (size_t)nb * 3) : 1000.f, py = valid ? __ldg(s + (size_t)nb * 3 + 1) : 1000.f,
                        pz = valid ? __ldg(s + (size_t)nb * 3 + 2) : 1000.f;
            const float dx = (px - qx) - kx, dy = (py - qy) - ky, dz = (pz - qz) - kz;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) { best = d2; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(SGB_FULL_MASK, best, o);
            const int oj = __shfl_xor_sync(SGB_FULL_MASK, bj, o);
            if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0) arg[(size_t)qi * K + k] = bj == 0x7fffffff ? 0 : bj;       // column of the neighbour row
    }
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
    return sgb_kpconv_deform_fwd(query_points, support_points, neighbors, features, K_points, nullptr, nullptr, K_values, n, n0, W, Cin, Cout, K,
                                 KP_extent, influence, closest, out, stream);
}
// offsets [n,K,3] (NULL = rigid KPConv), modulations [n,K] (or NULL): convolution_ops.py:371-493 `KPConv_deform_ops`
extern "C" int sgb_kpconv_deform_fwd(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                                     const float* K_points, const float* offsets, const float* modulations, const float* K_values,
                                     int n, int n0, int W, int Cin, int Cout, int K, float KP_extent, int influence, int closest,
                                     float* out, void* stream) {
    if (modulations && !offsets) return SGB_ERR_INVALID;
    if (n < 0 || n0 <= 0 || W < 0 || !(KP_extent > 0.f) || influence < 0 || influence > 2) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!query_points || !support_points || (!neighbors && W > 0) || !features || !K_points || !K_values || !out) return SGB_ERR_INVALID;
    KpPlan p;
    if (!kp_plan(K, Cin, Cout, p)) return SGB_ERR_UNSUPPORTED;
    KpArgs a{query_points, support_points, neighbors, features, K_points, K_values, out, n, n0, W, Cin, Cout, K, p.T, p.rows_b,
             KP_extent, influence, closest, offsets, modulations};
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
    return sgb_kpconv_deform_bwd(g, query_points, support_points, neighbors, features, K_points, nullptr, nullptr, K_values, n, n0, W, Cin, Cout,
                                 K, KP_extent, influence, closest, gfeat, gK, nullptr, nullptr, ws, ws_bytes, stream);
}
// deformable backward: additionally goffsets [n,K,3] and gmodulations [n,K] (overwritten; NULL when the inputs are NULL)
extern "C" int sgb_kpconv_deform_bwd(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                                     const float* features, const float* K_points, const float* offsets, const float* modulations,
                                     const float* K_values, int n, int n0, int W, int Cin, int Cout, int K, float KP_extent, int influence,
                                     int closest, float* gfeat, float* gK, float* goffsets, float* gmodulations,
                                     void* ws, size_t ws_bytes, void* stream) {
    if ((modulations && !offsets) || (offsets && !goffsets) || (modulations && !gmodulations)) return SGB_ERR_INVALID;
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
             KP_extent, influence, closest, offsets, modulations};
    const int tiles = sgb_div_up(n, p.T);
    const int grid = kp_bwd_grid(n, p.T);
    KpBwdArgs b{g, gfeat, (float*)ws, tiles, offsets ? goffsets : nullptr, modulations ? gmodulations : nullptr};
    if (offsets && influence == 1) SGB_CUDA(cudaMemsetAsync(goffsets, 0, (size_t)n * K * 3 * sizeof(float), (cudaStream_t)stream));   // step function
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

// arg [n,K] <- column (0..W-1) of the neighbour closest to every deformed kernel point (KPFCNN_model.py:249-259)
extern "C" int sgb_deform_closest_neighbor(const float* query_points, const float* support_points, const int* neighbors, const float* K_points,
                                           const float* offsets, int n, int n0, int W, int K, int* arg, void* stream) {
    if (n < 0 || n0 <= 0 || W <= 0 || K <= 0) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!query_points || !support_points || !neighbors || !K_points || !offsets || !arg) return SGB_ERR_INVALID;
    { deform_closest_kernel<<<sgb_div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(query_points, support_points, neighbors, K_points, offsets, n, n0, W, K, arg); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

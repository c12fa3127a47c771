// a4/a8: per-cluster fixed-size clouds (tile + farthest point sampling), their normalisation, and the
// centred-coordinate append.  seggroup/model.py:329-395, 398-426, 429-436.
//
// FPS: one CTA per cluster.  (x, y, z, running min d^2) live as float4 in shared memory when the
// cluster has <= FPS_SMEM_PTS members, else in a caller-provided global scratch indexed by member
// position (same code path through a generic pointer).  Each pick is a block-wide arg-max with the
// numpy tie rule (first maximum).  Squared distances use __fmul_rn/__fadd_rn in numpy's order
// ((dx^2 + dy^2) + dz^2) so the picks are bit-identical to the reference.
#include "common.cuh"

namespace {
constexpr int FPS_THREADS = 256;
constexpr int FPS_SMEM_PTS = 2048;

__device__ __forceinline__ float sq_dist_np(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// block-wide argmax, ties -> lowest index.  Result broadcast to all threads.
__device__ __forceinline__ int block_argmax(float v, int i, float* s_val, int* s_idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(SGB_FULL_MASK, v, o);
        const int oi = __shfl_xor_sync(SGB_FULL_MASK, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { s_val[w] = v; s_idx[w] = i; }
    __syncthreads();
    if (w == 0) {
        v = lane < (FPS_THREADS / 32) ? s_val[lane] : -INFINITY;
        i = lane < (FPS_THREADS / 32) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(SGB_FULL_MASK, v, o);
            const int oi = __shfl_xor_sync(SGB_FULL_MASK, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (lane == 0) s_idx[0] = i;
    }
    __syncthreads();
    return s_idx[0];
}

// Round 2b: clusters of at most FPSW_MAX members with P <= FPSW_P picks — every level-1 segment of the model (~150 points, P = 64)
// — are sampled by ONE WARP each: 8 points per lane in registers (coordinates mirrored in shared memory for the broadcast of the
// picked point), the arg-max of a pick is a 5-step shuffle reduction with the same first-maximum tie rule, no block barrier at all.
// The block kernel below spent two __syncthreads and a shared-memory round trip per pick for ~150 points on 256 threads
// (0.52 ms for the 8,571 segments of an 8-scene batch); it remains the path for large clusters (phase B: P = 1024).
constexpr int FPSW_PPL = 8, FPSW_MAX = 32 * FPSW_PPL, FPSW_P = 64, FPSW_WARPS = 4;

__global__ void __launch_bounds__(FPSW_WARPS * 32)
cluster_cloud_indices_warp_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ order,
                                  const int* __restrict__ cl_off, int S, int P, int* __restrict__ cloud_idx, int* __restrict__ status) {
    __shared__ float s_x[FPSW_WARPS][FPSW_MAX], s_y[FPSW_WARPS][FPSW_MAX], s_z[FPSW_WARPS][FPSW_MAX];
    __shared__ int s_choice[FPSW_WARPS][FPSW_P];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * FPSW_WARPS + w;
    if (c >= S) return;
    const int lo = __ldg(cl_off + c), n = __ldg(cl_off + c + 1) - lo;
    if (n <= 0 || n > FPSW_MAX) return;                  // large clusters: cluster_cloud_indices_kernel
    int* out = cloud_idx + (size_t)c * P;
    const int rep = P / n, rem = P % n;
    for (int i = lane; i < rep * n; i += 32) out[i] = __ldg(order + lo + (i % n));
    if (rem == 0) return;
    // member i = lane + 32 j lives in slot j of its lane
    float px[FPSW_PPL], py[FPSW_PPL], pz[FPSW_PPL], pd[FPSW_PPL];
#pragma unroll
    for (int j = 0; j < FPSW_PPL; ++j) {
        const int i = lane + 32 * j;
        px[j] = py[j] = pz[j] = 0.f;
        if (i < n) {
            const float* p = xyz + (size_t)__ldg(order + lo + i) * stride;
            px[j] = __ldg(p); py[j] = __ldg(p + 1); pz[j] = __ldg(p + 2);
            s_x[w][i] = px[j]; s_y[w][i] = py[j]; s_z[w][i] = pz[j];
        }
    }
    __syncwarp();
    auto warp_argmax = [&](float v, int i) {             // ties -> lowest index; result in every lane
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(SGB_FULL_MASK, v, o);
            const int oi = __shfl_xor_sync(SGB_FULL_MASK, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        return i;
    };
    // distances to member 0 -> first pick
    float sx = s_x[w][0], sy = s_y[w][0], sz = s_z[w][0];
    float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < FPSW_PPL; ++j) {
        const int i = lane + 32 * j;
        if (i < n) {
            pd[j] = sq_dist_np(sx, sy, sz, px[j], py[j], pz[j]);
            if (pd[j] > bv) { bv = pd[j]; bi = i; }
        } else pd[j] = 0.f;
    }
    int sel = warp_argmax(bv, bi);
    if (lane == 0) s_choice[w][0] = sel;
    // skip_initial: restart the running minimum from the first pick
    sx = s_x[w][sel]; sy = s_y[w][sel]; sz = s_z[w][sel];
    bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < FPSW_PPL; ++j) {
        const int i = lane + 32 * j;
        if (i < n) {
            pd[j] = sq_dist_np(sx, sy, sz, px[j], py[j], pz[j]);
            if (pd[j] > bv) { bv = pd[j]; bi = i; }
        }
    }
    for (int k = 1; k < rem; ++k) {
        sel = warp_argmax(bv, bi);
        if (lane == 0) s_choice[w][k] = sel;
        sx = s_x[w][sel]; sy = s_y[w][sel]; sz = s_z[w][sel];
        bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < FPSW_PPL; ++j) {
            const int i = lane + 32 * j;
            if (i < n) {
                const float d = sq_dist_np(sx, sy, sz, px[j], py[j], pz[j]);
                if (d < pd[j]) pd[j] = d;
                if (pd[j] > bv) { bv = pd[j]; bi = i; }
            }
        }
    }
    __syncwarp();
    // trailing picks equal to member 0 are replaced by the leading picks (model.py:407-412), as in the block kernel
    if (lane == 0 && s_choice[w][rem - 1] == 0) {
        int j = 1;
        for (; j <= rem; ++j) if (s_choice[w][rem - j] != 0) break;
        if (j > rem) j = rem;
        const int invalid = j - 1;
        if (invalid == 0) atomicOr(status, 1);
        for (int t = 0; t < invalid; ++t) s_choice[w][rem - invalid + t] = (t < rem - invalid) ? s_choice[w][t] : 0;
    }
    __syncwarp();
    for (int i = lane; i < rem; i += 32) out[rep * n + i] = __ldg(order + lo + s_choice[w][i]);
}

__global__ void __launch_bounds__(FPS_THREADS)
cluster_cloud_indices_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ order,
                             const int* __restrict__ cl_off, int P, int small_done, int* __restrict__ cloud_idx,
                             float4* __restrict__ scratch, int* __restrict__ status) {
    __shared__ float4 s_pts[FPS_SMEM_PTS];
    __shared__ float s_val[FPS_THREADS / 32];
    __shared__ int s_idx[FPS_THREADS / 32];
    extern __shared__ int s_choice[];          // [rem]
    const int c = blockIdx.x;
    const int lo = cl_off[c], n = cl_off[c + 1] - lo;
    int* out = cloud_idx + (size_t)c * P;
    if (n <= 0) return;
    if (small_done && n <= FPSW_MAX) return;   // sampled by cluster_cloud_indices_warp_kernel
    const int rep = P / n, rem = P % n;
    for (int i = threadIdx.x; i < rep * n; i += FPS_THREADS) out[i] = __ldg(order + lo + (i % n));
    if (rem == 0) return;

    float4* pts = (n <= FPS_SMEM_PTS) ? s_pts : (scratch + lo);
    // distances to member 0
    const float* p0 = xyz + (size_t)__ldg(order + lo) * stride;
    float sx = __ldg(p0), sy = __ldg(p0 + 1), sz = __ldg(p0 + 2);
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += FPS_THREADS) {
        const float* p = xyz + (size_t)__ldg(order + lo + i) * stride;
        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        const float d = sq_dist_np(sx, sy, sz, x, y, z);
        pts[i] = make_float4(x, y, z, d);
        if (d > bv) { bv = d; bi = i; }
    }
    int sel = block_argmax(bv, bi, s_val, s_idx);      // also orders the pts[] writes before the reads below
    if (threadIdx.x == 0) s_choice[0] = sel;
    // skip_initial: restart the running minimum from the first pick
    {
        const float4 q = pts[sel];
        __syncthreads();
        sx = q.x; sy = q.y; sz = q.z;
        bv = -INFINITY; bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += FPS_THREADS) {
            float4 t = pts[i];
            t.w = sq_dist_np(sx, sy, sz, t.x, t.y, t.z);
            pts[i] = t;
            if (t.w > bv) { bv = t.w; bi = i; }
        }
    }
    for (int k = 1; k < rem; ++k) {
        sel = block_argmax(bv, bi, s_val, s_idx);
        if (threadIdx.x == 0) s_choice[k] = sel;
        const float4 q = pts[sel];
        __syncthreads();
        sx = q.x; sy = q.y; sz = q.z;
        bv = -INFINITY; bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += FPS_THREADS) {
            float4 t = pts[i];
            const float d = sq_dist_np(sx, sy, sz, t.x, t.y, t.z);
            if (d < t.w) { t.w = d; pts[i] = t; }
            if (t.w > bv) { bv = t.w; bi = i; }
        }
    }
    __syncthreads();
    // trailing picks equal to member 0 are replaced by the leading picks (model.py:407-412)
    if (threadIdx.x == 0 && s_choice[rem - 1] == 0) {
        int j = 1;
        for (; j <= rem; ++j) if (s_choice[rem - j] != 0) break;
        if (j > rem) j = rem;                              // python leaves j at its last value
        const int invalid = j - 1;
        if (invalid == 0) atomicOr(status, 1);             // the reference raises here
        // numpy assigns through a temporary when the two slices overlap (invalid > rem / 2); the overwritten tail was all
        // zeros, so a source index inside it reads 0 whatever has been written there since
        for (int t = 0; t < invalid; ++t) s_choice[rem - invalid + t] = (t < rem - invalid) ? s_choice[t] : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rem; i += FPS_THREADS) out[rep * n + i] = __ldg(order + lo + s_choice[i]);
}

// one warp per cluster: mean over the P rows, centre, scale by max |xyz|.
// The mean reproduces torch-CPU's fp32 summation order for a [P<=64, 3] reduction over dim 0 (probed: four
// interleaved row accumulators, row i -> acc[i % 4], combined ((a0+a1)+a2)+a3), so the transformed clouds —
// and the kNN ranking computed on them — are bit-identical to the reference.
__global__ void cluster_cloud_transform_kernel(const float* __restrict__ data6, const int* __restrict__ cloud_idx,
                                               int S, int P, float* __restrict__ clouds) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= S) return;
    const int* idx = cloud_idx + (size_t)c * P;
    float m = 0.f;
    if (lane < 3) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = 0;
        for (; i + 3 < P; i += 4) {
            a0 = __fadd_rn(a0, __ldg(data6 + (size_t)__ldg(idx + i) * 6 + lane));
            a1 = __fadd_rn(a1, __ldg(data6 + (size_t)__ldg(idx + i + 1) * 6 + lane));
            a2 = __fadd_rn(a2, __ldg(data6 + (size_t)__ldg(idx + i + 2) * 6 + lane));
            a3 = __fadd_rn(a3, __ldg(data6 + (size_t)__ldg(idx + i + 3) * 6 + lane));
        }
        if (i < P) a0 = __fadd_rn(a0, __ldg(data6 + (size_t)__ldg(idx + i) * 6 + lane));
        if (i + 1 < P) a1 = __fadd_rn(a1, __ldg(data6 + (size_t)__ldg(idx + i + 1) * 6 + lane));
        if (i + 2 < P) a2 = __fadd_rn(a2, __ldg(data6 + (size_t)__ldg(idx + i + 2) * 6 + lane));
        m = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3), (float)P);
    }
    const float mx = __shfl_sync(SGB_FULL_MASK, m, 0), my = __shfl_sync(SGB_FULL_MASK, m, 1), mz = __shfl_sync(SGB_FULL_MASK, m, 2);
    float amax = 0.f;
    for (int i = lane; i < P; i += 32) {
        const float* p = data6 + (size_t)__ldg(idx + i) * 6;
        amax = fmaxf(amax, fmaxf(fabsf(__fsub_rn(__ldg(p), mx)), fmaxf(fabsf(__fsub_rn(__ldg(p + 1), my)), fabsf(__fsub_rn(__ldg(p + 2), mz)))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(SGB_FULL_MASK, amax, o));
    for (int i = lane; i < P; i += 32) {
        const float* p = data6 + (size_t)__ldg(idx + i) * 6;
        float* o = clouds + ((size_t)c * P + i) * 6;
        o[0] = __fdiv_rn(__fsub_rn(__ldg(p), mx), amax);
        o[1] = __fdiv_rn(__fsub_rn(__ldg(p + 1), my), amax);
        o[2] = __fdiv_rn(__fsub_rn(__ldg(p + 2), mz), amax);
        o[3] = __ldg(p + 3); o[4] = __ldg(p + 4); o[5] = __ldg(p + 5);
    }
}

// Cluster means.  Clusters range from 20 points to a whole room (the late levels: a few dozen clusters of thousands of
// points), so one CTA per cluster starves the GPU exactly when the clusters are large.  Instead every warp takes 32
// consecutive positions, converts xyz to 64-bit fixed point (2^-36 m: |x| < 128 m and < 2^20 points per cluster keep the
// sum inside int64) and adds warp-reduced sums to its cluster with integer atomics.  Integer addition is associative, so
// the sums -- and the fp32 means rounded from them -- do not depend on the order the warps run in (deterministic), and the
// mean carries ~1e-11 m of quantisation error, far below the fp32 ulp of a coordinate.
constexpr double FIX_SCALE = 68719476736.0;           // 2^36
__global__ void __launch_bounds__(256)
cluster_sum_kernel(const float* __restrict__ data6, int N, const int* __restrict__ order, const int* __restrict__ cl_off, int S,
                   long long* __restrict__ sums /*[S,3], zeroed*/, int* __restrict__ pt_cluster /*[N]: cluster of every POINT*/) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int q0 = q - lane;
    if (q0 >= N) return;
    const bool valid = q < N;
    long long fx = 0, fy = 0, fz = 0;
    int pid = 0;
    if (valid) {
        pid = __ldg(order + q);
        const float2* p = reinterpret_cast<const float2*>(data6 + (size_t)pid * 6);
        const float2 a = __ldg(p);
        const float z = __ldg(reinterpret_cast<const float*>(p + 1));
        fx = __double2ll_rn((double)a.x * FIX_SCALE); fy = __double2ll_rn((double)a.y * FIX_SCALE); fz = __double2ll_rn((double)z * FIX_SCALE);
    }
    int c = sgb_upper_segment(cl_off, S, q0);         // one search per warp, then a short forward walk
    const int c_first_end = __ldg(cl_off + c + 1);
    if (c_first_end >= min(q0 + 32, N)) {             // the whole warp lies in one cluster (the common case)
        if (valid) pt_cluster[pid] = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(SGB_FULL_MASK, fx, o); fy += __shfl_xor_sync(SGB_FULL_MASK, fy, o); fz += __shfl_xor_sync(SGB_FULL_MASK, fz, o);
        }
        if (lane < 3) atomicAdd(reinterpret_cast<unsigned long long*>(sums + (size_t)c * 3 + lane),
                                (unsigned long long)(lane == 0 ? fx : lane == 1 ? fy : fz));
    } else if (valid) {
        while (c + 1 < S && q >= __ldg(cl_off + c + 1)) ++c;
        pt_cluster[pid] = c;
        atomicAdd(reinterpret_cast<unsigned long long*>(sums + (size_t)c * 3 + 0), (unsigned long long)fx);
        atomicAdd(reinterpret_cast<unsigned long long*>(sums + (size_t)c * 3 + 1), (unsigned long long)fy);
        atomicAdd(reinterpret_cast<unsigned long long*>(sums + (size_t)c * 3 + 2), (unsigned long long)fz);
    }
}

// Second pass in POINT order: every global access is a full-width coalesced vector access.  The first pass left the cluster of
// every point in pt_cluster (one 4-byte scatter per point instead of the nine strided scalar stores per point that the
// position-ordered version of this kernel issued: ncu showed it bound by the load/store unit, l1tex 64 %, DRAM 6.5 %).  A CTA
// stages its CZ_PTS input rows (24 B each) through shared memory with 16-byte loads, every thread builds the 9-float row of
// one point in shared memory, and the CTA streams the 36-byte rows out with 16-byte stores.  The means are looked up through
// the cluster tables (24 B sums + two offsets per cluster: a few hundred KB, L1/L2 resident).
constexpr int CZ_PTS = 256;
__global__ void __launch_bounds__(CZ_PTS)
centralize_kernel(const float* __restrict__ data6, int N, const int* __restrict__ pt_cluster, const int* __restrict__ cl_off, int S,
                  const long long* __restrict__ sums, float* __restrict__ x9) {
    __shared__ __align__(16) float s_in[CZ_PTS * 6];
    __shared__ __align__(16) float s_out[CZ_PTS * 9];
    const int p0 = blockIdx.x * CZ_PTS, t = threadIdx.x;
    const int np = min(CZ_PTS, N - p0);
    const float* src = data6 + (size_t)p0 * 6;                   // CZ_PTS * 24 B per CTA: 16-byte aligned when data6 is
    const bool vec = ((reinterpret_cast<uintptr_t>(data6) | reinterpret_cast<uintptr_t>(x9)) & 15) == 0;
    if (vec && np == CZ_PTS) {
        for (int i = t; i < CZ_PTS * 6 / 4; i += CZ_PTS) reinterpret_cast<float4*>(s_in)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = t; i < np * 6; i += CZ_PTS) s_in[i] = __ldg(src + i);
    }
    int c = 0;
    if (t < np) c = pt_cluster[p0 + t];
    if ((unsigned)c >= (unsigned)S) c = 0;                      // a point outside every cluster (order is not a full permutation): row undefined, as before
    __syncthreads();
    if (t < np) {
        const float x = s_in[t * 6], y = s_in[t * 6 + 1], z = s_in[t * 6 + 2];
        const double inv = 1.0 / (FIX_SCALE * (double)(__ldg(cl_off + c + 1) - __ldg(cl_off + c)));
        float* o = s_out + t * 9;                                // stride 9 words: odd -> conflict-free
        o[0] = x; o[1] = y; o[2] = z; o[3] = s_in[t * 6 + 3]; o[4] = s_in[t * 6 + 4]; o[5] = s_in[t * 6 + 5];
        o[6] = x - (float)((double)__ldg(sums + (size_t)c * 3) * inv);
        o[7] = y - (float)((double)__ldg(sums + (size_t)c * 3 + 1) * inv);
        o[8] = z - (float)((double)__ldg(sums + (size_t)c * 3 + 2) * inv);
    }
    __syncthreads();
    float* dst = x9 + (size_t)p0 * 9;                            // CZ_PTS * 36 B per CTA: 16-byte aligned when x9 is
    if (vec && np == CZ_PTS) {
        for (int i = t; i < CZ_PTS * 9 / 4; i += CZ_PTS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_out)[i];
    } else {
        for (int i = t; i < np * 9; i += CZ_PTS) dst[i] = s_out[i];
    }
}
}  // namespace

extern "C" size_t sgb_cluster_cloud_ws_bytes(int N) { return (size_t)(N > 0 ? N : 0) * sizeof(float4); }

extern "C" int sgb_cluster_cloud_indices(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                                         int P, int* cloud_idx, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (S < 0 || P <= 0 || stride < 3 || N < 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!xyz || !order || !cl_off || !cloud_idx || !status || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_cluster_cloud_ws_bytes(N)) return SGB_ERR_WORKSPACE;
    if ((size_t)P * sizeof(int) > 40 * 1024) return SGB_ERR_UNSUPPORTED;
    const int small = P <= FPSW_P ? 1 : 0;
    if (small) {
        cluster_cloud_indices_warp_kernel<<<sgb_div_up(S, FPSW_WARPS), FPSW_WARPS * 32, 0, (cudaStream_t)stream>>>(
            xyz, stride, order, cl_off, S, P, cloud_idx, status); SGB_COUNT_LAUNCH();
    }
    { cluster_cloud_indices_kernel<<<S, FPS_THREADS, (size_t)P * sizeof(int), (cudaStream_t)stream>>>(
        xyz, stride, order, cl_off, P, small, cloud_idx, (float4*)ws, status); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_cluster_cloud_transform(const float* data6, const int* cloud_idx, int S, int P, float* clouds, void* stream) {
    if (S < 0 || P <= 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!data6 || !cloud_idx || !clouds) return SGB_ERR_INVALID;
    { cluster_cloud_transform_kernel<<<sgb_div_up(S, 4), 128, 0, (cudaStream_t)stream>>>(data6, cloud_idx, S, P, clouds); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

/* scratch of sgb_centralize: 24 B of fixed-point sums per cluster + the cluster id of every point */
extern "C" size_t sgb_centralize_ws_bytes(int N, int S) {
    return (((size_t)(S > 0 ? S : 0) * 24 + 15) & ~(size_t)15) + (size_t)(N > 0 ? N : 0) * 4;
}

extern "C" int sgb_centralize(const float* data6, int N, const int* order, const int* cl_off, int S, float* x9,
                              void* ws, size_t ws_bytes, void* stream) {
    if (N < 0 || S < 0) return SGB_ERR_INVALID;
    if (N == 0 || S == 0) return SGB_OK;
    if (!data6 || !order || !cl_off || !x9 || !ws || ((uintptr_t)ws & 7)) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_centralize_ws_bytes(N, S)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    long long* sums = (long long*)ws;
    int* pt_cluster = (int*)((unsigned char*)ws + (((size_t)S * 24 + 15) & ~(size_t)15));
    SGB_CUDA(cudaMemsetAsync(sums, 0, (size_t)S * 3 * sizeof(long long), st));
    { cluster_sum_kernel<<<sgb_div_up(N, 256), 256, 0, st>>>(data6, N, order, cl_off, S, sums, pt_cluster); SGB_COUNT_LAUNCH(); }
    { centralize_kernel<<<sgb_div_up(N, CZ_PTS), CZ_PTS, 0, st>>>(data6, N, pt_cluster, cl_off, S, sums, x9); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// a10: segment max pooling with argmax (forward) and its scatter backward.
//
// Forward layout: one warp owns R = 32 consecutive positions of the member list, lanes own channels
// (float2 per lane -> one 256 B row per warp-load at C=64).  A running (max, first position) is kept
// in registers while the positions stay inside one segment and is merged into a 64-bit key
//   (monotone float key << 32) | ~position
// with atomicMax, so the result is independent of warp scheduling and ties resolve to the first row
// in member-list order (torch.max(dim=0) semantics).  A second tiny kernel
// decodes keys into (value, argmax row id).
// HBM traffic: 4*N*C (rows, read once) + 4*N (member ids) + 8*S*C (keys) + 8*S*C (out+argmax).
#include "common.cuh"

namespace {
constexpr int POOL_R = 32;        // positions per warp, all in flight at once
constexpr int POOL_WARPS = 8;

__device__ __forceinline__ bool better(float v, float best) {     // torch.max: NaN wins
    return v > best || (v != v && best == best);
}

template <int VEC>
__global__ void __launch_bounds__(POOL_WARPS * 32)
segment_pool_fwd_kernel(const float* __restrict__ feat, int C, const int* __restrict__ members, int n_members,
                        const int* __restrict__ offsets, int S, unsigned long long* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * POOL_WARPS + (threadIdx.x >> 5);
    const int p0 = warp * POOL_R;
    if (p0 >= n_members) return;
    const int p1 = min(p0 + POOL_R, n_members);
    const int c0 = (blockIdx.y * 32 + lane) * VEC;
    const bool active = c0 < C;
    float best[VEC];
    int best_pos[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { best[v] = 0.f; best_pos[v] = -1; }

    auto flush = [&](int s) {
        if (!active) return;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (best_pos[v] >= 0 && c0 + v < C) {
                unsigned long long key = ((unsigned long long)sgb_float_key(best[v]) << 32) | (unsigned)(~(unsigned)best_pos[v]);
                atomicMax(keys + (size_t)s * C + c0 + v, key);
            }
            best_pos[v] = -1;
        }
    };

    // one coalesced load of the warp's member ids, then all POOL_R row loads in flight at once (the rows are a gather:
    // latency, not issue rate, bounds this kernel, so memory-level parallelism per warp is what matters)
    int my_row = 0;
    if (p0 + lane < p1) my_row = members ? __ldg(members + p0 + lane) : p0 + lane;
    float val[POOL_R][VEC];
#pragma unroll
    for (int u = 0; u < POOL_R; ++u) {
        const int row = __shfl_sync(SGB_FULL_MASK, my_row, u);
        if (p0 + u < p1 && active) {
            const float* src = feat + (size_t)row * C + c0;
            if (VEC == 2) {
                float2 t = __ldg(reinterpret_cast<const float2*>(src));
                val[u][0] = t.x; val[u][VEC - 1] = t.y;
            } else {
                val[u][0] = __ldg(src);
            }
        }
    }
    // the segment lookup (a chain of ~log2(S) dependent loads) is issued AFTER the row gathers so that its latency
    // hides under theirs instead of preceding them
    int seg = sgb_upper_segment(offsets, S, p0);
    int seg_end = __ldg(offsets + seg + 1);
    if (p1 - p0 == POOL_R && seg_end >= p1) {
        // Fast path (most warps: segments are ~150 rows, a warp owns 32): all rows in ONE segment.  Two cheap passes over
        // the registers instead of a compare-and-track per element: max (FMNMX) + NaN probe (a sum is NaN iff a term is,
        // or inf - inf: either way the generic loop below takes over), then the FIRST position attaining the max.
        if (active) {
            float m[VEC], sum[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { m[v] = val[0][v]; sum[v] = val[0][v]; }
#pragma unroll
            for (int u = 1; u < POOL_R; ++u) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { m[v] = fmaxf(m[v], val[u][v]); sum[v] += val[u][v]; }
            }
            bool clean = true;
#pragma unroll
            for (int v = 0; v < VEC; ++v) clean = clean && (sum[v] == sum[v]);
            if (clean) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    int pos = POOL_R - 1;
#pragma unroll
                    for (int u = POOL_R - 2; u >= 0; --u) pos = (val[u][v] == m[v]) ? u : pos;
                    best[v] = m[v]; best_pos[v] = p0 + pos;
                }
                flush(seg);
                return;
            }
        } else {
            return;
        }
    }
#pragma unroll
    for (int u = 0; u < POOL_R; ++u) {
        const int q = p0 + u;
        if (q < p1) {
            while (q >= seg_end) {          // warp-uniform
                flush(seg);
                ++seg;
                seg_end = __ldg(offsets + seg + 1);
            }
            if (active) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (best_pos[v] < 0 || better(val[u][v], best[v])) { best[v] = val[u][v]; best_pos[v] = q; }
                }
            }
        }
    }
    flush(seg);
}

__global__ void segment_pool_decode_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ members,
                                           long long total, float* __restrict__ out, int* __restrict__ argmax) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    unsigned long long k = keys[i];
    out[i] = sgb_key_float((uint32_t)(k >> 32));
    if (argmax) {
        int pos = (int)(~(uint32_t)(k & 0xffffffffu));
        argmax[i] = members ? __ldg(members + pos) : pos;
    }
}

__global__ void segment_pool_bwd_kernel(const float* __restrict__ grad_out, const int* __restrict__ argmax, long long total,
                                        int C, float* __restrict__ grad_feat) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    // segments are disjoint and (row, c) pairs are unique per (s, c) -> at most one writer per address
    // unless two segments share a row; atomicAdd keeps that case correct as well.
    atomicAdd(grad_feat + (size_t)argmax[i] * C + c, grad_out[i]);
}
}  // namespace

extern "C" size_t sgb_segment_pool_ws_bytes(int S, int C) { return (size_t)(S > 0 ? S : 0) * (size_t)(C > 0 ? C : 0) * 8; }

extern "C" int sgb_segment_pool_max_fwd(const float* feat, int n_rows, int C, const int* members, int n_members,
                                        const int* offsets, int S, float* out, int* argmax,
                                        void* ws, size_t ws_bytes, void* stream) {
    if (S < 0 || C <= 0 || n_members < 0 || n_rows < 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!feat || !offsets || !out || !ws || n_members == 0) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_segment_pool_ws_bytes(S, C)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* keys = (unsigned long long*)ws;
    SGB_CUDA(cudaMemsetAsync(keys, 0, (size_t)S * C * 8, st));
    const int warps = sgb_div_up(n_members, POOL_R);
    const bool vec2 = (C % 2 == 0) && (((uintptr_t)feat & 7) == 0);
    dim3 grid(sgb_div_up(warps, POOL_WARPS), sgb_div_up(C, 32 * (vec2 ? 2 : 1)));
    if (vec2) { segment_pool_fwd_kernel<2><<<grid, POOL_WARPS * 32, 0, st>>>(feat, C, members, n_members, offsets, S, keys); SGB_COUNT_LAUNCH(); }
    else      { segment_pool_fwd_kernel<1><<<grid, POOL_WARPS * 32, 0, st>>>(feat, C, members, n_members, offsets, S, keys); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)S * C;
    { segment_pool_decode_kernel<<<sgb_div_up(total, 256), 256, 0, st>>>(keys, members, total, out, argmax); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_segment_pool_max_bwd(const float* grad_out, const int* argmax, int S, int C, float* grad_feat, void* stream) {
    if (S < 0 || C <= 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!grad_out || !argmax || !grad_feat) return SGB_ERR_INVALID;
    const long long total = (long long)S * C;
    { segment_pool_bwd_kernel<<<sgb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, argmax, total, C, grad_feat); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

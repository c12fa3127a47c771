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
                const uint32_t kh = best[v] != best[v] ? 0xffffffffu : sgb_float_key(best[v]);      // NaN: top of the key order
                unsigned long long key = ((unsigned long long)kh << 32) | (unsigned)(~(unsigned)best_pos[v]);
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

// ---- staged variant for the point-sized pools (C = 64 / 128, thousands of rows per warp).
// The generic kernel above keeps a warp's 32 gathered rows in REGISTERS, which caps the bytes in flight per SM at
// (resident warps x 8 KB) only while those warps sit in their load phase, and spends ~15 instructions per element.
// Here the gather is staged through shared memory with 16-byte cp.async (LDGSTS, L2-only): every warp owns a private ring
// of STAGES x 32 rows (8 KB each at C = 64), so a CTA of 8 warps keeps up to 192 KB of row data in flight regardless of
// what its warps are computing, and the arithmetic reads conflict-free LDS.64/128.  Persistent grid (one CTA per SM),
// each warp walks a contiguous range of 32-row chunks, so the segment of a position advances incrementally (one binary
// search per warp) and a (max, first position) pair is flushed with one 64-bit atomicMax per (segment part, channel).
// Per element: FSETP.GT + SEL (first position attaining the running max) + FMNMX.NaN (NaN poisons the max; the rare NaN
// case re-scans that segment part from global memory for the first NaN, torch.max semantics).
constexpr int ST_STAGES = 3;
constexpr int ST_ROWS = 32;
constexpr int ST_MIN_ROWS = 4096;                 // below this the generic kernel's single wave is just as fast

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

template <int C, int ST_WARPS>
__global__ void __launch_bounds__(ST_WARPS * 32, 1)
segment_pool_staged_kernel(const float* __restrict__ feat, const int* __restrict__ members, int n_members,
                           const int* __restrict__ offsets, int S, unsigned long long* __restrict__ keys) {
    constexpr int VEC = C / 32;                       // channels per lane in the compute phase (2 or 4)
    constexpr int CH = C / 4;                         // 16-byte chunks per row
    constexpr int RPI = 32 / CH;                      // rows per warp-wide copy instruction (2 or 1)
    constexpr int ROW_BYTES = C * 4;
    constexpr int STAGE_BYTES = ST_ROWS * ROW_BYTES;
    extern __shared__ __align__(128) unsigned char st_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* ring = st_smem + (size_t)wib * (ST_STAGES * STAGE_BYTES);
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);

    // contiguous, balanced range of chunks for this warp
    const int n_chunks = (n_members + ST_ROWS - 1) / ST_ROWS;
    const int n_warps = gridDim.x * ST_WARPS, w = blockIdx.x * ST_WARPS + wib;
    const int per = n_chunks / n_warps, rem = n_chunks % n_warps;
    const int k_begin = w * per + min(w, rem);
    const int k_count = per + (w < rem ? 1 : 0);
    if (k_count == 0) return;

    auto load_ids = [&](int k) -> int {               // member id of position (chunk k, lane); -1 past the end
        if (k >= k_count) return -1;
        const int p = (k_begin + k) * ST_ROWS + lane;
        return p < n_members ? (members ? __ldg(members + p) : p) : -1;
    };
    auto issue = [&](int k, int ids) {                // rows of chunk k -> ring stage k % STAGES
        const uint32_t dst = ring_u32 + (uint32_t)((k % ST_STAGES) * STAGE_BYTES) + (uint32_t)lane * 16u;
#pragma unroll
        for (int i = 0; i < ST_ROWS / RPI; ++i) {
            const int u = i * RPI + (RPI == 2 ? (lane >> 4) : 0);
            const int row = __shfl_sync(SGB_FULL_MASK, ids, u);
            if (row >= 0) cp_async16(dst + (uint32_t)(i * 512), feat + (size_t)row * C + (lane & (CH - 1)) * 4);
        }
        cp_async_commit();
    };

    // member ids run three chunks ahead of the row copies they feed (an id load that is consumed one iteration later
    // is a DRAM round trip on the critical path of every chunk: ncu long-scoreboard stalls)
    int ids_a, ids_b, ids_c;
    {
        const int i0 = load_ids(0), i1 = load_ids(1), i2 = load_ids(2);
        ids_a = load_ids(ST_STAGES); ids_b = load_ids(ST_STAGES + 1); ids_c = load_ids(ST_STAGES + 2);
        issue(0, i0); issue(1, i1); issue(2, i2);
    }
    static_assert(ST_STAGES == 3, "prologue written for three stages");

    const int q_begin = k_begin * ST_ROWS, q_end = min(n_members, (k_begin + k_count) * ST_ROWS);
    int seg = sgb_upper_segment(offsets, S, q_begin);
    int seg_end = __ldg(offsets + seg + 1);
    int part_begin = q_begin;                         // first position of the running (segment part)
    float m[VEC];
    int pos[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { m[v] = -INFINITY; pos[v] = q_begin; }

    auto flush = [&](int s, int part_end) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int c = lane * VEC + v;
            const bool nan = m[v] != m[v];
            if (nan) {                                // NaN in this part: the FIRST NaN wins (rare: plain re-scan)
                for (int q = part_begin; q < part_end; ++q) {
                    const int row = members ? __ldg(members + q) : q;
                    const float x = __ldg(feat + (size_t)row * C + c);
                    if (x != x) { pos[v] = q; break; }
                }
            }
            const uint32_t kh = nan ? 0xffffffffu : sgb_float_key(m[v]);             // NaN: top of the key order
            const unsigned long long key = ((unsigned long long)kh << 32) | (unsigned)(~(unsigned)pos[v]);
            atomicMax(keys + (size_t)s * C + c, key);
        }
    };

    for (int k = 0; k < k_count; ++k) {
        cp_async_wait<ST_STAGES - 1>();
        __syncwarp();
        const unsigned char* stage = ring + (k % ST_STAGES) * STAGE_BYTES + lane * (VEC * 4);
        const int q0 = (k_begin + k) * ST_ROWS;
        const int nrows = min(ST_ROWS, q_end - q0);
        if (nrows == ST_ROWS && seg_end >= q0 + ST_ROWS) {
            // whole chunk inside the running segment (the common case): no boundary checks
#pragma unroll
            for (int u = 0; u < ST_ROWS; ++u) {
                float x[VEC];
                if (VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(stage + u * ROW_BYTES); x[0] = t.x; x[VEC - 1] = t.y; }
                else { const float4 t = *reinterpret_cast<const float4*>(stage + u * ROW_BYTES); x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w; }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    pos[v] = x[v] > m[v] ? q0 + u : pos[v];
                    m[v] = fmax_nan(m[v], x[v]);
                }
            }
        } else {
            for (int u = 0; u < nrows; ++u) {
                const int q = q0 + u;
                while (q >= seg_end) {                // warp-uniform; empty segments are skipped without a flush
                    if (part_begin < seg_end) flush(seg, seg_end);
                    ++seg;
                    part_begin = max(part_begin, seg_end);
                    seg_end = __ldg(offsets + seg + 1);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { m[v] = -INFINITY; pos[v] = q; }
                }
                float x[VEC];
                if (VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(stage + u * ROW_BYTES); x[0] = t.x; x[VEC - 1] = t.y; }
                else { const float4 t = *reinterpret_cast<const float4*>(stage + u * ROW_BYTES); x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w; }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    pos[v] = x[v] > m[v] ? q : pos[v];
                    m[v] = fmax_nan(m[v], x[v]);
                }
            }
        }
        __syncwarp();                                 // every lane is done reading the stage before it is refilled
        const int ids = ids_a;
        ids_a = ids_b; ids_b = ids_c;
        ids_c = load_ids(k + ST_STAGES + 3);
        issue(k + ST_STAGES, ids);                    // commits an (empty) group past the end: keeps wait_group counting uniform
    }
    if (part_begin < q_end) flush(seg, q_end);
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
    // point-sized pools: shared-memory staged persistent kernel (rows must be 16-byte aligned)
    if ((C == 64 || C == 128) && n_members >= ST_MIN_ROWS && (((uintptr_t)feat & 15) == 0)) {
        int dev = 0, sms = 148;
        SGB_CUDA(cudaGetDevice(&dev));
        SGB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int chunks = sgb_div_up(n_members, ST_ROWS);
        const int warps = C == 64 ? 8 : 4;            // 8 (4) warps x 3 stages x 8 (16) KB = 192 KB of rows in flight per SM
        const int grid = chunks < sms * warps ? sgb_div_up(chunks, warps) : sms;
        const size_t smem = (size_t)warps * ST_STAGES * ST_ROWS * C * 4;
        if (C == 64) {
            SGB_OPT_IN_SMEM(segment_pool_staged_kernel<64, 8>);
            { segment_pool_staged_kernel<64, 8><<<grid, warps * 32, smem, st>>>(feat, members, n_members, offsets, S, keys); SGB_COUNT_LAUNCH(); }
        } else {
            SGB_OPT_IN_SMEM(segment_pool_staged_kernel<128, 4>);
            { segment_pool_staged_kernel<128, 4><<<grid, warps * 32, smem, st>>>(feat, members, n_members, offsets, S, keys); SGB_COUNT_LAUNCH(); }
        }
        const long long total = (long long)S * C;
        { segment_pool_decode_kernel<<<sgb_div_up(total, 256), 256, 0, st>>>(keys, members, total, out, argmax); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
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

// ---------------------------------------------------------------------------------------------------------------------------
// use_avg variant of aggregate_cluster_feature (seggroup/model.py:282-284): the mean over a segment's rows.  One warp per
// segment, lanes over channels (coalesced row reads), rows added in member-list order in fp32 (torch.mean sums in another,
// unspecified order: equal to ~1e-6 relative, not bitwise).  An empty segment gives NaN, as torch.mean of an empty slice.
namespace {
__global__ void __launch_bounds__(256)
segment_mean_fwd_kernel(const float* __restrict__ feat, int C, const int* __restrict__ members, const int* __restrict__ offsets, int S,
                        float* __restrict__ out) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= S) return;
    const int lo = __ldg(offsets + s), hi = __ldg(offsets + s + 1);
    const float inv = 1.f / (float)(hi - lo);          // 1 / 0 = inf -> 0 * inf = NaN below
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int q = lo; q < hi; ++q) {
            const int r = members ? __ldg(members + q) : q;
            acc += __ldg(feat + (size_t)r * C + c);
        }
        out[(size_t)s * C + c] = acc * inv;
    }
}
__global__ void __launch_bounds__(256)
segment_mean_bwd_kernel(const float* __restrict__ gout, int C, const int* __restrict__ members, const int* __restrict__ offsets, int S,
                        float* __restrict__ gfeat) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= S) return;
    const int lo = __ldg(offsets + s), hi = __ldg(offsets + s + 1);
    if (hi <= lo) return;
    const float inv = 1.f / (float)(hi - lo);
    for (int c = lane; c < C; c += 32) {
        const float g = __ldg(gout + (size_t)s * C + c) * inv;
        for (int q = lo; q < hi; ++q) {
            const int r = members ? __ldg(members + q) : q;
            atomicAdd(gfeat + (size_t)r * C + c, g);      // one owner per row in the model -> deterministic there
        }
    }
}
}  // namespace

extern "C" int sgb_segment_pool_mean_fwd(const float* feat, int n_rows, int C, const int* members, int n_members,
                                         const int* offsets, int S, float* out, void* stream) {
    if (S < 0 || C <= 0 || n_rows < 0 || n_members < 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!feat || !offsets || !out) return SGB_ERR_INVALID;
    { segment_mean_fwd_kernel<<<sgb_div_up(S, 8), 256, 0, (cudaStream_t)stream>>>(feat, C, members, offsets, S, out); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
extern "C" int sgb_segment_pool_mean_bwd(const float* grad_out, int S, int C, const int* members, int n_members, const int* offsets,
                                         float* grad_feat, void* stream) {
    if (S < 0 || C <= 0 || n_members < 0) return SGB_ERR_INVALID;
    if (S == 0) return SGB_OK;
    if (!grad_out || !offsets || !grad_feat) return SGB_ERR_INVALID;
    { segment_mean_bwd_kernel<<<sgb_div_up(S, 8), 256, 0, (cudaStream_t)stream>>>(grad_out, C, members, offsets, S, grad_feat); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// a18: hashed-voxel grid subsampling (barycentres, feature means, majority labels), single cloud or batch.
// replaces kpconv/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 and
// kpconv/tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:5-150.
//
// The reference walks the points once, inserting into std::unordered_map<size_t, SampledData>; its output order is
// the hash map's iteration order.  Here:
//   1. per batch element: min / max corner -> origin = floor(min * (1/dl)) * dl, NX, NY     (fp32, same ops)
//   2. per point: voxel key = iX + NX*iY + NX*NY*iZ from floor((p - origin) / dl), inserted into an open-addressing
//      hash table with atomicCAS; the table keeps the smallest point index of every voxel (atomicMin)
//   3. voxels are numbered by FIRST OCCURRENCE (scan over "this point is the first of its voxel") — the order in
//      which the reference creates them, and, because batch elements are contiguous, automatically grouped by batch
//   4. point ids are bucketed per voxel and every bucket is sorted ascending, so that
//   5. one thread per voxel sums its points IN INPUT ORDER in fp32 and scales by (float)(1.0 / count):
//      bit-identical barycentres (and feature means, sum / (float)count) to the reference.
// Majority label ties resolve to the smallest label (the reference: first maximum in hash-map order).
// HBM traffic: 12 N + 12 M (+ 4 N d + 4 M d features, + 4 N l + 4 M l labels) compulsory; the hash table and the
// buckets add ~40 B/point of scratch traffic.
#include "common.cuh"

namespace {
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int KEY_BITS = 44;                       // voxel key bits; batch index above

struct BatchParams { float ox, oy, oz; unsigned long long nx, ny; };

__global__ void gs_batch_offsets(const int* __restrict__ batches, int B, int N, int* __restrict__ boff, int* __restrict__ status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int b = 0; b < B; ++b) { boff[b] = s; s += batches ? batches[b] : N; }
        boff[B] = s;
        if (s != N) atomicOr(status, 4);           // batch lengths do not add up to N
    }
}

// Bounding boxes of the batch elements, grid-wide: every warp owns BND_PTS consecutive points (coalesced, BND_PTS/32 per
// lane), keeps a running (min, max) while its points stay inside one batch element and merges through monotone uint keys
// with atomicMin / atomicMax (order independent -> bit-identical to any sequential min / max).  A second tiny kernel
// turns the keys into the floats and the voxel-grid parameters.
constexpr int BND_WARPS = 4;
constexpr int BND_PTS = 256;        // points per warp

__global__ void gs_bounds_init(unsigned* __restrict__ keys, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 6) keys[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(BND_WARPS * 32)
gs_bounds_partial(const float* __restrict__ xyz, int N, const int* __restrict__ boff, int B, unsigned* __restrict__ keys /*[B][6]*/) {
    const int lane = threadIdx.x & 31;
    const int w0 = (blockIdx.x * BND_WARPS + (threadIdx.x >> 5)) * BND_PTS;
    if (w0 >= N) return;
    const int w1 = min(w0 + BND_PTS, N);
    float v[BND_PTS / 32][3];
#pragma unroll
    for (int k = 0; k < BND_PTS / 32; ++k) {
        const int i = w0 + k * 32 + lane;
        if (i < w1) { v[k][0] = __ldg(xyz + (size_t)i * 3); v[k][1] = __ldg(xyz + (size_t)i * 3 + 1); v[k][2] = __ldg(xyz + (size_t)i * 3 + 2); }
    }
    // batch element of the warp's first point (after the loads are in flight); the warp walks forward from there
    int b = B > 1 ? sgb_upper_segment(boff, B, w0) : 0;
    int b_end = __ldg(boff + b + 1);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    auto flush = [&](int bb) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[d] = fminf(mn[d], __shfl_xor_sync(SGB_FULL_MASK, mn[d], o));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(SGB_FULL_MASK, mx[d], o));
            }
        }
        const float lo = lane == 0 ? mn[0] : (lane == 1 ? mn[1] : mn[2]);
        const float hi = lane == 0 ? mx[0] : (lane == 1 ? mx[1] : mx[2]);
        if (lane < 3 && lo <= hi) {                                    // nothing accumulated: +inf > -inf
            atomicMin(keys + bb * 6 + lane, sgb_float_key(lo));
            atomicMax(keys + bb * 6 + 3 + lane, sgb_float_key(hi));
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] = INFINITY; mx[d] = -INFINITY; }
    };
#pragma unroll
    for (int k = 0; k < BND_PTS / 32; ++k) {
        const int i0 = w0 + k * 32;
        if (i0 >= w1) break;                                           // warp-uniform
        const int i = i0 + lane;
        if (min(i0 + 32, w1) <= b_end) {                               // the whole 32-point row is inside batch element b
            if (i < w1) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], v[k][d]); mx[d] = fmaxf(mx[d], v[k][d]); }
            }
        } else {                                                       // a row straddling batch elements: one element at a time
            int lo = i0;
            while (lo < min(i0 + 32, w1)) {
                while (lo >= b_end) { flush(b); ++b; b_end = __ldg(boff + b + 1); }
                const int hi = min(min(i0 + 32, w1), b_end);
                if (i >= lo && i < hi) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], v[k][d]); mx[d] = fmaxf(mx[d], v[k][d]); }
                }
                lo = hi;
            }
        }
    }
    flush(b);
}

__global__ void gs_bounds_finish(const unsigned* __restrict__ keys, const int* __restrict__ boff, int B, float dl,
                                 BatchParams* __restrict__ bp, float* __restrict__ minmax /*[B][6] or null*/) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int lo = boff[b], hi = boff[b + 1];
    float mn[3], mx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        mn[d] = hi > lo ? sgb_key_float(keys[b * 6 + d]) : INFINITY;
        mx[d] = hi > lo ? sgb_key_float(keys[b * 6 + 3 + d]) : -INFINITY;
    }
    if (minmax) for (int d = 0; d < 3; ++d) { minmax[b * 6 + d] = mn[d]; minmax[b * 6 + 3 + d] = mx[d]; }
    if (bp) {
        const float inv = __fdiv_rn(1.f, dl);                       // `1/sampleDl`
        BatchParams p;
        p.ox = __fmul_rn(floorf(__fmul_rn(mn[0], inv)), dl);
        p.oy = __fmul_rn(floorf(__fmul_rn(mn[1], inv)), dl);
        p.oz = __fmul_rn(floorf(__fmul_rn(mn[2], inv)), dl);
        p.nx = hi > lo ? (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mx[0], p.ox), dl)) + 1ull : 1ull;
        p.ny = hi > lo ? (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mx[1], p.oy), dl)) + 1ull : 1ull;
        bp[b] = p;
    }
}

int bounds_launch(const float* xyz, int N, const int* boff, int B, float dl, BatchParams* bp, float* minmax, unsigned* keys, cudaStream_t st) {
    gs_bounds_init<<<sgb_div_up(B * 6, 256), 256, 0, st>>>(keys, B); SGB_COUNT_LAUNCH();
    gs_bounds_partial<<<sgb_div_up(N, BND_WARPS * BND_PTS), BND_WARPS * 32, 0, st>>>(xyz, N, boff, B, keys); SGB_COUNT_LAUNCH();
    gs_bounds_finish<<<sgb_div_up(B, 128), 128, 0, st>>>(keys, boff, B, dl, bp, minmax); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

__device__ __forceinline__ unsigned hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__global__ void gs_insert(const float* __restrict__ xyz, int N, const int* __restrict__ boff, int B, const BatchParams* __restrict__ bp, float dl,
                          unsigned long long* __restrict__ tkeys, int* __restrict__ tfirst, int* __restrict__ tcount, unsigned tmask,
                          int* __restrict__ slot_of, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = sgb_upper_segment(boff, B, i);
    const BatchParams p = bp[b];
    const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)i * 3), p.ox), dl));
    const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)i * 3 + 1), p.oy), dl));
    const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)i * 3 + 2), p.oz), dl));
    const unsigned long long vk = ix + p.nx * iy + p.nx * p.ny * iz;
    if (vk >> KEY_BITS) atomicOr(status, 8);                             // grid too fine for the packed key
    const unsigned long long key = ((unsigned long long)b << KEY_BITS) | (vk & ((1ull << KEY_BITS) - 1));
    unsigned h = hash64(key) & tmask;
    while (true) {
        const unsigned long long prev = atomicCAS(tkeys + h, EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) break;
        h = (h + 1) & tmask;
    }
    slot_of[i] = (int)h;
    atomicMin(tfirst + h, i);
    atomicAdd(tcount + h, 1);
}

__global__ void gs_flag_first(const int* __restrict__ slot_of, const int* __restrict__ tfirst, int N, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flag[i] = tfirst[slot_of[i]] == i ? 1 : 0;
}
// rank = exclusive scan of flag.  first points publish voxel id / count of their slot, and count voxels per batch
// The per-batch voxel counts are aggregated warp -> CTA -> one global atomic per (CTA, batch element): with one atomicAdd per
// voxel every first point of a cloud hit the SAME address (B = 1: 358k serialised atomics, 240 us of a 400 us entry in ncu).
__global__ void __launch_bounds__(256)
gs_number_voxels(const int* __restrict__ slot_of, const int* __restrict__ flag, const int* __restrict__ rank, int N,
                 const int* __restrict__ tcount, int* __restrict__ tvox, int* __restrict__ vcount, int* __restrict__ vfirst,
                 const int* __restrict__ boff, int B, int* __restrict__ out_batches, int* __restrict__ counts) {
    __shared__ int s_cnt, s_b0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counts[0] = rank[N];
    const bool first = i < N && flag[i];
    if (first) {
        const int v = rank[i], s = slot_of[i];
        tvox[s] = v;
        vcount[v] = tcount[s];
        vfirst[v] = i;
    }
    if (!out_batches) return;                                  // kernel-uniform
    if (threadIdx.x == 0) { s_cnt = 0; s_b0 = sgb_upper_segment(boff, B, min(i, N - 1)); }
    __syncthreads();
    const int b0 = s_b0;
    // points are grouped by batch element, so almost every CTA lies inside one element: its first points are counted in
    // shared memory; the few threads of a CTA that straddles a boundary fall back to a global atomic
    int b = -1;
    if (first) b = (i >= __ldg(boff + b0) && i < __ldg(boff + b0 + 1)) ? b0 : sgb_upper_segment(boff, B, i);
    const unsigned m = __ballot_sync(SGB_FULL_MASK, b == b0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    if (b >= 0 && b != b0) atomicAdd(out_batches + b, 1);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(out_batches + b0, s_cnt);
}
__global__ void gs_bucket(const int* __restrict__ slot_of, const int* __restrict__ tvox, int N, const int* __restrict__ voff,
                          int* __restrict__ cursor, int* __restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int v = tvox[slot_of[i]];
    list[voff[v] + atomicAdd(cursor + v, 1)] = i;
}
// ascending sort of every bucket: thread-local insertion sort for small buckets, big ones are queued
__global__ void gs_sort_small(const int* __restrict__ voff, const int* __restrict__ counts, int* __restrict__ list,
                              int* __restrict__ big, int* __restrict__ nbig) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= counts[0]) return;
    const int a = voff[v], b = voff[v + 1];
    if (b - a > 48) { big[atomicAdd(nbig, 1)] = v; return; }
    for (int i = a + 1; i < b; ++i) {
        const int k = list[i];
        int j = i - 1;
        while (j >= a && list[j] > k) { list[j + 1] = list[j]; --j; }
        list[j + 1] = k;
    }
}
// one CTA per queued (big) bucket: rank sort — every element counts the smaller ones (ids are distinct).
__global__ void __launch_bounds__(256) gs_ranksort_big(const int* __restrict__ voff, const int* __restrict__ big, const int* __restrict__ nbig,
                                                       int* __restrict__ list, int* __restrict__ scratch) {
    for (int q = blockIdx.x; q < *nbig; q += gridDim.x) {
        const int v = big[q];
        const int off = voff[v];
        const int n = voff[v + 1] - off;
        for (int i = threadIdx.x; i < n; i += blockDim.x) scratch[off + i] = list[off + i];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int x = scratch[off + i];
            int r = 0;
            for (int j = 0; j < n; ++j) r += scratch[off + j] < x;   // ids are distinct
            list[off + r] = x;
        }
        __syncthreads();
    }
}

__global__ void gs_reduce_points(const float* __restrict__ xyz, const int* __restrict__ voff, const int* __restrict__ list,
                                 const int* __restrict__ counts, float* __restrict__ out_xyz) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= counts[0]) return;
    const int a = voff[v], b = voff[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int t = a; t < b; ++t) {
        const float* p = xyz + (size_t)list[t] * 3;
        sx = __fadd_rn(sx, __ldg(p)); sy = __fadd_rn(sy, __ldg(p + 1)); sz = __fadd_rn(sz, __ldg(p + 2));
    }
    const float sc = (float)(1.0 / (double)(b - a));               // `point * (1.0 / count)`: double -> float, then fp32 multiply
    out_xyz[(size_t)v * 3] = __fmul_rn(sx, sc);
    out_xyz[(size_t)v * 3 + 1] = __fmul_rn(sy, sc);
    out_xyz[(size_t)v * 3 + 2] = __fmul_rn(sz, sc);
}
__global__ void gs_reduce_features(const float* __restrict__ feat, int fdim, const int* __restrict__ voff, const int* __restrict__ list,
                                   const int* __restrict__ counts, float* __restrict__ out_feat) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)counts[0] * fdim) return;
    const int v = (int)(t / fdim), d = (int)(t % fdim);
    const int a = voff[v], b = voff[v + 1];
    float s = 0.f;
    for (int q = a; q < b; ++q) s = __fadd_rn(s, __ldg(feat + (size_t)list[q] * fdim + d));
    out_feat[t] = __fdiv_rn(s, (float)(b - a));
}
__global__ void gs_reduce_labels(const int* __restrict__ cls, int ldim, const int* __restrict__ voff, const int* __restrict__ list,
                                 const int* __restrict__ counts, int* __restrict__ out_cls) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)counts[0] * ldim) return;
    const int v = (int)(t / ldim), d = (int)(t % ldim);
    const int a = voff[v], b = voff[v + 1];
    int best = 0, best_n = 0;
    for (int q = a; q < b; ++q) {
        const int l = __ldg(cls + (size_t)list[q] * ldim + d);
        int n = 0;
        for (int r = a; r < b; ++r) n += __ldg(cls + (size_t)list[r] * ldim + d) == l;
        if (n > best_n || (n == best_n && l < best)) { best = l; best_n = n; }
    }
    out_cls[t] = best;
}

inline unsigned table_size(int N) {
    unsigned t = 1024;
    while (t < 2u * (unsigned)N) t <<= 1;
    return t;
}
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" size_t sgb_grid_subsample_ws_bytes(int N, int B) {
    const size_t T = table_size(N > 0 ? N : 1);
    size_t b = align256(T * 8) + 3 * align256(T * 4);                       // keys, first, count, vox
    b += 9 * align256((size_t)(N + 2) * 4);                                 // slot, flag, rank, vcount, vfirst, voff, cursor, list, big/scratch
    b += align256((size_t)(N + 2) * 4);
    b += align256((size_t)(B + 1) * 4) + align256((size_t)B * sizeof(BatchParams)) + 256;
    b += align256((size_t)B * 24);                                          // bounding-box keys
    b += sgb_scan_ws_bytes(N + 1) + 256;
    return b;
}

// points [N,3]; feat [N,fdim] or NULL; cls [N,ldim] or NULL; batches [B] device lengths (NULL: one cloud).
// Outputs are sized for the worst case M = N: out_xyz [N,3], out_feat [N,fdim], out_cls [N,ldim], out_first [N] (input
// index of the first point of every voxel; may be NULL), out_batches [B] (may be NULL).  counts[0] <- M.
// status: |4 batch lengths do not sum to N, |8 voxel grid exceeds 2^44 cells.
extern "C" int sgb_grid_subsample(const float* xyz, const float* feat, const int* cls, int N, int fdim, int ldim,
                                  const int* batches, int B, float dl, float* out_xyz, float* out_feat, int* out_cls,
                                  int* out_first, int* out_batches, int* counts, int* status,
                                  void* ws, size_t ws_bytes, void* stream) {
    if (N <= 0 || B <= 0 || !(dl > 0.f) || fdim < 0 || ldim < 0) return SGB_ERR_INVALID;
    if (!xyz || !out_xyz || !counts || !status || !ws) return SGB_ERR_INVALID;
    if ((fdim && (!feat || !out_feat)) || (ldim && (!cls || !out_cls))) return SGB_ERR_INVALID;
    if (B >= (1 << (63 - KEY_BITS))) return SGB_ERR_UNSUPPORTED;
    if (ws_bytes < sgb_grid_subsample_ws_bytes(N, B)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned T = table_size(N);
    unsigned char* w = (unsigned char*)ws;
    auto take = [&](size_t bytes) { void* p = w; w += align256(bytes); return p; };
    unsigned long long* tkeys = (unsigned long long*)take((size_t)T * 8);
    int* tfirst = (int*)take((size_t)T * 4);
    int* tcount = (int*)take((size_t)T * 4);
    int* tvox = (int*)take((size_t)T * 4);
    int* slot_of = (int*)take((size_t)(N + 2) * 4);
    int* flag = (int*)take((size_t)(N + 2) * 4);
    int* rank = (int*)take((size_t)(N + 2) * 4);
    int* vcount = (int*)take((size_t)(N + 2) * 4);
    int* vfirst = (int*)take((size_t)(N + 2) * 4);
    int* voff = (int*)take((size_t)(N + 2) * 4);
    int* cursor = (int*)take((size_t)(N + 2) * 4);
    int* list = (int*)take((size_t)(N + 2) * 4);
    int* big = (int*)take((size_t)(N + 2) * 4);
    int* scratch = (int*)take((size_t)(N + 2) * 4);
    int* boff = (int*)take((size_t)(B + 1) * 4);
    BatchParams* bp = (BatchParams*)take((size_t)B * sizeof(BatchParams));
    int* nbig = (int*)take(256);
    unsigned* bkeys = (unsigned*)take((size_t)B * 24);
    void* scan_ws = w;
    const size_t scan_bytes = sgb_scan_ws_bytes(N + 1);

    SGB_CUDA(cudaMemsetAsync(tkeys, 0xff, (size_t)T * 8, st));
    SGB_CUDA(cudaMemsetAsync(tfirst, 0x7f, (size_t)T * 4, st));
    SGB_CUDA(cudaMemsetAsync(tcount, 0, (size_t)T * 4, st));
    SGB_CUDA(cudaMemsetAsync(vcount, 0, (size_t)(N + 2) * 4, st));
    SGB_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(N + 2) * 4, st));
    SGB_CUDA(cudaMemsetAsync(nbig, 0, 4, st));
    if (out_batches) SGB_CUDA(cudaMemsetAsync(out_batches, 0, (size_t)B * 4, st));
    const int g = sgb_div_up(N, 256);
    int rc;
    { gs_batch_offsets<<<1, 32, 0, st>>>(batches, B, N, boff, status); SGB_COUNT_LAUNCH(); }
    if ((rc = bounds_launch(xyz, N, boff, B, dl, bp, nullptr, bkeys, st))) return rc;
    { gs_insert<<<g, 256, 0, st>>>(xyz, N, boff, B, bp, dl, tkeys, tfirst, tcount, T - 1, slot_of, status); SGB_COUNT_LAUNCH(); }
    { gs_flag_first<<<g, 256, 0, st>>>(slot_of, tfirst, N, flag); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(flag, rank, N, scan_ws, scan_bytes, st))) return rc;
    { gs_number_voxels<<<g, 256, 0, st>>>(slot_of, flag, rank, N, tcount, tvox, vcount, out_first ? out_first : vfirst, boff, B, out_batches, counts); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(vcount, voff, N, scan_ws, scan_bytes, st))) return rc;
    { gs_bucket<<<g, 256, 0, st>>>(slot_of, tvox, N, voff, cursor, list); SGB_COUNT_LAUNCH(); }
    { gs_sort_small<<<g, 256, 0, st>>>(voff, counts, list, big, nbig); SGB_COUNT_LAUNCH(); }
    { gs_ranksort_big<<<148, 256, 0, st>>>(voff, big, nbig, list, scratch); SGB_COUNT_LAUNCH(); }
    { gs_reduce_points<<<g, 256, 0, st>>>(xyz, voff, list, counts, out_xyz); SGB_COUNT_LAUNCH(); }
    if (fdim) { gs_reduce_features<<<sgb_div_up((long long)N * fdim, 256), 256, 0, st>>>(feat, fdim, voff, list, counts, out_feat); SGB_COUNT_LAUNCH(); }
    if (ldim) { gs_reduce_labels<<<sgb_div_up((long long)N * ldim, 256), 256, 0, st>>>(cls, ldim, voff, list, counts, out_cls); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// min / max corner per batch element (used by the neighbour grid as well): minmax [B][6]
extern "C" int sgb_batch_bounds(const float* xyz, int N, const int* batches, int B, float* minmax, int* boff_out /*[B+1]*/, int* status, void* stream) {
    if (N <= 0 || B <= 0 || !xyz || !minmax || !boff_out || !status) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    { gs_batch_offsets<<<1, 32, 0, st>>>(batches, B, N, boff_out, status); SGB_COUNT_LAUNCH(); }
    // the uint keys are accumulated in place: minmax [B][6] is 4-byte words either way, decoded by the finish kernel
    return bounds_launch(xyz, N, boff_out, B, 1.f, nullptr, minmax, reinterpret_cast<unsigned*>(minmax), st);
}

// a19: batched radius-neighbour search with a hashed uniform grid.
// replaces kpconv/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:211-332 `batch_nanoflann_neighbors`
// (and the brute-force equivalents `ordered_neighbors` 58-123, `batch_ordered_neighbors` 125-208).
//
// The reference builds a nanoflann KD-tree per batch element and runs a sorted radius search per query.  Here the
// supports of every batch element are binned into cells of edge 1.001*r (hash table keyed by batch | cx | cy | cz),
// a query inspects the 27 cells around it, keeps the supports with d2 < r*r (d2 = dx*dx + dy*dy + dz*dz evaluated
// left to right in fp32 without FMA contraction, nanoflann's L2_Simple_Adaptor; strict '<', RadiusResultSet), and
// sorts its hits by (d2, index) with a warp-wide bitonic network in shared memory.  Rows are padded with Ns to the
// global maximum count, which is data dependent -> two entry points: _count (also builds the grid) and _fill.
// Equal-distance ties: the reference's order is std::sort's, here index-ascending (canonical form used by the tests).
// HBM traffic (compulsory): 12 (Nq + Ns) + 4 Nq W.
#include "common.cuh"

namespace {
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int RN_CAP = 1024;            // max neighbours per query held in shared memory for the sort
constexpr int RN_WARPS = 4;
constexpr float CELL_SLACK = 1.001f;    // cell edge = 1.001 r: fp32 rounding of the cell index can never hide a hit

struct GridHeader { int T; int n_cell_slots; };

__device__ __forceinline__ unsigned rn_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}
__device__ __forceinline__ unsigned long long cell_key(int b, int cx, int cy, int cz) {
    return ((unsigned long long)b << 51) | ((unsigned long long)cx << 34) | ((unsigned long long)cy << 17) | (unsigned long long)cz;
}
__device__ __forceinline__ void cell_of(const float* p, const float* mn, float inv_cell, int& cx, int& cy, int& cz) {
    // clamped so that far-away queries cannot overflow the int conversion (they see no cell either way)
    cx = (int)fminf(fmaxf(floorf((p[0] - mn[0]) * inv_cell), -2.f), 131072.f);
    cy = (int)fminf(fmaxf(floorf((p[1] - mn[1]) * inv_cell), -2.f), 131072.f);
    cz = (int)fminf(fmaxf(floorf((p[2] - mn[2]) * inv_cell), -2.f), 131072.f);
}

__global__ void rn_insert(const float* __restrict__ s, int Ns, const int* __restrict__ sboff, int B, const float* __restrict__ minmax,
                          float inv_cell, unsigned long long* __restrict__ tkeys, int* __restrict__ tcount, unsigned tmask,
                          int* __restrict__ slot_of, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    const int b = sgb_upper_segment(sboff, B, i);
    const float p[3] = {__ldg(s + (size_t)i * 3), __ldg(s + (size_t)i * 3 + 1), __ldg(s + (size_t)i * 3 + 2)};
    int cx, cy, cz;
    cell_of(p, minmax + b * 6, inv_cell, cx, cy, cz);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= (1 << 17) - 1 || cy >= (1 << 17) - 1 || cz >= (1 << 17) - 1) { atomicOr(status, 16); cx = cy = cz = 0; }
    const unsigned long long key = cell_key(b, cx, cy, cz);
    unsigned h = rn_hash(key) & tmask;
    while (true) {
        const unsigned long long prev = atomicCAS(tkeys + h, EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) break;
        h = (h + 1) & tmask;
    }
    slot_of[i] = (int)h;
    atomicAdd(tcount + h, 1);
}
__global__ void rn_bucket(int Ns, const int* __restrict__ slot_of, const int* __restrict__ toff, int* __restrict__ cursor, int* __restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    const int h = slot_of[i];
    list[toff[h] + atomicAdd(cursor + h, 1)] = i;
}

__device__ __forceinline__ int lookup(const unsigned long long* __restrict__ tkeys, unsigned tmask, unsigned long long key) {
    unsigned h = rn_hash(key) & tmask;
    while (true) {
        const unsigned long long k = tkeys[h];
        if (k == key) return (int)h;
        if (k == EMPTY_KEY) return -1;
        h = (h + 1) & tmask;
    }
}
__device__ __forceinline__ float d2_ref(const float* q, const float* s) {       // nanoflann L2_Simple_Adaptor::evalMetric
    const float dx = __fsub_rn(q[0], s[0]), dy = __fsub_rn(q[1], s[1]), dz = __fsub_rn(q[2], s[2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// one thread per query: number of supports with d2 < r2
__global__ void rn_count(const float* __restrict__ q, int Nq, const int* __restrict__ qboff, const float* __restrict__ s, int B,
                         const float* __restrict__ minmax, float inv_cell, float r2, const unsigned long long* __restrict__ tkeys,
                         unsigned tmask, const int* __restrict__ toff, const int* __restrict__ list, int* __restrict__ counts,
                         int* __restrict__ max_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if (i < Nq) {
        const int b = sgb_upper_segment(qboff, B, i);
        const float p[3] = {__ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1), __ldg(q + (size_t)i * 3 + 2)};
        int cx, cy, cz;
        cell_of(p, minmax + b * 6, inv_cell, cx, cy, cz);
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            const int x = cx + dx, y = cy + dy, z = cz + dz;
            if (x < 0 || y < 0 || z < 0 || x >= (1 << 17) - 1 || y >= (1 << 17) - 1 || z >= (1 << 17) - 1) continue;
            const int h = lookup(tkeys, tmask, cell_key(b, x, y, z));
            if (h < 0) continue;
            for (int t = toff[h]; t < toff[h + 1]; ++t) {
                const int j = list[t];
                const float sp[3] = {__ldg(s + (size_t)j * 3), __ldg(s + (size_t)j * 3 + 1), __ldg(s + (size_t)j * 3 + 2)};
                n += d2_ref(p, sp) < r2;
            }
        }
        counts[i] = n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(SGB_FULL_MASK, n, o));
    if ((threadIdx.x & 31) == 0 && n > 0) atomicMax(max_count, n);
}

// one warp per query: collect (d2, index) keys, bitonic sort, write the row
__global__ void __launch_bounds__(RN_WARPS * 32)
rn_fill(const float* __restrict__ q, int Nq, const int* __restrict__ qboff, const float* __restrict__ s, int Ns, int B,
        const float* __restrict__ minmax, float inv_cell, float r2, const unsigned long long* __restrict__ tkeys, unsigned tmask,
        const int* __restrict__ toff, const int* __restrict__ list, const int* __restrict__ counts, int W, int* __restrict__ out,
        int* __restrict__ status) {
    __shared__ unsigned long long s_keys[RN_WARPS][RN_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * RN_WARPS + warp;
    if (i >= Nq) return;
    unsigned long long* keys = s_keys[warp];
    const int cnt = counts[i];
    int* row = out + (size_t)i * W;
    if (cnt > RN_CAP || cnt > W) {
        if (lane == 0) atomicOr(status, 32);
        for (int t = lane; t < W; t += 32) row[t] = Ns;
        return;
    }
    const int b = sgb_upper_segment(qboff, B, i);
    const float p[3] = {__ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1), __ldg(q + (size_t)i * 3 + 2)};
    int cx, cy, cz;
    cell_of(p, minmax + b * 6, inv_cell, cx, cy, cz);
    int n = 0;
    for (int c = 0; c < 27; ++c) {
        const int x = cx + (c % 3) - 1, y = cy + ((c / 3) % 3) - 1, z = cz + (c / 9) - 1;
        if (x < 0 || y < 0 || z < 0 || x >= (1 << 17) - 1 || y >= (1 << 17) - 1 || z >= (1 << 17) - 1) continue;
        const int h = lookup(tkeys, tmask, cell_key(b, x, y, z));
        if (h < 0) continue;
        const int t0 = toff[h], t1 = toff[h + 1];
        for (int tb = t0; tb < t1; tb += 32) {
            const int t = tb + lane;
            bool hit = false;
            unsigned long long key = 0;
            if (t < t1) {
                const int j = list[t];
                const float sp[3] = {__ldg(s + (size_t)j * 3), __ldg(s + (size_t)j * 3 + 1), __ldg(s + (size_t)j * 3 + 2)};
                const float d2 = d2_ref(p, sp);
                hit = d2 < r2;
                key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
            }
            const unsigned m = __ballot_sync(SGB_FULL_MASK, hit);
            if (hit) keys[n + __popc(m & ((1u << lane) - 1))] = key;
            n += __popc(m);
        }
    }
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int t = n + lane; t < p2; t += 32) keys[t] = ~0ull;
    __syncwarp();
    for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < p2; t += 32) {
                const int l = t ^ j;
                if (l > t) {
                    const unsigned long long a = keys[t], c = keys[l];
                    const bool up = (t & k) == 0;
                    if ((a > c) == up) { keys[t] = c; keys[l] = a; }
                }
            }
            __syncwarp();
        }
    }
    for (int t = lane; t < W; t += 32) row[t] = t < n ? (int)(unsigned)(keys[t] & 0xffffffffull) : Ns;
}

inline unsigned rn_table_size(int N) {
    unsigned t = 1024;
    while (t < 2u * (unsigned)N) t <<= 1;
    return t;
}
inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct RnLayout {
    unsigned T;
    unsigned long long* tkeys; int* tcount; int* toff; int* cursor; int* slot_of; int* list; int* qboff; int* sboff; float* minmax; float* qminmax;
    int* counts; int* max_count; void* scan_ws; size_t scan_bytes;
};
inline RnLayout rn_layout(void* ws, int Nq, int Ns, int B) {
    RnLayout L;
    L.T = rn_table_size(Ns);
    unsigned char* w = (unsigned char*)ws;
    auto take = [&](size_t bytes) { void* p = w; w += al(bytes); return p; };
    L.tkeys = (unsigned long long*)take((size_t)L.T * 8);
    L.tcount = (int*)take((size_t)(L.T + 1) * 4);
    L.toff = (int*)take((size_t)(L.T + 1) * 4);
    L.cursor = (int*)take((size_t)(L.T + 1) * 4);
    L.slot_of = (int*)take((size_t)Ns * 4);
    L.list = (int*)take((size_t)Ns * 4);
    L.qboff = (int*)take((size_t)(B + 1) * 4);
    L.sboff = (int*)take((size_t)(B + 1) * 4);
    L.minmax = (float*)take((size_t)B * 6 * 4);
    L.qminmax = (float*)take((size_t)B * 6 * 4);
    L.counts = (int*)take((size_t)Nq * 4);
    L.max_count = (int*)take(256);
    L.scan_ws = w;
    L.scan_bytes = sgb_scan_ws_bytes((int)L.T + 1);
    return L;
}
}  // namespace

extern "C" int sgb_batch_bounds(const float* xyz, int N, const int* batches, int B, float* minmax, int* boff_out, int* status, void* stream);

extern "C" size_t sgb_radius_neighbors_ws_bytes(int Nq, int Ns, int B) {
    const size_t T = rn_table_size(Ns > 0 ? Ns : 1);
    return al(T * 8) + 3 * al((T + 1) * 4) + 2 * al((size_t)Ns * 4) + 2 * al((size_t)(B + 1) * 4) + 2 * al((size_t)B * 24) + al((size_t)Nq * 4) + 256 +
           sgb_scan_ws_bytes((int)T + 1) + 256;
}

// Phase 1: build the support grid in `ws` and count the neighbours of every query.  max_count_out (device int) <- W.
// q_batches / s_batches: device [B] lengths (NULL: a single cloud).  status |= 16: more than 2^17 cells along an axis.
extern "C" int sgb_radius_neighbors_count(const float* queries, int Nq, const float* supports, int Ns, const int* q_batches,
                                          const int* s_batches, int B, float radius, int* max_count_out, int* status,
                                          void* ws, size_t ws_bytes, void* stream) {
    if (Nq <= 0 || Ns <= 0 || B <= 0 || !(radius > 0.f) || B >= 4096) return SGB_ERR_INVALID;
    if (!queries || !supports || !max_count_out || !status || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_radius_neighbors_ws_bytes(Nq, Ns, B)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    RnLayout L = rn_layout(ws, Nq, Ns, B);
    const float inv_cell = 1.f / (radius * CELL_SLACK);
    const float r2 = radius * radius;
    SGB_CUDA(cudaMemsetAsync(L.tkeys, 0xff, (size_t)L.T * 8, st));
    SGB_CUDA(cudaMemsetAsync(L.tcount, 0, (size_t)(L.T + 1) * 4, st));
    SGB_CUDA(cudaMemsetAsync(L.cursor, 0, (size_t)(L.T + 1) * 4, st));
    SGB_CUDA(cudaMemsetAsync(L.max_count, 0, 4, st));
    int rc;
    if ((rc = sgb_batch_bounds(supports, Ns, s_batches, B, L.minmax, L.sboff, status, st))) return rc;
    // query batch offsets (the queries' own bounds are a by-product: cells are relative to the supports' corner)
    if ((rc = sgb_batch_bounds(queries, Nq, q_batches, B, L.qminmax, L.qboff, status, st))) return rc;
    { rn_insert<<<sgb_div_up(Ns, 256), 256, 0, st>>>(supports, Ns, L.sboff, B, L.minmax, inv_cell, L.tkeys, L.tcount, L.T - 1, L.slot_of, status); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(L.tcount, L.toff, (int)L.T, L.scan_ws, L.scan_bytes, st))) return rc;
    { rn_bucket<<<sgb_div_up(Ns, 256), 256, 0, st>>>(Ns, L.slot_of, L.toff, L.cursor, L.list); SGB_COUNT_LAUNCH(); }
    { rn_count<<<sgb_div_up(Nq, 128), 128, 0, st>>>(queries, Nq, L.qboff, supports, B, L.minmax, inv_cell, r2, L.tkeys, L.T - 1, L.toff, L.list,
                                                  L.counts, L.max_count); SGB_COUNT_LAUNCH(); }
    SGB_CUDA(cudaMemcpyAsync(max_count_out, L.max_count, 4, cudaMemcpyDeviceToDevice, st));
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// Phase 2: neighbors [Nq, W] (W >= the maximum count from phase 1), rows sorted by (d2, index), padded with Ns.
// `ws` must be the workspace phase 1 filled.  status |= 32: a query has more than 1024 neighbours (row left as padding).
extern "C" int sgb_radius_neighbors_fill(const float* queries, int Nq, const float* supports, int Ns, int B, float radius, int W,
                                         int* neighbors, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (Nq <= 0 || Ns <= 0 || B <= 0 || W < 0 || !(radius > 0.f)) return SGB_ERR_INVALID;
    if (W == 0) return SGB_OK;
    if (!queries || !supports || !neighbors || !status || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_radius_neighbors_ws_bytes(Nq, Ns, B)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    RnLayout L = rn_layout(ws, Nq, Ns, B);
    const float inv_cell = 1.f / (radius * CELL_SLACK);
    const float r2 = radius * radius;
    { rn_fill<<<sgb_div_up(Nq, RN_WARPS), RN_WARPS * 32, 0, st>>>(queries, Nq, L.qboff, supports, Ns, B, L.minmax, inv_cell, r2, L.tkeys, L.T - 1,
                                                               L.toff, L.list, L.counts, W, neighbors, status); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

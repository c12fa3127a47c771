// a19: batched radius-neighbour search with a hashed uniform grid.
// replaces kpconv/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:211-332 `batch_nanoflann_neighbors`
// (and the brute-force equivalents `ordered_neighbors` 58-123, `batch_ordered_neighbors` 125-208).
//
// The reference builds a nanoflann KD-tree per batch element and runs a sorted radius search per query.  Here the
// supports of every batch element are binned into cells of edge 1.001*r (hash table keyed by batch | cx | cy | cz),
// and copied cell by cell into a float4 array (x, y, z, index) so that a cell is one contiguous, coalesced read.  One
// warp per query: lanes 0..26 look up the 27 cells around it in parallel, the cell populations are prefix-summed and
// the warp walks the FLATTENED candidate list 32 at a time (every lane busy, whatever the cell sizes), keeps the
// supports with d2 < r*r (d2 = dx*dx + dy*dy + dz*dz evaluated left to right in fp32 without FMA contraction,
// nanoflann's L2_Simple_Adaptor; strict '<', RadiusResultSet), and sorts its hits by (d2, index) with a warp-wide
// bitonic network in shared memory.  Rows are padded with Ns to the global maximum count, which is data dependent ->
// two entry points: _count builds the grid, SEARCHES ONCE and parks every sorted row (up to RN_TMPW entries) in the
// workspace; _fill copies the parked rows into the [Nq, W] output (only rows longer than RN_TMPW are searched again).
// Equal-distance ties: the reference's order is std::sort's, here index-ascending (canonical form used by the tests).
// HBM traffic (compulsory): 12 (Nq + Ns) + 4 Nq W.
#include "common.cuh"

namespace {
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int RN_CAP = 1024;            // max neighbours per query held in shared memory for the sort
constexpr int RN_WARPS = 4;
constexpr int RN_TMPW = 64;             // row capacity of the parked rows between _count and _fill
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
__global__ void rn_bucket(const float* __restrict__ s, int Ns, const int* __restrict__ slot_of, const int* __restrict__ toff,
                          int* __restrict__ cursor, float4* __restrict__ sorted4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    const int h = slot_of[i];
    const float x = __ldg(s + (size_t)i * 3), y = __ldg(s + (size_t)i * 3 + 1), z = __ldg(s + (size_t)i * 3 + 2);
    sorted4[toff[h] + atomicAdd(cursor + h, 1)] = make_float4(x, y, z, __int_as_float(i));
}
__global__ void rn_offsets(const int* __restrict__ batches, int B, int N, int* __restrict__ boff, int* __restrict__ status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < B; ++b) { boff[b] = acc; acc += batches ? batches[b] : N; }
        boff[B] = acc;
        if (acc != N) atomicOr(status, 4);
    }
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

// Warp-cooperative search of query i: returns the number of hits n; the first min(n, RN_CAP) hits are left in `keys`
// (shared memory) sorted by (d2, index).  Every lane of the warp must call it (full-mask shuffles inside).
__device__ __forceinline__ int search_sorted(int i, const float* __restrict__ q, const int* __restrict__ qboff, int B,
                                             const float* __restrict__ minmax, float inv_cell, float r2,
                                             const unsigned long long* __restrict__ tkeys, unsigned tmask, const int* __restrict__ toff,
                                             const float4* __restrict__ sorted4, unsigned long long* keys) {
    const int lane = threadIdx.x & 31;
    const float p[3] = {__ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1), __ldg(q + (size_t)i * 3 + 2)};
    const int b = B > 1 ? sgb_upper_segment(qboff, B, i) : 0;
    int cx, cy, cz;
    cell_of(p, minmax + b * 6, inv_cell, cx, cy, cz);
    // lanes 0..26: one cell each
    int start = 0, len = 0;
    if (lane < 27) {
        const int x = cx + (lane % 3) - 1, y = cy + ((lane / 3) % 3) - 1, z = cz + (lane / 9) - 1;
        if (!(x < 0 || y < 0 || z < 0 || x >= (1 << 17) - 1 || y >= (1 << 17) - 1 || z >= (1 << 17) - 1)) {
            const int h = lookup(tkeys, tmask, cell_key(b, x, y, z));
            if (h >= 0) { start = toff[h]; len = toff[h + 1] - start; }
        }
    }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(SGB_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    const int total = __shfl_sync(SGB_FULL_MASK, inc, 31);
    const int exc = inc - len;                                  // first flattened candidate of this lane's cell
    int n = 0;
    for (int u0 = 0; u0 < total; u0 += 32) {                    // warp-uniform trip count
        const int u = u0 + lane;
        int c = 0;                                              // largest cell c with exc[c] <= u
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = c + step;
            const int e = __shfl_sync(SGB_FULL_MASK, exc, cand & 31);
            if (cand < 32 && e <= u) c = cand;
        }
        const int st = __shfl_sync(SGB_FULL_MASK, start, c);
        const int ex = __shfl_sync(SGB_FULL_MASK, exc, c);
        bool hit = false;
        unsigned long long key = 0;
        if (u < total) {
            const float4 sp4 = __ldg(sorted4 + st + (u - ex));
            const float sp[3] = {sp4.x, sp4.y, sp4.z};
            const float d2 = d2_ref(p, sp);
            hit = d2 < r2;
            key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(sp4.w);
        }
        const unsigned m = __ballot_sync(SGB_FULL_MASK, hit);
        const int pos = n + __popc(m & ((1u << lane) - 1));
        if (hit && pos < RN_CAP) keys[pos] = key;
        n += __popc(m);
    }
    const int ns = min(n, RN_CAP);
    int p2 = 1;
    while (p2 < ns) p2 <<= 1;
    for (int t = ns + lane; t < p2; t += 32) keys[t] = ~0ull;
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
    return n;
}

// one warp per query: search, count, park the sorted row (first RN_TMPW entries) in the workspace
__global__ void __launch_bounds__(RN_WARPS * 32)
rn_search(const float* __restrict__ q, int Nq, const int* __restrict__ qboff, int B, const float* __restrict__ minmax, float inv_cell,
          float r2, const unsigned long long* __restrict__ tkeys, unsigned tmask, const int* __restrict__ toff,
          const float4* __restrict__ sorted4, int* __restrict__ counts, int* __restrict__ max_count, int* __restrict__ rows_tmp) {
    __shared__ unsigned long long s_keys[RN_WARPS][RN_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * RN_WARPS + warp;
    if (i >= Nq) return;
    unsigned long long* keys = s_keys[warp];
    const int n = search_sorted(i, q, qboff, B, minmax, inv_cell, r2, tkeys, tmask, toff, sorted4, keys);
    if (n <= RN_TMPW) {
        int* row = rows_tmp + (size_t)i * RN_TMPW;
        for (int t = lane; t < n; t += 32) row[t] = (int)(unsigned)(keys[t] & 0xffffffffull);
    }
    if (lane == 0) {
        counts[i] = n;
        if (n > 0) atomicMax(max_count, n);
    }
}

// one warp per query: copy the parked row (or search again when it was longer than RN_TMPW), pad with Ns
__global__ void __launch_bounds__(RN_WARPS * 32)
rn_fill(const float* __restrict__ q, int Nq, const int* __restrict__ qboff, int Ns, int B,
        const float* __restrict__ minmax, float inv_cell, float r2, const unsigned long long* __restrict__ tkeys, unsigned tmask,
        const int* __restrict__ toff, const float4* __restrict__ sorted4, const int* __restrict__ counts, const int* __restrict__ rows_tmp,
        int W, int* __restrict__ out, int* __restrict__ status) {
    __shared__ unsigned long long s_keys[RN_WARPS][RN_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * RN_WARPS + warp;
    if (i >= Nq) return;
    const int cnt = counts[i];
    int* row = out + (size_t)i * W;
    if (cnt > RN_CAP || cnt > W) {
        if (lane == 0) atomicOr(status, 32);
        for (int t = lane; t < W; t += 32) row[t] = Ns;
        return;
    }
    if (cnt <= RN_TMPW) {
        const int* src = rows_tmp + (size_t)i * RN_TMPW;
        for (int t = lane; t < W; t += 32) row[t] = t < cnt ? __ldg(src + t) : Ns;
        return;
    }
    unsigned long long* keys = s_keys[warp];
    const int n = search_sorted(i, q, qboff, B, minmax, inv_cell, r2, tkeys, tmask, toff, sorted4, keys);
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
    unsigned long long* tkeys; int* tcount; int* toff; int* cursor; int* slot_of; float4* sorted4; int* qboff; int* sboff; float* minmax;
    int* counts; int* max_count; int* rows_tmp; void* scan_ws; size_t scan_bytes;
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
    L.sorted4 = (float4*)take((size_t)Ns * 16);
    L.qboff = (int*)take((size_t)(B + 1) * 4);
    L.sboff = (int*)take((size_t)(B + 1) * 4);
    L.minmax = (float*)take((size_t)B * 6 * 4);
    L.counts = (int*)take((size_t)Nq * 4);
    L.max_count = (int*)take(256);
    L.rows_tmp = (int*)take((size_t)Nq * RN_TMPW * 4);
    L.scan_ws = w;
    L.scan_bytes = sgb_scan_ws_bytes((int)L.T + 1);
    return L;
}
}  // namespace

extern "C" int sgb_batch_bounds(const float* xyz, int N, const int* batches, int B, float* minmax, int* boff_out, int* status, void* stream);

extern "C" size_t sgb_radius_neighbors_ws_bytes(int Nq, int Ns, int B) {
    const size_t T = rn_table_size(Ns > 0 ? Ns : 1);
    return al(T * 8) + 3 * al((T + 1) * 4) + al((size_t)Ns * 4) + al((size_t)Ns * 16) + 2 * al((size_t)(B + 1) * 4) + al((size_t)B * 24) +
           al((size_t)Nq * 4) + 256 + al((size_t)Nq * RN_TMPW * 4) + sgb_scan_ws_bytes((int)T + 1) + 256;
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
    { rn_offsets<<<1, 32, 0, st>>>(q_batches, B, Nq, L.qboff, status); SGB_COUNT_LAUNCH(); }      // cells are relative to the supports' corner
    { rn_insert<<<sgb_div_up(Ns, 256), 256, 0, st>>>(supports, Ns, L.sboff, B, L.minmax, inv_cell, L.tkeys, L.tcount, L.T - 1, L.slot_of, status); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(L.tcount, L.toff, (int)L.T, L.scan_ws, L.scan_bytes, st))) return rc;
    { rn_bucket<<<sgb_div_up(Ns, 256), 256, 0, st>>>(supports, Ns, L.slot_of, L.toff, L.cursor, L.sorted4); SGB_COUNT_LAUNCH(); }
    { rn_search<<<sgb_div_up(Nq, RN_WARPS), RN_WARPS * 32, 0, st>>>(queries, Nq, L.qboff, B, L.minmax, inv_cell, r2, L.tkeys, L.T - 1, L.toff,
                                                                 L.sorted4, L.counts, L.max_count, L.rows_tmp); SGB_COUNT_LAUNCH(); }
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
    { rn_fill<<<sgb_div_up(Nq, RN_WARPS), RN_WARPS * 32, 0, st>>>(queries, Nq, L.qboff, Ns, B, L.minmax, inv_cell, r2, L.tkeys, L.T - 1,
                                                               L.toff, L.sorted4, L.counts, L.rows_tmp, W, neighbors, status); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

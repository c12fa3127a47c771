// a5/a7: exact kNN inside clusters (seggroup/model.py:30-36, 512-522).
//
// The reference ranks by  score(i,j) = (-|x_j|^2 - (-2 <x_i,x_j>)) - |x_i|^2  from a matmul in fp32 and takes the top k;
// the expression is cancellation-prone, so the ranking is only reproducible by evaluating exactly that expression
// (explicit __fmul_rn/__fmaf_rn/__fsub_rn: bit-identical to torch-CPU, probed) and ranking (score desc, member position
// asc).  A brute force over every pair of a cluster costs sum n_c^2 evaluations (9e8 for a 30k-point floor cluster).
// Here every cluster is sorted along its longest bounding-box axis and each query sweeps outwards from its own place
// in that order, left and right (32 neighbouring queries share one window, see knn_sweep_kernel), keeping its k best in
// registers.  A side is finished when
//       dx^2 > -score_k + 2 err,        err = 2^-19 * max |x|^2   (bound on |score + d^2| of the fp32 expression)
// because a candidate with dx^2 beyond that cannot reach score_k: d^2 >= dx^2 and score <= -d^2 + err.  The result is
// therefore EXACTLY the brute-force top k (ties included), at ~n * (2 r_k / extent) evaluations per cluster.
// Pipeline: per-cluster bounding box -> 64-bit keys (cluster, monotone float key of the coordinate) -> radix sort
// (cub::DeviceRadixSort: a sort primitive, not a hot-path kernel) -> gather (x, y, z, |x|^2) in sorted order -> sweep.
// HBM traffic: 16*N (xyz + order) + 4*N*k (output) + 40*N (keys, sorted records); candidate re-reads come from L1/L2.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {
constexpr int KNN_THREADS = 128;

// packed fp32x2 FMA (Blackwell FFMA2), each half an IEEE fma.rn:  v = v * a + b   /   acc = a * b + acc
__device__ __forceinline__ void sgb_ffma2(float2& v, const float2 a, const float2 b) {
    unsigned long long d = *reinterpret_cast<unsigned long long*>(&v);
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    v = *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ void sgb_ffma2_acc(float2& acc, const float2 a, const float2 b) {
    unsigned long long d = *reinterpret_cast<unsigned long long*>(&acc);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    acc = *reinterpret_cast<float2*>(&d);
}

__device__ __forceinline__ float sq_norm_ref(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float score_ref(float xi, float yi, float zi, float xxi, const float4 c) {
    float m = __fmul_rn(xi, c.x);
    m = __fmaf_rn(yi, c.y, m);
    m = __fmaf_rn(zi, c.z, m);
    const float inner = __fmul_rn(-2.f, m);
    return __fsub_rn(__fsub_rn(-c.w, inner), xxi);
}

// one CTA per cluster: bounding box -> sweep axis; scene-wide max |x|^2 (positive floats order like their bit patterns)
__global__ void __launch_bounds__(KNN_THREADS)
knn_axis_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ order, const int* __restrict__ cl_off,
                int* __restrict__ axis, unsigned* __restrict__ max_sq) {
    __shared__ float s_mn[KNN_THREADS / 32][3], s_mx[KNN_THREADS / 32][3], s_b[KNN_THREADS / 32];
    const int c = blockIdx.x;
    const int lo = cl_off[c], hi = cl_off[c + 1];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, b = 0.f;
    for (int q = lo + threadIdx.x; q < hi; q += KNN_THREADS) {
        const float* p = xyz + (size_t)__ldg(order + q) * stride;
        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        b = fmaxf(b, sq_norm_ref(x, y, z));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(SGB_FULL_MASK, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(SGB_FULL_MASK, mx[a], o));
        }
        b = fmaxf(b, __shfl_xor_sync(SGB_FULL_MASK, b, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_mn[w][a] = mn[a]; s_mx[w][a] = mx[a]; }
        s_b[w] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ext[3], bb = 0.f;
        for (int a = 0; a < 3; ++a) {
            float lo_ = INFINITY, hi_ = -INFINITY;
            for (int v = 0; v < KNN_THREADS / 32; ++v) { lo_ = fminf(lo_, s_mn[v][a]); hi_ = fmaxf(hi_, s_mx[v][a]); }
            ext[a] = hi_ - lo_;
        }
        for (int v = 0; v < KNN_THREADS / 32; ++v) bb = fmaxf(bb, s_b[v]);
        int best = 0;
        if (ext[1] > ext[best]) best = 1;
        if (ext[2] > ext[best]) best = 2;
        axis[c] = best;
        atomicMax(max_sq, __float_as_uint(bb));
    }
}

__global__ void knn_keys_kernel(const float* __restrict__ xyz, int stride, int N, const int* __restrict__ order,
                                const int* __restrict__ cl_off, int S, const int* __restrict__ axis,
                                unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    const int c = sgb_upper_segment(cl_off, S, q);
    const float v = __ldg(xyz + (size_t)__ldg(order + q) * stride + axis[c]);
    keys[q] = ((unsigned long long)c << 32) | sgb_float_key(v + 0.f);       // +0: -0.0 and +0.0 share a key
    vals[q] = q;
}

// sorted index i -> record (x, y, z, |x|^2), member position and cluster
__global__ void knn_gather_kernel(const float* __restrict__ xyz, int stride, int N, const int* __restrict__ order,
                                  const unsigned long long* __restrict__ keys, const int* __restrict__ spos,
                                  float4* __restrict__ rec, int* __restrict__ cid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* p = xyz + (size_t)__ldg(order + spos[i]) * stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    rec[i] = make_float4(x, y, z, sq_norm_ref(x, y, z));
    cid[i] = (int)(keys[i] >> 32);
}

// Warp-cooperative sweep: a warp owns 32 consecutive queries of the sorted order; their candidate windows overlap almost
// completely, so the warp walks ONE shared window outwards in chunks of 32 records (coalesced load -> per-warp shared
// memory -> broadcast reads), every lane scoring every record of the chunk that lies in its own cluster.  A lane closes a
// side once the axis gap to the nearest unprocessed record on that side rules out a better score (see the file header);
// the warp stops when every lane has closed both sides.
template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
knn_sweep_kernel(int N, const int* __restrict__ order, const int* __restrict__ cl_off, const int* __restrict__ axis,
                 const float4* __restrict__ rec, const int* __restrict__ spos, const int* __restrict__ cid,
                 const unsigned* __restrict__ max_sq, int* __restrict__ knn, const int* __restrict__ scene_pt_off, int n_scenes) {
    // chunk of 32 candidate records per warp, structure-of-arrays so that two neighbouring candidates load as one float2
    __shared__ __align__(8) float s_x[KNN_THREADS / 32][32], s_y[KNN_THREADS / 32][32], s_z[KNN_THREADS / 32][32], s_w[KNN_THREADS / 32][32];
    __shared__ int s_pos[KNN_THREADS / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i0 = (blockIdx.x * (KNN_THREADS / 32) + w) * 32;
    if (i0 >= N) return;
    const int i = i0 + lane;
    const bool have = i < N;
    int lo = 0, hi = 0, pos_i = 0, ax = 0;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) {
        const int c = cid[i];
        lo = __ldg(cl_off + c); hi = __ldg(cl_off + c + 1);
        pos_i = spos[i];
        ax = axis[c];
        me = rec[i];
    }
    const int n = hi - lo;
    const bool searching = have && n > K;
    const float err2 = 2.f * 1.9073486e-6f * __uint_as_float(*max_sq);       // 2 * 2^-19 * max |x|^2
    const float cq = ax == 0 ? me.x : (ax == 1 ? me.y : me.z);
    const float2 one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
    const float2 nx2 = make_float2(-me.x, -me.x), ny2 = make_float2(-me.y, -me.y), nz2 = make_float2(-me.z, -me.z);

    // top K as 64-bit keys (monotone key of the score << 32 | ~member position): a larger key is a better candidate under the
    // ranking (score desc, member position asc), so one unsigned 64-bit comparison (two ISETP) decides, ties included
    uint32_t khi[K], klo[K];
    const uint32_t kmin = sgb_float_key(-INFINITY);
#pragma unroll
    for (int t = 0; t < K; ++t) { khi[t] = kmin; klo[t] = 0u; }            // klo 0 = position 0xffffffff: loses every tie
    float T = INFINITY;                             // stopping / rejection radius^2 = -score_k + 2 err

    auto consider = [&](int t) {                    // exact score of candidate t of the staged chunk; insertion into the top K
        const float4 cand = make_float4(s_x[w][t], s_y[w][t], s_z[w][t], s_w[w][t]);
        const float s = score_ref(me.x, me.y, me.z, me.w, cand) + 0.f;       // + 0: -0.0 and +0.0 are one score
        const uint32_t nh = sgb_float_key(s), nl = ~(uint32_t)s_pos[w][t];
        const unsigned long long nk = ((unsigned long long)nh << 32) | nl;
        auto key_at = [&](int q) { return ((unsigned long long)khi[q] << 32) | klo[q]; };
        bool c_q = nk > key_at(K - 1);              // better than the current K-th?
        if (c_q) {
            // Shift-insertion, descending order: the entries the new key beats move down by one slot, the new key takes the first
            // such slot.  Every step reads only OLD values (walking upwards from the bottom), so the K steps are independent
            // (round 1 bubbled the new entry up through a chain of 19 dependent compare-and-swaps: ~150 instructions, serial).
#pragma unroll
            for (int q = K - 1; q > 0; --q) {
                const bool c_lo = nk > key_at(q - 1);                        // beats entry q - 1 as well?
                if (c_q) { khi[q] = c_lo ? khi[q - 1] : nh; klo[q] = c_lo ? klo[q - 1] : nl; }
                c_q = c_lo;
            }
            if (c_q) { khi[0] = nh; klo[0] = nl; }
            T = -sgb_key_float(khi[K - 1]) + err2;  // +inf until K records have been seen
        }
    };

    auto process_chunk = [&](int base) {            // records [base, base + 32) of the sorted order
        const int j = base + lane;
        __syncwarp();
        if (j >= 0 && j < N) {
            const float4 r = rec[j];
            s_x[w][lane] = r.x; s_y[w][lane] = r.y; s_z[w][lane] = r.z; s_w[w][lane] = r.w; s_pos[w][lane] = spos[j];
        }
        __syncwarp();
        const int t0 = max(0, lo - base), t1 = min(32, hi - base);          // my cluster's part of the chunk
        const int u0 = max(0, -base), u1 = min(32, N - base);
        const int a0 = searching ? max(u0, t0) : 32, a1 = searching ? min(u1, t1) : 0;
        // Two candidates per step on the packed fp32x2 pipe (FFMA2): the same prefilter value as the scalar expression
        //     d2 = fma(dz, dz, fma(dy, dy, dx * dx)),   dx = cand.x - me.x   (cand * 1 + (-me) rounds like the subtraction)
        // bit for bit, so the set of candidates that reach the exact score is unchanged.  Exact-safe prefilter:
        // score <= -d^2 + err, so a record with d^2 > -score_k + 2 err can never enter.
        // The threshold T of the chunk start is used for all 32 records (T only shrinks, so this admits a superset; the exact
        // comparison in consider() decides), which keeps the filter loop branch-free and the insertion code in one place.
        if (__any_sync(SGB_FULL_MASK, a0 < a1)) {
            unsigned pass = 0;
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
                float2 dx = *reinterpret_cast<const float2*>(&s_x[w][t]);
                float2 dy = *reinterpret_cast<const float2*>(&s_y[w][t]);
                float2 dz = *reinterpret_cast<const float2*>(&s_z[w][t]);
                sgb_ffma2(dx, one2, nx2); sgb_ffma2(dy, one2, ny2); sgb_ffma2(dz, one2, nz2);
                float2 d2 = zero2;
                sgb_ffma2_acc(d2, dx, dx); sgb_ffma2_acc(d2, dy, dy); sgb_ffma2_acc(d2, dz, dz);
                pass |= (!(d2.x * 0.999999f > T) ? 1u : 0u) << t;
                pass |= (!(d2.y * 0.999999f > T) ? 2u : 0u) << t;
            }
            // keep bits [a0, a1): my cluster's records of this chunk
            const unsigned range = (a0 < a1) ? ((a1 >= 32 ? 0xffffffffu : ((1u << a1) - 1u)) & ~((1u << a0) - 1u)) : 0u;
            pass &= range;
            while (pass) {
                const int t = __ffs(pass) - 1;
                pass &= pass - 1;
                consider(t);
            }
        }
    };
    auto coord = [&](const float4 v) { return ax == 0 ? v.x : (ax == 1 ? v.y : v.z); };

    process_chunk(i0);
    int L = i0, R = i0 + 32;                        // unprocessed records: (.., L) on the left, [R, ..) on the right
    bool flip = false;
    while (true) {
        // per lane: is a side still able to contribute?
        bool open_l = false, open_r = false;
        if (searching) {
            if (L > lo) { const float d = cq - coord(rec[L - 1]); open_l = !(d * d * 0.999999f > T); }
            if (R < hi) { const float d = coord(rec[R]) - cq; open_r = !(d * d * 0.999999f > T); }
        }
        const bool any_l = __any_sync(SGB_FULL_MASK, open_l), any_r = __any_sync(SGB_FULL_MASK, open_r);
        if (!any_l && !any_r) break;
        const bool go_left = any_l && (!any_r || flip);
        flip = !flip;
        if (go_left) { L -= 32; process_chunk(L); }
        else { process_chunk(R); R += 32; }
    }
    if (!have) return;
    const int me_pt = __ldg(order + pos_i);
    int* out = knn + (size_t)me_pt * K;
    // scene batch: ids relative to the first point of the query's scene (what the reference builds for that scene alone)
    const int base = scene_pt_off ? __ldg(scene_pt_off + sgb_upper_segment(scene_pt_off, n_scenes, me_pt)) : 0;
    if (n <= K) {                                  // model.py:516-518: all members in member order, the rest stays 0
#pragma unroll
        for (int t = 0; t < K; ++t) out[t] = (t < n) ? __ldg(order + lo + t) - base : 0;
    } else {
#pragma unroll
        for (int t = 0; t < K; ++t) out[t] = __ldg(order + (int)~klo[t]) - base;
    }
}

struct KnnWs {
    unsigned long long *keys_in, *keys_out;
    int *vals_in, *vals_out, *cid, *axis;
    unsigned* max_sq;
    float4* rec;
    void* cub_tmp;
    size_t cub_bytes, total;
};
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
inline KnnWs knn_layout(void* ws, int N, int S) {
    KnnWs w;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                    (int*)nullptr, N > 0 ? N : 1, 0, 64);
    unsigned char* p = (unsigned char*)ws;
    size_t o = 0;
    w.keys_in = (unsigned long long*)(p + o); o = align256(o + (size_t)N * 8);
    w.keys_out = (unsigned long long*)(p + o); o = align256(o + (size_t)N * 8);
    w.rec = (float4*)(p + o); o = align256(o + (size_t)N * 16);
    w.vals_in = (int*)(p + o); o = align256(o + (size_t)N * 4);
    w.vals_out = (int*)(p + o); o = align256(o + (size_t)N * 4);
    w.cid = (int*)(p + o); o = align256(o + (size_t)N * 4);
    w.axis = (int*)(p + o); o = align256(o + (size_t)(S + 1) * 4);
    w.max_sq = (unsigned*)(p + o); o = align256(o + 4);
    w.cub_tmp = p + o; w.cub_bytes = cub_bytes; o = align256(o + cub_bytes);
    w.total = o;
    return w;
}
}  // namespace

extern "C" size_t sgb_cluster_knn_ws_bytes(int N, int S) { return knn_layout(nullptr, N > 0 ? N : 1, S > 0 ? S : 1).total + 256; }

extern "C" int sgb_cluster_knn(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                               int k, int* knn, void* ws, size_t ws_bytes, void* stream) {
    return sgb_cluster_knn_scenes(xyz, stride, N, order, cl_off, S, k, knn, nullptr, 1, ws, ws_bytes, stream);
}
extern "C" int sgb_cluster_knn_scenes(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                                      int k, int* knn, const int* scene_pt_off, int n_scenes, void* ws, size_t ws_bytes, void* stream) {
    if (n_scenes < 1 || (n_scenes > 1 && !scene_pt_off)) return SGB_ERR_INVALID;
    if (n_scenes == 1) scene_pt_off = nullptr;
    if (N < 0 || S < 0 || stride < 3) return SGB_ERR_INVALID;
    if (N == 0) return SGB_OK;
    if (!xyz || !order || !cl_off || !knn || S == 0 || !ws) return SGB_ERR_INVALID;
    if (k != 20 && k != 10 && k != 16) return SGB_ERR_UNSUPPORTED;
    if (ws_bytes < sgb_cluster_knn_ws_bytes(N, S)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    KnnWs w = knn_layout(ws, N, S);
    const int grid = sgb_div_up(N, KNN_THREADS);
    SGB_CUDA(cudaMemsetAsync(w.max_sq, 0, 4, st));
    { knn_axis_kernel<<<S, KNN_THREADS, 0, st>>>(xyz, stride, order, cl_off, w.axis, w.max_sq); SGB_COUNT_LAUNCH(); }
    { knn_keys_kernel<<<grid, KNN_THREADS, 0, st>>>(xyz, stride, N, order, cl_off, S, w.axis, w.keys_in, w.vals_in); SGB_COUNT_LAUNCH(); }
    int cbits = 1;
    while ((1 << cbits) < S) ++cbits;
    size_t tmp = w.cub_bytes;
    SGB_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, N, 0, 32 + cbits, st));
    { knn_gather_kernel<<<grid, KNN_THREADS, 0, st>>>(xyz, stride, N, order, w.keys_out, w.vals_out, w.rec, w.cid); SGB_COUNT_LAUNCH(); }
    if (k == 20) { knn_sweep_kernel<20><<<grid, KNN_THREADS, 0, st>>>(N, order, cl_off, w.axis, w.rec, w.vals_out, w.cid, w.max_sq, knn, scene_pt_off, n_scenes); SGB_COUNT_LAUNCH(); }
    else if (k == 10) { knn_sweep_kernel<10><<<grid, KNN_THREADS, 0, st>>>(N, order, cl_off, w.axis, w.rec, w.vals_out, w.cid, w.max_sq, knn, scene_pt_off, n_scenes); SGB_COUNT_LAUNCH(); }
    else { knn_sweep_kernel<16><<<grid, KNN_THREADS, 0, st>>>(N, order, cl_off, w.axis, w.rec, w.vals_out, w.cid, w.max_sq, knn, scene_pt_off, n_scenes); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

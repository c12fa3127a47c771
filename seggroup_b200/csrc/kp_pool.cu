// a21: the two index-pooling operators that sit beside KPConv in every strided / upsampling block
// (kpconv/models/network_blocks.py:49-81): ind_max_pool (max over the rows listed per pooling cell, shadow
// index = the column minimum) and closest_pool (copy of the first listed row, shadow index = zeros).
//
// Layout: one warp per output row, lanes own channels (float4 per lane when d % 4 == 0 -> one 512 B request per
// warp-load); the row's W indices are read with one coalesced load and broadcast by shuffle, and up to POOLW_U
// gathered rows are in flight per lane before the max is folded, so the gather is bounded by memory-level
// parallelism rather than by latency.
// HBM traffic (algorithmic): 4*n1*d (rows; every support row is listed by some cell) + 4*n2*W + 4*n2*d.
//
// Gradient (tf.reduce_max / tf.reduce_min / tf.gather): the upstream gradient of an output element is split
// EQUALLY between all listed entries that attain the maximum (duplicates of an index count once per listing);
// the share of shadow entries flows on to the rows attaining the column minimum, again split equally.
#include "common.cuh"

namespace {
constexpr int POOLW_WARPS = 8;
constexpr int POOLW_U = 8;

// ---- column minimum of x [n1,d] as monotone uint keys (order independent)
__global__ void colmin_kernel(const float* __restrict__ x, int n1, int d, unsigned* __restrict__ keys) {
    // block (32, 8): x = channel within a 32-wide slab, y = row lane; grid.x tiles channels, grid.y tiles rows
    const int c = blockIdx.x * 32 + threadIdx.x;
    unsigned best = 0xffffffffu;
    if (c < d) {
        for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < n1; r += gridDim.y * blockDim.y)
            best = min(best, sgb_float_key(__ldg(x + (size_t)r * d + c)));
    }
    __shared__ unsigned sm[8][33];
    sm[threadIdx.y][threadIdx.x] = best;
    __syncthreads();
    if (threadIdx.y == 0 && c < d) {
#pragma unroll
        for (int j = 1; j < 8; ++j) best = min(best, sm[j][threadIdx.x]);
        atomicMin(keys + c, best);
    }
}

__global__ void colmin_decode_kernel(const unsigned* __restrict__ keys, int d, float* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < d) out[c] = sgb_key_float(keys[c]);
}

template <int VEC>
__device__ __forceinline__ void vload(const float* p, float (&v)[VEC]) {
    if (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const float (&v)[VEC]) {
    if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else *p = v[0];
}

// tf.reduce_max semantics for NaN do not matter here (features are finite); plain fmaxf-free compare keeps -0/+0 as is.
template <int VEC>
__global__ void __launch_bounds__(POOLW_WARPS * 32)
ind_max_pool_fwd_kernel(const float* __restrict__ x, int n1, int d, const int* __restrict__ inds, int n2, int W,
                        const float* __restrict__ colmin, float* __restrict__ out, int G) {
    // G lanes (power of two, <= 32) own one output row; a warp holds 32/G rows.  Every loop bound below is warp-uniform,
    // so the full-mask shuffles are legal; inactive lanes only skip their loads and the store.
    const int lane = threadIdx.x & 31;
    const int lg = lane & (G - 1);
    const int row = (blockIdx.x * POOLW_WARPS + (threadIdx.x >> 5)) * (32 / G) + lane / G;
    const bool valid = row < n2;
    const int* ir = inds + (size_t)(valid ? row : 0) * W;
    const int npass = (d + G * VEC - 1) / (G * VEC);
    for (int pass = 0; pass < npass; ++pass) {
        const int c0 = (pass * G + lg) * VEC;
        const bool active = valid && c0 < d;
        float best[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) best[k] = 0.f;
        bool have = false;
        for (int w0 = 0; w0 < W; w0 += G) {
            const int my = (valid && w0 + lg < W) ? __ldg(ir + w0 + lg) : -1;
            const int cnt = min(G, W - w0);
            for (int u0 = 0; u0 < cnt; u0 += POOLW_U) {
                float v[POOLW_U][VEC];
#pragma unroll
                for (int u = 0; u < POOLW_U; ++u) {
                    const int j = __shfl_sync(SGB_FULL_MASK, my, (u0 + u) & (G - 1), G);
                    if (active && u0 + u < cnt) {
                        const float* src = (j >= 0 && j < n1) ? x + (size_t)j * d + c0 : colmin + c0;
                        vload<VEC>(src, v[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < POOLW_U; ++u) {
                    if (active && u0 + u < cnt) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) best[k] = have ? fmaxf(best[k], v[u][k]) : v[u][k];
                        have = true;
                    }
                }
            }
        }
        if (active && have) vstore<VEC>(out + (size_t)row * d + c0, best);
    }
}

// backward: thread per (row, channel); ties share the gradient equally (tf.reduce_max gradient)
__global__ void ind_max_pool_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, int n1, int d,
                                        const int* __restrict__ inds, int n2, int W, const float* __restrict__ colmin,
                                        const float* __restrict__ out, float* __restrict__ gx, float* __restrict__ gshadow) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n2 * d) return;
    const int row = (int)(i / d), c = (int)(i % d);
    const float m = out[i];
    const int* ir = inds + (size_t)row * W;
    int ties = 0;
    for (int w = 0; w < W; ++w) {
        int j = __ldg(ir + w);
        float v = (j >= 0 && j < n1) ? __ldg(x + (size_t)j * d + c) : colmin[c];
        ties += (v == m);
    }
    if (ties == 0) return;
    const float share = g[i] / (float)ties;
    for (int w = 0; w < W; ++w) {
        int j = __ldg(ir + w);
        const bool real = (j >= 0 && j < n1);
        float v = real ? __ldg(x + (size_t)j * d + c) : colmin[c];
        if (v == m) atomicAdd(real ? gx + (size_t)j * d + c : gshadow + c, share);
    }
}

// rows attaining the column minimum share the shadow gradient (tf.reduce_min gradient)
__global__ void colmin_count_kernel(const float* __restrict__ x, int n1, int d, const float* __restrict__ colmin,
                                    const float* __restrict__ gshadow, int* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n1 * d) return;
    const int c = (int)(i % d);
    if (gshadow[c] != 0.f && x[i] == colmin[c]) atomicAdd(count + c, 1);
}
__global__ void colmin_bwd_kernel(const float* __restrict__ x, int n1, int d, const float* __restrict__ colmin,
                                  const float* __restrict__ gshadow, const int* __restrict__ count, float* __restrict__ gx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n1 * d) return;
    const int c = (int)(i % d);
    const float gs = gshadow[c];
    if (gs != 0.f && x[i] == colmin[c]) gx[i] += gs / (float)count[c];
}

template <int VEC>
__global__ void __launch_bounds__(256)
closest_pool_fwd_kernel(const float* __restrict__ x, int n1, int d, const int* __restrict__ inds, int n2, int W,
                        float* __restrict__ out) {
    const int dv = d / VEC;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n2 * dv) return;
    const int row = (int)(i / dv), c0 = (int)(i % dv) * VEC;
    const int j = __ldg(inds + (size_t)row * W);
    float v[VEC];
    if (j >= 0 && j < n1) vload<VEC>(x + (size_t)j * d + c0, v);
    else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = 0.f;
    }
    vstore<VEC>(out + (size_t)row * d + c0, v);
}

__global__ void closest_pool_bwd_kernel(const float* __restrict__ g, int n1, int d, const int* __restrict__ inds, int n2, int W,
                                        float* __restrict__ gx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n2 * d) return;
    const int row = (int)(i / d), c = (int)(i % d);
    const int j = __ldg(inds + (size_t)row * W);
    if (j >= 0 && j < n1) atomicAdd(gx + (size_t)j * d + c, g[i]);      // several pooled positions may share a nearest row
}

int colmin_launch(const float* x, int n1, int d, unsigned* keys, float* colmin, cudaStream_t st) {
    SGB_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)d * 4, st));
    if (n1 > 0) {
        dim3 grid(sgb_div_up(d, 32), min(sgb_div_up(n1, 64), 1184));
        colmin_kernel<<<grid, dim3(32, 8), 0, st>>>(x, n1, d, keys); SGB_COUNT_LAUNCH();
    }
    colmin_decode_kernel<<<sgb_div_up(d, 256), 256, 0, st>>>(keys, d, colmin); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
}  // namespace

// ws: [d] uint keys | [d] float colmin | [d] float gshadow | [d] int count
extern "C" size_t sgb_ind_max_pool_ws_bytes(int d) { return (size_t)(d > 0 ? d : 0) * 16; }

extern "C" int sgb_ind_max_pool_fwd(const float* x, int n1, int d, const int* inds, int n2, int W, float* out,
                                    void* ws, size_t ws_bytes, void* stream) {
    if (n1 < 0 || n2 < 0 || d <= 0 || W <= 0) return SGB_ERR_INVALID;
    if (n2 == 0) return SGB_OK;
    if (!inds || !out || !ws || (n1 > 0 && !x)) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_ind_max_pool_ws_bytes(d)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* keys = (unsigned*)ws;
    float* colmin = (float*)ws + d;
    int rc = colmin_launch(x, n1, d, keys, colmin, st);
    if (rc != SGB_OK) return rc;
    const bool v4 = (d % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)out & 15) == 0) && (((uintptr_t)colmin & 15) == 0);
    int G = 1;
    while (G < 32 && G * (v4 ? 4 : 1) < d) G <<= 1;
    dim3 grid(sgb_div_up(n2, POOLW_WARPS * (32 / G)), 1);
    if (v4) { ind_max_pool_fwd_kernel<4><<<grid, POOLW_WARPS * 32, 0, st>>>(x, n1, d, inds, n2, W, colmin, out, G); SGB_COUNT_LAUNCH(); }
    else    { ind_max_pool_fwd_kernel<1><<<grid, POOLW_WARPS * 32, 0, st>>>(x, n1, d, inds, n2, W, colmin, out, G); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_ind_max_pool_bwd(const float* g, const float* x, int n1, int d, const int* inds, int n2, int W,
                                    const float* out, float* gx, void* ws, size_t ws_bytes, void* stream) {
    if (n1 < 0 || n2 < 0 || d <= 0 || W <= 0) return SGB_ERR_INVALID;
    if (n1 == 0) return SGB_OK;
    if (!gx || !ws || !x) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_ind_max_pool_ws_bytes(d)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    SGB_CUDA(cudaMemsetAsync(gx, 0, (size_t)n1 * d * 4, st));
    if (n2 == 0) return SGB_OK;
    if (!g || !inds || !out) return SGB_ERR_INVALID;
    unsigned* keys = (unsigned*)ws;
    float* colmin = (float*)ws + d;
    float* gshadow = (float*)ws + 2 * (size_t)d;
    int* count = (int*)ws + 3 * (size_t)d;
    int rc = colmin_launch(x, n1, d, keys, colmin, st);            // recomputed: the forward workspace is not kept alive
    if (rc != SGB_OK) return rc;
    SGB_CUDA(cudaMemsetAsync(gshadow, 0, (size_t)d * 8, st));      // gshadow + count
    const long long tot2 = (long long)n2 * d, tot1 = (long long)n1 * d;
    ind_max_pool_bwd_kernel<<<sgb_div_up(tot2, 256), 256, 0, st>>>(g, x, n1, d, inds, n2, W, colmin, out, gx, gshadow); SGB_COUNT_LAUNCH();
    colmin_count_kernel<<<sgb_div_up(tot1, 256), 256, 0, st>>>(x, n1, d, colmin, gshadow, count); SGB_COUNT_LAUNCH();
    colmin_bwd_kernel<<<sgb_div_up(tot1, 256), 256, 0, st>>>(x, n1, d, colmin, gshadow, count, gx); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_closest_pool_fwd(const float* x, int n1, int d, const int* inds, int n2, int W, float* out, void* stream) {
    if (n1 < 0 || n2 < 0 || d <= 0 || W <= 0) return SGB_ERR_INVALID;
    if (n2 == 0) return SGB_OK;
    if (!inds || !out || (n1 > 0 && !x)) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (d % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)out & 15) == 0);
    if (v4) { closest_pool_fwd_kernel<4><<<sgb_div_up((long long)n2 * (d / 4), 256), 256, 0, st>>>(x, n1, d, inds, n2, W, out); SGB_COUNT_LAUNCH(); }
    else    { closest_pool_fwd_kernel<1><<<sgb_div_up((long long)n2 * d, 256), 256, 0, st>>>(x, n1, d, inds, n2, W, out); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_closest_pool_bwd(const float* g, int n1, int d, const int* inds, int n2, int W, float* gx, void* stream) {
    if (n1 < 0 || n2 < 0 || d <= 0 || W <= 0) return SGB_ERR_INVALID;
    if (n1 == 0) return SGB_OK;
    if (!gx) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    SGB_CUDA(cudaMemsetAsync(gx, 0, (size_t)n1 * d * 4, st));
    if (n2 == 0) return SGB_OK;
    if (!g || !inds) return SGB_ERR_INVALID;
    closest_pool_bwd_kernel<<<sgb_div_up((long long)n2 * d, 256), 256, 0, st>>>(g, n1, d, inds, n2, W, gx); SGB_COUNT_LAUNCH();
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

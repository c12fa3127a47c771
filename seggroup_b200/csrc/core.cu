// Status plumbing and the int32 exclusive scan used by the CSR builders.
#include "common.cuh"
#include <atomic>

static thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};

void sgb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long sgb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int sgb_cuda_error(int code) { g_last_cuda_error = code; return SGB_ERR_CUDA; }

extern "C" int sgb_version(void) { return 100; }

extern "C" const char* sgb_status_string(int s) {
    switch (s) {
        case SGB_OK: return "ok";
        case SGB_ERR_INVALID: return "invalid argument";
        case SGB_ERR_WORKSPACE: return "workspace too small";
        case SGB_ERR_CUDA: return "CUDA runtime error";
        case SGB_ERR_UNSUPPORTED: return "unsupported shape";
        case SGB_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}
extern "C" int sgb_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" const char* sgb_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)g_last_cuda_error); }

extern "C" int sgb_device_arch(int* major, int* minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SGB_ERR_NO_DEVICE;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return SGB_ERR_NO_DEVICE;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return SGB_OK;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan: 3 phases (block sums -> scan of block sums in one CTA -> block scan + base)
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 4096

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*[32]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(SGB_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(SGB_FULL_MASK, si, o);
            if (lane >= o) si += t;
        }
        smem[lane] = si - s;            // exclusive warp bases
        if (lane == 31) smem[32] = si;  // block total
    }
    __syncthreads();
    int base = smem[w];
    if (total) *total = smem[32];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int* __restrict__ in, int n, int* __restrict__ sums) {
    __shared__ int sm[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
    int tot;
    block_exclusive_scan(s, &tot, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(int* __restrict__ sums, int nb, int* __restrict__ grand) {
    __shared__ int sm[33];
    int carry = 0;
    for (int start = 0; start < nb; start += SCAN_THREADS) {
        int i = start + threadIdx.x;
        int v = i < nb ? sums[i] : 0;
        int tot;
        int ex = block_exclusive_scan(v, &tot, sm);
        if (i < nb) sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int* __restrict__ in, int n, const int* __restrict__ sums,
                                                           const int* __restrict__ grand, int* __restrict__ out) {
    __shared__ int sm[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    int ex = block_exclusive_scan(s, nullptr, sm) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand;
}
__global__ void scan_empty(int* out) { out[0] = 0; }

// whole scan in ONE CTA (n <= SCAN_SINGLE_MAX): the CSR builders of the segment graph scan a few thousand entries, where
// three dependent launches cost more than the work.
constexpr int SCAN_SINGLE_MAX = 8 * SCAN_TILE;
__global__ void __launch_bounds__(SCAN_THREADS) scan_single(const int* __restrict__ in, int n, int* __restrict__ out) {
    __shared__ int sm[33];
    int carry = 0;
    for (int t0 = 0; t0 < n; t0 += SCAN_TILE) {
        const int base = t0 + threadIdx.x * SCAN_ITEMS;
        int v[SCAN_ITEMS];
        int s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
        int tot;
        int ex = block_exclusive_scan(s, &tot, sm) + carry;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
        carry += tot;
    }
    if (threadIdx.x == 0) out[n] = carry;
}
}  // namespace

extern "C" size_t sgb_scan_ws_bytes(int n) {
    int nb = n > 0 ? sgb_div_up(n, SCAN_TILE) : 1;
    return (size_t)(nb + 1) * sizeof(int);
}

extern "C" int sgb_exclusive_scan_i32(const int* in, int* out, int n, void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || !out) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { { scan_empty<<<1, 1, 0, st>>>(out); SGB_COUNT_LAUNCH(); } SGB_CHECK_LAUNCH(); return SGB_OK; }
    if (!in || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_scan_ws_bytes(n)) return SGB_ERR_WORKSPACE;
    if (n <= SCAN_SINGLE_MAX) {
        { scan_single<<<1, SCAN_THREADS, 0, st>>>(in, n, out); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
    int nb = sgb_div_up(n, SCAN_TILE);
    int* sums = (int*)ws;
    int* grand = sums + nb;
    { scan_tile_sums<<<nb, SCAN_THREADS, 0, st>>>(in, n, sums); SGB_COUNT_LAUNCH(); }
    { scan_block_sums<<<1, SCAN_THREADS, 0, st>>>(sums, nb, grand); SGB_COUNT_LAUNCH(); }
    { scan_tiles<<<nb, SCAN_THREADS, 0, st>>>(in, n, sums, grand, out); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

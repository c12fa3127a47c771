// N1: batch normalisation (+ residual) + LeakyReLU of the KPFCNN blocks as library kernels.
// replaces kpconv/models/network_blocks.py:147-163 `batch_norm` (tf.layers.batch_normalization, training = batch statistics over
// the points of the stacked batch, epsilon 1e-6) followed by `leaky_relu` (network_blocks.py:166-173) — the pair that ends every
// unary / simple / resnetb block (176-337, 530-581) — and the residual join `leaky_relu(features + shortcut)` of the bottleneck
// blocks.  Features are [n points, d channels] row major, so the statistics are COLUMN sums over n rows:
//     pass 1  per (row chunk, 32-column slab): fp32 partial sums of x and x^2 (or of dz and dz * zhat in the backward), rows read
//             once, coalesced along the channels; partials reduced in fp64 in a fixed order (deterministic);
//     pass 2  y = slope-activation(gamma * (x - mean) * invstd + beta + residual), one streaming pass.
// HBM traffic: forward 4 n d (stats) + 4 n d (apply read) + 4 n d (write) [+ 4 n d residual]; backward the same with dz.
#include "common.cuh"
#include "bn_moments.cuh"

namespace {
constexpr int BA_ROWS = 256;          // rows per partial-sum block
constexpr int BA_THREADS = 256;       // 32 columns x 8 row lanes

// partial[chunk][0][c] = sum a * (b ? b : 1) ... generic: s0 = sum u, s1 = sum u * v  over the rows of the chunk
//   forward:  u = x,  v = x            (sum x, sum x^2)
//   backward: u = dz, v = zhat         (sum dz, sum dz zhat),  dz = dy * act'(y)
template <bool BWD>
__global__ void __launch_bounds__(BA_THREADS)
bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y, int n, int d,
                  const float* __restrict__ stat /*[2][d] mean, invstd (backward)*/, float slope, float* __restrict__ part /*[chunks][2][d]*/) {
    __shared__ float s0[8][33], s1[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int r0 = blockIdx.y * BA_ROWS, r1 = min(n, r0 + BA_ROWS);
    float a0 = 0.f, a1 = 0.f;
    if (c < d) {
        float mean = 0.f, invstd = 0.f;
        if (BWD) { mean = stat[c]; invstd = stat[d + c]; }
        for (int r = r0 + ry; r < r1; r += 8) {
            const size_t o = (size_t)r * d + c;
            if (BWD) {
                const float dz = dy[o] * (y[o] > 0.f ? 1.f : slope);
                a0 += dz;
                a1 = fmaf(dz, (x[o] - mean) * invstd, a1);
            } else {
                const float v = x[o];
                a0 += v;
                a1 = fmaf(v, v, a1);
            }
        }
    }
    s0[ry][cx] = a0; s1[ry][cx] = a1;
    __syncthreads();
    if (ry == 0 && c < d) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) { t0 += s0[g][cx]; t1 += s1[g][cx]; }
        part[((size_t)blockIdx.y * 2 + 0) * d + c] = t0;
        part[((size_t)blockIdx.y * 2 + 1) * d + c] = t1;
    }
}

// sums [2][d] (fp64) -> stat [2][d] = mean, invstd; batch mean / biased variance out; running statistics (optional) updated with
// torch's rule (momentum m, unbiased variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int n, int d, float eps, float momentum, float* __restrict__ stat,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const double mean = sums[c] / n;
    double var = sums[d + c] / n - mean * mean;
    if (var < 0) var = 0;
    stat[c] = (float)mean;
    stat[d + c] = (float)(1.0 / sqrt(var + (double)eps));
    stat[2 * d + c] = (float)var;
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(n > 1 ? var * n / (n - 1) : var);
    }
}
// evaluation mode: stat from the running statistics
__global__ void bn_stat_from_running_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int d, float eps,
                                            float* __restrict__ stat) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    stat[c] = running_mean[c];
    stat[d + c] = 1.f / sqrtf(running_var[c] + eps);
    stat[2 * d + c] = running_var[c];
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ residual, long long total, int d, const float* __restrict__ stat,
                const float* __restrict__ gamma, const float* __restrict__ beta, float slope, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % d);
    float v = fmaf((x[i] - stat[c]) * stat[d + c], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f);
    if (residual) v += residual[i];
    y[i] = v > 0.f ? v : v * slope;
}

// dx = gamma invstd (dz - mean(dz) - zhat mean(dz zhat)),  dres = dz,  dz = dy * act'(y);  sums [2][d] = sum dz, sum dz zhat
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y, long long total, int n, int d,
                    const float* __restrict__ stat, const float* __restrict__ gamma, const double* __restrict__ sums, float slope, int batch_stats,
                    float* __restrict__ dx, float* __restrict__ dres) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % d);
    const float dz = dy[i] * (y[i] > 0.f ? 1.f : slope);
    if (dres) dres[i] = dz;
    const float g = (gamma ? gamma[c] : 1.f) * stat[d + c];
    if (batch_stats) {
        const float zhat = (x[i] - stat[c]) * stat[d + c];
        const float m0 = (float)(sums[c] / n), m1 = (float)(sums[d + c] / n);
        dx[i] = g * (dz - m0 - zhat * m1);
    } else {
        dx[i] = g * dz;
    }
}
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int d, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[d + c];
}
inline int ba_chunks(int n) { return sgb_div_up(n > 0 ? n : 1, BA_ROWS); }
}  // namespace

// workspace: partial sums [chunks][2][d] f32 + reduced sums [2][d] f64
extern "C" size_t sgb_bn_act_ws_bytes(int n, int d) {
    return (size_t)ba_chunks(n) * 2 * (size_t)(d > 0 ? d : 1) * sizeof(float) + 2 * (size_t)(d > 0 ? d : 1) * sizeof(double) + 512;
}

// y [n,d] = act(gamma * (x - mean) * invstd + beta + residual), act = LeakyReLU(slope) (slope = 1: none).
// training != 0: batch statistics (stat [3][d] <- mean, invstd, biased variance; running_mean / running_var, if given, updated with
// momentum); training == 0: running statistics.  gamma / beta / residual may be NULL.
extern "C" int sgb_bn_act_fwd(const float* x, int n, int d, const float* gamma, const float* beta, const float* residual, float eps,
                              float slope, int training, float momentum, float* running_mean, float* running_var, float* y, float* stat,
                              void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || d <= 0) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!x || !y || !stat || !ws) return SGB_ERR_INVALID;
    if (!training && (!running_mean || !running_var)) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_bn_act_ws_bytes(n, d)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = ba_chunks(n);
    float* part = (float*)ws;
    double* sums = (double*)(((uintptr_t)(part + (size_t)chunks * 2 * d) + 255) & ~(uintptr_t)255);
    if (training) {
        dim3 grid(sgb_div_up(d, 32), chunks);
        { bn_partial_kernel<false><<<grid, BA_THREADS, 0, st>>>(x, nullptr, nullptr, n, d, nullptr, slope, part); SGB_COUNT_LAUNCH(); }
        sgb_bn::reduce_partials(part, chunks, 2 * d, sums, st);
        { bn_finalize_kernel<<<sgb_div_up(d, 128), 128, 0, st>>>(sums, n, d, eps, momentum, stat, running_mean, running_var); SGB_COUNT_LAUNCH(); }
    } else {
        { bn_stat_from_running_kernel<<<sgb_div_up(d, 128), 128, 0, st>>>(running_mean, running_var, d, eps, stat); SGB_COUNT_LAUNCH(); }
    }
    const long long total = (long long)n * d;
    { bn_apply_kernel<<<sgb_div_up(total, 256), 256, 0, st>>>(x, residual, total, d, stat, gamma, beta, slope, y); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// Gradients of sgb_bn_act_fwd: dx [n,d], dres [n,d] (optional: gradient of the residual input), dgamma / dbeta [d] (optional).
// x, y, stat as given to / returned by the forward; training as in the forward.
extern "C" int sgb_bn_act_bwd(const float* dy, const float* x, const float* y, int n, int d, const float* gamma, const float* stat,
                              float slope, int training, float* dx, float* dres, float* dgamma, float* dbeta,
                              void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || d <= 0) return SGB_ERR_INVALID;
    if (n == 0) return SGB_OK;
    if (!dy || !x || !y || !stat || !dx || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_bn_act_ws_bytes(n, d)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = ba_chunks(n);
    float* part = (float*)ws;
    double* sums = (double*)(((uintptr_t)(part + (size_t)chunks * 2 * d) + 255) & ~(uintptr_t)255);
    dim3 grid(sgb_div_up(d, 32), chunks);
    { bn_partial_kernel<true><<<grid, BA_THREADS, 0, st>>>(x, dy, y, n, d, stat, slope, part); SGB_COUNT_LAUNCH(); }
    sgb_bn::reduce_partials(part, chunks, 2 * d, sums, st);
    if (dgamma || dbeta) { bn_param_grads_kernel<<<sgb_div_up(d, 128), 128, 0, st>>>(sums, d, dgamma, dbeta); SGB_COUNT_LAUNCH(); }
    const long long total = (long long)n * d;
    { bn_bwd_apply_kernel<<<sgb_div_up(total, 256), 256, 0, st>>>(x, dy, y, total, n, d, stat, gamma, sums, slope, training, dx, dres); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// a15: classifier head + loss of the pseudo-label model, fused, one CTA per scene.
// replaces seggroup/model.py:154-166 `Classifier` (Linear 256->128 no bias, BatchNorm1d in TRAINING mode — the reference never
// calls .eval() —, LeakyReLU 0.2, Dropout 0.5, Linear 128->40) applied to the per-instance features of model.py:902-921, and
// seggroup/util.py:12-29 `cross_entropy_loss` (label smoothing eps = 0.2, SUM over the instances), forward and backward.
//
// The reference runs this on ~40 rows per scene: a dozen framework kernels forward and twice that backward, each a launch that
// does microseconds of work.  In a batch of scenes the head is still PER SCENE (BatchNorm statistics over one scene's
// instances), so the eager form costs ~40 launches per scene and step and leaves the device idle in between (measured: the
// largest source of idle gaps of the batched step, profiles/r02q_profile.txt).  Here: one launch forward, one backward, grid =
// number of scenes, W1 transposed in shared memory, per-scene gradient partials summed afterwards in scene order (deterministic).
#include "common.cuh"

namespace {
constexpr int CH_IN = 256, CH_HID = 128, CH_OUT = 40;
constexpr int CH_THREADS = 256;                 // 2 row groups x 128 hidden columns
constexpr int CH_WS = CH_HID + 1;               // row stride of the transposed W1 tile: odd -> conflict-free transposed stores
constexpr float CH_SLOPE = 0.2f, CH_BN_EPS = 1e-5f, CH_SMOOTH = 0.2f;

__device__ __forceinline__ float ch_lrelu(float v) { return v > 0.f ? v : CH_SLOPE * v; }

// logits [40] in shared memory -> loss of the row; also writes softmax - target into dl[40] when dl != nullptr
__device__ float ch_row_loss(const float* lg, int gold, float* dl) {
    float m = lg[0];
    for (int c = 1; c < CH_OUT; ++c) m = fmaxf(m, lg[c]);
    float s = 0.f;
    for (int c = 0; c < CH_OUT; ++c) s += expf(lg[c] - m);
    const float lse = m + logf(s);
    float loss = 0.f;
    for (int c = 0; c < CH_OUT; ++c) {
        const float t = c == gold ? 1.f - CH_SMOOTH : CH_SMOOTH / (CH_OUT - 1);
        const float lp = lg[c] - lse;
        loss -= t * lp;
        if (dl) dl[c] = expf(lp) - t;           // sum_c t_c = 1 -> d(-sum t log_softmax)/d logit = softmax - t
    }
    return loss;
}

// grid = scenes.  feat [G,256], g_off [B+1] instance ranges, mask [G,128] (0/1 floats) or NULL, drop_scale = 1 / (1 - p).
// out: hpre [G,128] (Linear1 output), stats [B,256] (batch mean, biased variance), logits [G,40], loss_raw [B,2] (sum, count)
__global__ void __launch_bounds__(CH_THREADS)
cls_head_fwd_kernel(const float* __restrict__ feat, const int* __restrict__ g_off, const int* __restrict__ gold,
                    const float* __restrict__ W1, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ mask, float drop_scale,
                    float* __restrict__ hpre, float* __restrict__ stats, float* __restrict__ logits, float* __restrict__ loss_raw) {
    extern __shared__ __align__(16) float ch_smem[];
    float* s_w1t = ch_smem;                              // [256][CH_WS]: W1 transposed
    float* s_row = s_w1t + CH_IN * CH_WS;                // [2][256]
    float* s_h = s_row + 2 * CH_IN;                      // [2][128]
    float* s_lg = s_h + 2 * CH_HID;                      // [2][40]
    float* s_red = s_lg + 2 * CH_OUT;                    // [2][128]
    float* s_mean = s_red + 2 * CH_HID;                  // [128]
    float* s_istd = s_mean + CH_HID;                     // [128]
    __shared__ float s_loss[2];
    const int b = blockIdx.x, tid = threadIdx.x, grp = tid >> 7, j = tid & 127;
    const int g0 = g_off[b], g1 = g_off[b + 1], I = g1 - g0;
    for (int i = tid; i < CH_IN * CH_HID; i += CH_THREADS) { const int jj = i / CH_IN, k = i % CH_IN; s_w1t[k * CH_WS + jj] = __ldg(W1 + i); }   // coalesced reads, stride-129 stores
    if (tid < 2) s_loss[tid] = 0.f;
    __syncthreads();
    // ---- Linear1 + column sums
    float csum = 0.f;
    for (int r0 = g0; r0 < g1; r0 += 2) {
        const int r = r0 + grp;
        __syncthreads();
        if (r < g1) { s_row[grp * CH_IN + j] = __ldg(feat + (size_t)r * CH_IN + j); s_row[grp * CH_IN + 128 + j] = __ldg(feat + (size_t)r * CH_IN + 128 + j); }
        __syncthreads();
        if (r < g1) {
            float acc = 0.f;
            const float* x = s_row + grp * CH_IN;
#pragma unroll 8
            for (int k = 0; k < CH_IN; ++k) acc = fmaf(x[k], s_w1t[k * CH_WS + j], acc);
            hpre[(size_t)r * CH_HID + j] = acc;
            csum += acc;
        }
    }
    s_red[grp * CH_HID + j] = csum;
    __syncthreads();
    const float mean = (s_red[j] + s_red[CH_HID + j]) / (float)I;
    __syncthreads();
    float cvar = 0.f;
    for (int r = g0 + grp; r < g1; r += 2) { const float d = hpre[(size_t)r * CH_HID + j] - mean; cvar = fmaf(d, d, cvar); }
    s_red[grp * CH_HID + j] = cvar;
    __syncthreads();
    const float var = (s_red[j] + s_red[CH_HID + j]) / (float)I;        // biased (normalisation); the caller derives the unbiased one
    const float istd = rsqrtf(var + CH_BN_EPS);
    if (grp == 0) { s_mean[j] = mean; s_istd[j] = istd; stats[(size_t)b * 2 * CH_HID + j] = mean; stats[(size_t)b * 2 * CH_HID + CH_HID + j] = var; }
    const float ga = __ldg(gamma + j), be = __ldg(beta + j);
    // ---- BN -> LeakyReLU -> dropout -> Linear2 -> smoothed cross entropy
    float loss = 0.f;
    for (int r0 = g0; r0 < g1; r0 += 2) {
        const int r = r0 + grp;
        __syncthreads();
        if (r < g1) {
            const float y = fmaf((hpre[(size_t)r * CH_HID + j] - mean) * istd, ga, be);
            const float m = mask ? __ldg(mask + (size_t)r * CH_HID + j) * drop_scale : 1.f;
            s_h[grp * CH_HID + j] = ch_lrelu(y) * m;
        }
        __syncthreads();
        if (r < g1 && j < CH_OUT) {
            float acc = __ldg(b2 + j);
            const float* h = s_h + grp * CH_HID;
#pragma unroll 8
            for (int k = 0; k < CH_HID; ++k) acc = fmaf(h[k], __ldg(W2 + j * CH_HID + k), acc);
            s_lg[grp * CH_OUT + j] = acc;
            logits[(size_t)r * CH_OUT + j] = acc;
        }
        __syncthreads();
        if (r < g1 && j == 0) loss += ch_row_loss(s_lg + grp * CH_OUT, gold[r], nullptr);
    }
    if (j == 0) s_loss[grp] = loss;
    __syncthreads();
    if (tid == 0) { loss_raw[b * 2] = s_loss[0] + s_loss[1]; loss_raw[b * 2 + 1] = (float)I; }
}

// grid = scenes.  gl [B] = dL / d loss_sum of the scene.  Per-scene partial gradients (summed by the caller in scene order):
// dW1p [B,128,256], dgp / dbp [B,128], dW2p [B,40,128], db2p [B,40]; dfeat [G,256] written directly.  scratch [G,128].
__global__ void __launch_bounds__(CH_THREADS)
cls_head_bwd_kernel(const float* __restrict__ feat, const int* __restrict__ g_off, const int* __restrict__ gold,
                    const float* __restrict__ W1, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ W2, const float* __restrict__ mask, float drop_scale,
                    const float* __restrict__ hpre, const float* __restrict__ stats, const float* __restrict__ logits,
                    const float* __restrict__ gl, float* __restrict__ scratch,
                    float* __restrict__ dfeat, float* __restrict__ dW1p, float* __restrict__ dgp, float* __restrict__ dbp,
                    float* __restrict__ dW2p, float* __restrict__ db2p) {
    extern __shared__ __align__(16) float ch_smem[];
    float* s_dw1 = ch_smem;                              // [128][256] accumulators of dW1 for this scene
    float* s_row = s_dw1 + CH_HID * CH_IN;               // [2][256] feat rows
    float* s_dh = s_row + 2 * CH_IN;                     // [2][128]
    float* s_dl = s_dh + 2 * CH_HID;                     // [2][40]
    float* s_red = s_dl + 2 * CH_OUT;                    // [4][128]
    const int b = blockIdx.x, tid = threadIdx.x, grp = tid >> 7, j = tid & 127;
    const int g0 = g_off[b], g1 = g_off[b + 1], I = g1 - g0;
    const float mean = stats[(size_t)b * 2 * CH_HID + j], var = stats[(size_t)b * 2 * CH_HID + CH_HID + j];
    const float istd = rsqrtf(var + CH_BN_EPS);
    const float ga = __ldg(gamma + j), be = __ldg(beta + j);
    const float up = gl[b];
    for (int i = tid; i < CH_HID * CH_IN; i += CH_THREADS) s_dw1[i] = 0.f;
    // ---- pass 1: d logits -> dW2, db2, d hidden -> through dropout / LeakyReLU: dy; column sums for the BatchNorm backward
    float dw2[CH_OUT];
#pragma unroll
    for (int c = 0; c < CH_OUT; ++c) dw2[c] = 0.f;
    float db2 = 0.f, sum_dy = 0.f, sum_dyx = 0.f;
    for (int r0 = g0; r0 < g1; r0 += 2) {
        const int r = r0 + grp;
        __syncthreads();
        if (r < g1 && j == 0) {
            float lg[CH_OUT];
            for (int c = 0; c < CH_OUT; ++c) lg[c] = logits[(size_t)r * CH_OUT + c];
            ch_row_loss(lg, gold[r], s_dl + grp * CH_OUT);
        }
        __syncthreads();
        if (r < g1) {
            const float xhat = (hpre[(size_t)r * CH_HID + j] - mean) * istd;
            const float y = fmaf(xhat, ga, be);
            const float m = mask ? __ldg(mask + (size_t)r * CH_HID + j) * drop_scale : 1.f;
            const float d = ch_lrelu(y) * m;
            const float* dl = s_dl + grp * CH_OUT;
            float dd = 0.f;
#pragma unroll
            for (int c = 0; c < CH_OUT; ++c) {
                const float g = up * dl[c];
                dw2[c] = fmaf(g, d, dw2[c]);
                dd = fmaf(g, __ldg(W2 + c * CH_HID + j), dd);
            }
            if (j < CH_OUT) db2 += up * dl[j];
            const float dy = dd * m * (y > 0.f ? 1.f : CH_SLOPE);
            scratch[(size_t)r * CH_HID + j] = dy;
            sum_dy += dy;
            sum_dyx = fmaf(dy, xhat, sum_dyx);
        }
    }
    __syncthreads();
    s_red[grp * CH_HID + j] = sum_dy; s_red[(2 + grp) * CH_HID + j] = sum_dyx;
    __syncthreads();
    const float dbeta = s_red[j] + s_red[CH_HID + j], dgamma = s_red[2 * CH_HID + j] + s_red[3 * CH_HID + j];
    __syncthreads();
    // dW2 / db2 partials of the two row groups
    for (int c = 0; c < CH_OUT; ++c) {
        s_red[grp * CH_HID + j] = dw2[c];
        __syncthreads();
        if (grp == 0) dW2p[((size_t)b * CH_OUT + c) * CH_HID + j] = s_red[j] + s_red[CH_HID + j];
        __syncthreads();
    }
    s_red[grp * CH_HID + j] = db2;
    __syncthreads();
    if (grp == 0) {
        if (j < CH_OUT) db2p[(size_t)b * CH_OUT + j] = s_red[j] + s_red[CH_HID + j];
        dgp[(size_t)b * CH_HID + j] = dgamma;
        dbp[(size_t)b * CH_HID + j] = dbeta;
    }
    // ---- pass 2: BatchNorm backward (training statistics) -> d Linear1 output; dW1 accumulation and d feat
    const float k1 = ga * istd / (float)I;
    for (int r0 = g0; r0 < g1; r0 += 2) {
        const int r = r0 + grp;
        __syncthreads();
        if (r < g1) {
            const float xhat = (hpre[(size_t)r * CH_HID + j] - mean) * istd;
            const float dy = scratch[(size_t)r * CH_HID + j];
            s_dh[grp * CH_HID + j] = k1 * ((float)I * dy - dbeta - xhat * dgamma);
            s_row[grp * CH_IN + j] = __ldg(feat + (size_t)r * CH_IN + j); s_row[grp * CH_IN + 128 + j] = __ldg(feat + (size_t)r * CH_IN + 128 + j);
        }
        __syncthreads();
        // dW1[jj][k] += dh[jj] * x[k] for the (up to) two rows of this step: thread owns columns k = tid of every jj
        {
            const bool two = r0 + 1 < g1;
            const float x0 = s_row[tid], x1 = two ? s_row[CH_IN + tid] : 0.f;
#pragma unroll 4
            for (int jj = 0; jj < CH_HID; ++jj) {
                float a = s_dw1[jj * CH_IN + tid];
                a = fmaf(s_dh[jj], x0, a);
                if (two) a = fmaf(s_dh[CH_HID + jj], x1, a);
                s_dw1[jj * CH_IN + tid] = a;
            }
        }
        // d feat[r][k] = sum_jj dh[jj] W1[jj][k]: thread (grp, j) takes k = j and k = 128 + j of its row
        if (r < g1) {
            const float* dh = s_dh + grp * CH_HID;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
            for (int jj = 0; jj < CH_HID; ++jj) {
                const float d = dh[jj];
                a0 = fmaf(d, __ldg(W1 + jj * CH_IN + j), a0);
                a1 = fmaf(d, __ldg(W1 + jj * CH_IN + 128 + j), a1);
            }
            dfeat[(size_t)r * CH_IN + j] = a0; dfeat[(size_t)r * CH_IN + 128 + j] = a1;
        }
    }
    __syncthreads();
    for (int i = tid; i < CH_HID * CH_IN; i += CH_THREADS) dW1p[(size_t)b * CH_HID * CH_IN + i] = s_dw1[i];
}


// ---------------------------------------------------------------------------------------------
// Per-instance grouping of the final clusters (model.py:902-916): per scene `np.unique` of the clusters' weak instance labels
// (ascending), the clusters of a label in ascending id order, the semantic label of the FIRST cluster of each group.  The eager
// form (unique / argsort / bincount / cumsum / repeat_interleave on ~300 rows) is ~35 framework launches and four blocking
// read-backs per step with the device idle in between (profiles/r03d_timeline_8x150k_train.txt: 1.25 ms per 8-scene step for
// microseconds of work).  Here: ONE single-CTA launch — the scenes one after the other, rank of a cluster = number of clusters
// of its scene with a smaller (label, id) key (O(n^2) comparisons on n <= a few thousand rows), group heads by a block scan over
// the sorted order — and one 8-byte read-back (group count: it sizes the classifier's tensors).
// out: order [S] cluster ids sorted by (scene, label, id); off [G+1] positions of the group heads in `order` (off[G] = S);
//      gold [G] semantic weak label of the first cluster of the group; g_off [n_scenes+1] group range of every scene;
//      counts [2] = (G, smallest number of groups of a scene).
// ---------------------------------------------------------------------------------------------
constexpr int CG_THREADS = 1024;

__global__ void __launch_bounds__(CG_THREADS)
cls_groups_kernel(const int* __restrict__ cl_ins, const int* __restrict__ cl_sem, const int* __restrict__ scene_cl_off, int n_scenes, int S,
                  int* order, int* off, int* gold, int* g_off, int* counts) {
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    int gbase = 0, min_groups = 0x7fffffff;
    if (tid == 0) g_off[0] = 0;
    for (int b = 0; b < n_scenes; ++b) {
        const int c0 = scene_cl_off ? __ldg(scene_cl_off + b) : 0;
        const int c1 = scene_cl_off ? __ldg(scene_cl_off + b + 1) : S;
        const int n = c1 - c0;
        for (int i = tid; i < n; i += CG_THREADS) {          // rank by (label, id): a stable sort of the scene's clusters by label
            const int ki = __ldg(cl_ins + c0 + i);
            int r = 0;
            for (int j = 0; j < n; ++j) {
                const int kj = __ldg(cl_ins + c0 + j);
                r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
            }
            order[c0 + r] = c0 + i;
        }
        __syncthreads();
        const int before = gbase;
        for (int p0 = 0; p0 < n; p0 += CG_THREADS) {         // group heads in sorted order, scanned in chunks of one CTA
            const int p = p0 + tid;
            int flag = 0, cl = 0;
            if (p < n) {
                cl = order[c0 + p];
                flag = (p == 0 || __ldg(cl_ins + order[c0 + p - 1]) != __ldg(cl_ins + cl)) ? 1 : 0;
            }
            int inc = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(SGB_FULL_MASK, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) s_warp[wp] = inc;
            __syncthreads();
            if (wp == 0) {
                const int v = s_warp[lane];
                int iv = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(SGB_FULL_MASK, iv, o); if (lane >= o) iv += u; }
                s_warp[lane] = iv - v;
                if (lane == 31) s_warp[32] = iv;
            }
            __syncthreads();
            if (flag) {
                const int g = gbase + s_warp[wp] + inc - 1;
                off[g] = c0 + p;
                gold[g] = __ldg(cl_sem + cl);
            }
            gbase += s_warp[32];
            __syncthreads();                                 // s_warp is rewritten by the next chunk
        }
        min_groups = min(min_groups, gbase - before);
        if (tid == 0) g_off[b + 1] = gbase;
    }
    if (tid == 0) {
        off[gbase] = S;
        counts[0] = gbase;
        counts[1] = n_scenes > 0 ? min_groups : 0;
    }
}

constexpr size_t CH_FWD_SMEM = (size_t)(CH_IN * CH_WS + 2 * CH_IN + 2 * CH_HID + 2 * CH_OUT + 2 * CH_HID + 2 * CH_HID) * sizeof(float);
constexpr size_t CH_BWD_SMEM = (size_t)(CH_HID * CH_IN + 2 * CH_IN + 2 * CH_HID + 2 * CH_OUT + 4 * CH_HID) * sizeof(float);
}  // namespace

extern "C" int sgb_classifier_head_fwd(const float* feat, int G, const int* g_off, int n_scenes, const int* gold,
                                       const float* W1, const float* gamma, const float* beta, const float* W2, const float* b2,
                                       const float* mask, float drop_scale, float* hpre, float* stats, float* logits, float* loss_raw,
                                       void* stream) {
    if (G <= 0 || n_scenes <= 0 || !feat || !g_off || !gold || !W1 || !gamma || !beta || !W2 || !b2 || !hpre || !stats || !logits || !loss_raw)
        return SGB_ERR_INVALID;
    SGB_OPT_IN_SMEM(cls_head_fwd_kernel);
    { cls_head_fwd_kernel<<<n_scenes, CH_THREADS, CH_FWD_SMEM, (cudaStream_t)stream>>>(feat, g_off, gold, W1, gamma, beta, W2, b2, mask, drop_scale,
                                                                                     hpre, stats, logits, loss_raw); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_classifier_head_bwd(const float* feat, int G, const int* g_off, int n_scenes, const int* gold,
                                       const float* W1, const float* gamma, const float* beta, const float* W2,
                                       const float* mask, float drop_scale, const float* hpre, const float* stats, const float* logits,
                                       const float* grad_loss_sum, float* scratch, float* dfeat, float* dW1_part, float* dgamma_part,
                                       float* dbeta_part, float* dW2_part, float* db2_part, void* stream) {
    if (G <= 0 || n_scenes <= 0 || !feat || !g_off || !gold || !W1 || !gamma || !beta || !W2 || !hpre || !stats || !logits || !grad_loss_sum ||
        !scratch || !dfeat || !dW1_part || !dgamma_part || !dbeta_part || !dW2_part || !db2_part) return SGB_ERR_INVALID;
    SGB_OPT_IN_SMEM(cls_head_bwd_kernel);
    { cls_head_bwd_kernel<<<n_scenes, CH_THREADS, CH_BWD_SMEM, (cudaStream_t)stream>>>(feat, g_off, gold, W1, gamma, beta, W2, mask, drop_scale,
                                                                                     hpre, stats, logits, grad_loss_sum, scratch, dfeat, dW1_part,
                                                                                     dgamma_part, dbeta_part, dW2_part, db2_part); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_classifier_groups(const int* cl_ins, const int* cl_sem, const int* scene_cl_off, int n_scenes, int S,
                                     int* order, int* off, int* gold, int* g_off, int* counts, void* stream) {
    if (S <= 0 || n_scenes < 1 || !cl_ins || !cl_sem || !order || !off || !gold || !g_off || !counts) return SGB_ERR_INVALID;
    if (n_scenes > 1 && !scene_cl_off) return SGB_ERR_INVALID;
    { cls_groups_kernel<<<1, CG_THREADS, 0, (cudaStream_t)stream>>>(cl_ins, cl_sem, scene_cl_off, n_scenes, S, order, off, gold, g_off, counts); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

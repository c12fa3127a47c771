// a17: seggroup/model.py:608-655 `evaluate` on the device, two launches.
//
// The reference loops over 40 classes and over the predicted instance ids with numpy masks; every quantity it reports is a
// count, so one pass over the raw vertices with shared-memory histograms gives all of them:
//   semantic   I[c] = #(pred == c & true == c),  U[c] = #pred==c + #true==c - I[c]                       c = 1..40
//   instance   per predicted id != -1: I = #(pred == id & true == id), U = #pred==id + #true==id - I, credited to the class
//              sem_pred[first valid vertex with that id] - 1 (numpy index semantics: -1 wraps to class 40)
//   accuracy   sem / ins over valid vertices (sem_true != 0) and over the SEM_VALID / INS_VALID subsets of sem_true / ins_true
//              (the reference filters ins_true by INS_VALID *class* ids, model.py:648-651: reproduced)
// Counts are integers (exact, order independent); the four accuracies are divided in fp64 (sklearn) and rounded to fp32.
#include "common.cuh"

namespace {
constexpr int EV_THREADS = 256;
constexpr int EV_SMEM_IDS = 1024;       // instance ids below this are histogrammed in shared memory first
constexpr int EV_MAX_IDS = 65536;       // ids at or above this are not supported (status bit 64)
constexpr int EV_NCNT = 8;              // n_valid, sem_eq, ins_eq, n_semsel, semsel_eq, n_inssel, inssel_eq, overflow

struct EvWs {
    int* sem;        // [3][41]  pred, true, both
    int* cnt;        // [EV_NCNT]
    int* ins;        // [3][EV_MAX_IDS] pred, true, both
    int* first;      // [EV_MAX_IDS]  smallest valid vertex with ins_pred == id
};
__host__ __device__ inline size_t ev_ints() { return 3 * 41 + EV_NCNT + 4 * (size_t)EV_MAX_IDS; }
__host__ __device__ inline EvWs ev_layout(void* ws) {
    EvWs w;
    w.sem = (int*)ws; w.cnt = w.sem + 3 * 41; w.ins = w.cnt + EV_NCNT; w.first = w.ins + 3 * EV_MAX_IDS;
    return w;
}

__global__ void evaluate_init_kernel(int* __restrict__ ws_ints, size_t n_zero, int* __restrict__ first) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_zero) ws_ints[i] = 0;
    if (i < EV_MAX_IDS) first[i] = 0x7fffffff;
}

__global__ void __launch_bounds__(EV_THREADS)
evaluate_count_kernel(const long long* __restrict__ real_label, const int* __restrict__ sem_pred, const int* __restrict__ ins_pred,
                      int n, unsigned long long sem_valid, unsigned long long ins_valid, EvWs w) {
    __shared__ int s_sem[3][41];
    __shared__ int s_cnt[EV_NCNT];
    __shared__ int s_ins[3][EV_SMEM_IDS];
    __shared__ int s_first[EV_SMEM_IDS];
    for (int i = threadIdx.x; i < 3 * 41; i += EV_THREADS) (&s_sem[0][0])[i] = 0;
    if (threadIdx.x < EV_NCNT) s_cnt[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < EV_SMEM_IDS; i += EV_THREADS) { s_ins[0][i] = 0; s_ins[1][i] = 0; s_ins[2][i] = 0; s_first[i] = 0x7fffffff; }
    __syncthreads();
    int c_valid = 0, c_sem = 0, c_ins = 0, c_ss = 0, c_sse = 0, c_is = 0, c_ise = 0, c_ovf = 0;
    for (int i = blockIdx.x * EV_THREADS + threadIdx.x; i < n; i += gridDim.x * EV_THREADS) {
        const longlong2 tl = *reinterpret_cast<const longlong2*>(real_label + 2 * (size_t)i);
        const long long st = tl.x, it = tl.y;
        if (st == 0) continue;
        const int sp = __ldg(sem_pred + i), ip = __ldg(ins_pred + i);
        ++c_valid;
        const bool seq = (long long)sp == st, ieq = (long long)ip == it;
        c_sem += seq; c_ins += ieq;
        if (st > 0 && st < 64 && ((sem_valid >> st) & 1ull)) { ++c_ss; c_sse += seq; }
        if (it > 0 && it < 64 && ((ins_valid >> it) & 1ull)) { ++c_is; c_ise += ieq; }
        const bool p_in = sp >= 1 && sp <= 40, t_in = st >= 1 && st <= 40;
        if (p_in) atomicAdd(&s_sem[0][sp], 1);
        if (t_in) atomicAdd(&s_sem[1][(int)st], 1);
        if (p_in && seq) atomicAdd(&s_sem[2][sp], 1);
        if (ip >= 0) {                                             // predicted id (ids are >= 0; -1 = unlabeled)
            if (ip < EV_SMEM_IDS) { atomicAdd(&s_ins[0][ip], 1); atomicMin(&s_first[ip], i); if (ieq) atomicAdd(&s_ins[2][ip], 1); }
            else if (ip < EV_MAX_IDS) { atomicAdd(w.ins + ip, 1); atomicMin(w.first + ip, i); if (ieq) atomicAdd(w.ins + 2 * EV_MAX_IDS + ip, 1); }
            else ++c_ovf;
        }
        if (it >= 0) {
            if (it < EV_SMEM_IDS) atomicAdd(&s_ins[1][(int)it], 1);
            else if (it < EV_MAX_IDS) atomicAdd(w.ins + EV_MAX_IDS + (int)it, 1);
        }
    }
    c_valid = sgb_warp_sum(c_valid); c_sem = sgb_warp_sum(c_sem); c_ins = sgb_warp_sum(c_ins); c_ss = sgb_warp_sum(c_ss);
    c_sse = sgb_warp_sum(c_sse); c_is = sgb_warp_sum(c_is); c_ise = sgb_warp_sum(c_ise); c_ovf = sgb_warp_sum(c_ovf);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], c_valid); atomicAdd(&s_cnt[1], c_sem); atomicAdd(&s_cnt[2], c_ins); atomicAdd(&s_cnt[3], c_ss);
        atomicAdd(&s_cnt[4], c_sse); atomicAdd(&s_cnt[5], c_is); atomicAdd(&s_cnt[6], c_ise); atomicAdd(&s_cnt[7], c_ovf);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 41; i += EV_THREADS) { const int v = (&s_sem[0][0])[i]; if (v) atomicAdd(w.sem + i, v); }
    if (threadIdx.x < EV_NCNT && s_cnt[threadIdx.x]) atomicAdd(w.cnt + threadIdx.x, s_cnt[threadIdx.x]);
    for (int i = threadIdx.x; i < EV_SMEM_IDS; i += EV_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const int v = s_ins[a][i]; if (v) atomicAdd(w.ins + a * EV_MAX_IDS + i, v); }
        if (s_first[i] != 0x7fffffff) atomicMin(w.first + i, s_first[i]);
    }
}

// out [164] = IoU_sem [2][40], IoU_ins [2][40], acc [4]
__global__ void __launch_bounds__(1024)
evaluate_finalize_kernel(EvWs w, const int* __restrict__ sem_pred, float* __restrict__ out, int* __restrict__ status) {
    __shared__ int s_I[40], s_U[40];
    if (threadIdx.x < 40) { s_I[threadIdx.x] = 0; s_U[threadIdx.x] = 0; }
    __syncthreads();
    for (int id = threadIdx.x; id < EV_MAX_IDS; id += blockDim.x) {
        const int np_ = w.ins[id];
        if (np_ == 0) continue;
        const int both = w.ins[2 * EV_MAX_IDS + id];
        int idx = __ldg(sem_pred + w.first[id]) - 1;
        if (idx < 0) idx += 40;                                    // numpy / torch negative index
        if (idx >= 0 && idx < 40) { atomicAdd(&s_I[idx], both); atomicAdd(&s_U[idx], np_ + w.ins[EV_MAX_IDS + id] - both); }
    }
    __syncthreads();
    if (threadIdx.x < 40) {
        const int c = threadIdx.x + 1;
        const int both = w.sem[2 * 41 + c];
        out[threadIdx.x] = (float)both;
        out[40 + threadIdx.x] = (float)(w.sem[c] + w.sem[41 + c] - both);
        out[80 + threadIdx.x] = (float)s_I[threadIdx.x];
        out[120 + threadIdx.x] = (float)s_U[threadIdx.x];
    }
    if (threadIdx.x < 4) {
        const int num[4] = {w.cnt[1], w.cnt[2], w.cnt[4], w.cnt[6]};
        const int den[4] = {w.cnt[0], w.cnt[0], w.cnt[3], w.cnt[5]};
        out[160 + threadIdx.x] = (float)((double)num[threadIdx.x] / (double)den[threadIdx.x]);      // 0/0 -> NaN as np.mean([])
    }
    if (threadIdx.x == 0 && w.cnt[7] && status) atomicOr(status, 64);
}
}  // namespace

extern "C" size_t sgb_evaluate_ws_bytes(void) { return ev_ints() * sizeof(int) + 256; }

extern "C" int sgb_evaluate(const long long* real_label, const int* sem_pred, const int* ins_pred, int n,
                            const int* sem_valid_ids, int n_sem_valid, const int* ins_valid_ids, int n_ins_valid,
                            float* out, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (n <= 0 || !real_label || !sem_pred || !ins_pred || !out || !ws) return SGB_ERR_INVALID;
    if (((uintptr_t)real_label & 15) != 0) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_evaluate_ws_bytes()) return SGB_ERR_WORKSPACE;
    unsigned long long sv = 0, iv = 0;                             // host arrays: class id lists (ids 1..63)
    for (int i = 0; i < n_sem_valid; ++i) { if (sem_valid_ids[i] <= 0 || sem_valid_ids[i] > 63) return SGB_ERR_INVALID; sv |= 1ull << sem_valid_ids[i]; }
    for (int i = 0; i < n_ins_valid; ++i) { if (ins_valid_ids[i] <= 0 || ins_valid_ids[i] > 63) return SGB_ERR_INVALID; iv |= 1ull << ins_valid_ids[i]; }
    cudaStream_t st = (cudaStream_t)stream;
    EvWs w = ev_layout(ws);
    const size_t n_zero = 3 * 41 + EV_NCNT + 3 * (size_t)EV_MAX_IDS;
    { evaluate_init_kernel<<<sgb_div_up((long long)n_zero, 256), 256, 0, st>>>((int*)ws, n_zero, w.first); SGB_COUNT_LAUNCH(); }
    int grid = sgb_div_up(n, EV_THREADS * 4);
    if (grid > 148 * 2) grid = 148 * 2;
    { evaluate_count_kernel<<<grid, EV_THREADS, 0, st>>>(real_label, sem_pred, ins_pred, n, sv, iv, w); SGB_COUNT_LAUNCH(); }
    { evaluate_finalize_kernel<<<1, 1024, 0, st>>>(w, sem_pred, out, status); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// scene batch: the scenes of a batch one after the other on the same stream and workspace (stream order keeps them apart), driven
// from C: one library call per step instead of one per scene (the per-scene calls left the device idle for ~15 us each,
// profiles/r03d_timeline_8x150k_train.txt).  raw_off [n_scenes+1]: HOST array, raw-vertex range of every scene; out [n_scenes,164].
extern "C" int sgb_evaluate_scenes(const long long* real_label, const int* sem_pred, const int* ins_pred, const int* raw_off, int n_scenes,
                                   const int* sem_valid_ids, int n_sem_valid, const int* ins_valid_ids, int n_ins_valid,
                                   float* out, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (n_scenes < 1 || !raw_off || !real_label || !sem_pred || !ins_pred || !out) return SGB_ERR_INVALID;
    for (int b = 0; b < n_scenes; ++b) {
        const int lo = raw_off[b], hi = raw_off[b + 1];
        const int rc = sgb_evaluate(real_label + 2 * (size_t)lo, sem_pred + lo, ins_pred + lo, hi - lo, sem_valid_ids, n_sem_valid,
                                    ins_valid_ids, n_ins_valid, out + (size_t)b * 164, status, ws, ws_bytes, stream);
        if (rc != SGB_OK) return rc;
    }
    return SGB_OK;
}

// Microbenchmark: cycles per tcgen05.mma (kind::tf32, K = 8 per instruction, operands in the no-swizzle K-major layout of tc_common.cuh)
// as a function of M, N, the operand source of A (shared memory / TMEM) and the number of independent accumulators the instruction stream
// cycles through.  One CTA per SM, one issuing thread; all SMs run the same loop (operand fetches share nothing between SMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o gpurun_out/mma_rate tools/ubench/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../seggroup_b200/csrc/tc_common.cuh"
using namespace sgb_tc;

template <int M, int N, bool A_TMEM, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 64 * 4 + 256 * 64 * 4) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = slot;
    long long dt = 0;
    if (warp == 0) {
        const uint32_t idesc = make_idesc_tf32(M, N, false, false);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 64 * 4;
        const uint64_t da = make_desc(a0, M * 16, 128), db = make_desc(b0, N * 16, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (elect_one_sync()) {
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    const uint32_t d = tmem + (uint32_t)((i % NACC) * (NACC > 1 ? 512 / NACC : 0));
                    const uint64_t dak = da + (uint64_t)(((i % 8) * 2 * M * 16) >> 4), dbk = db + (uint64_t)(((i % 8) * 2 * N * 16) >> 4);
                    if (A_TMEM) mma_tf32_ts(d, tmem + 448u + (uint32_t)((i % 8) * 8), dbk, idesc, true);
                    else mma_tf32(d, dak, dbk, idesc, true);
                }
            }
            __syncwarp();
        }
        if (elect_one_sync()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        dt = clock64() - t0;
        if (blockIdx.x == 0 && tid == 0) out[0] = dt;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int M, int N, bool A_TMEM, int NACC>
void run(const char* label) {
    long long* out;
    cudaMalloc(&out, 8);
    const int reps = 200;
    const size_t sm = 128 * 64 * 4 + 256 * 64 * 4 + 1024;
    cudaFuncSetAttribute(rate_kernel<M, N, A_TMEM, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int w = 0; w < 2; ++w) rate_kernel<M, N, A_TMEM, NACC><<<148, 128, sm>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s M %3d N %3d  A %-4s acc %d : %7.1f cycles per tcgen05.mma  (N/2 = %d)  %s\n", label, M, N, A_TMEM ? "tmem" : "smem", NACC,
           (double)h / (reps * 24.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}


// The instruction mix of one tile of ec2_tc1_kernel: 9 x (M 128, N 64) into one accumulator, [10 x (M 128, N 80) into a second,] 24 x (M 64, N TE)
// into a third — does changing shape / accumulator between batches cost anything beyond the sum of the parts?
template <int TE, bool GRAM>
__global__ void __launch_bounds__(128, 1) mix_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 64 * 4 + 256 * 64 * 4) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = slot;
    if (warp == 0) {
        const uint32_t i1 = make_idesc_tf32(128, 64, false, false), i2 = make_idesc_tf32(64, TE, false, false), ig = make_idesc_tf32(128, 80, false, false);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 64 * 4;
        const uint64_t da1 = make_desc(a0, 128 * 16, 128), db1 = make_desc(b0, 64 * 16, 128), db2 = make_desc(b0, TE * 16, 128), dbg = make_desc(b0, 80 * 16, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (elect_one_sync()) {
#pragma unroll
                for (int i = 0; i < 9; ++i) mma_tf32(tmem + 256u, da1 + (uint64_t)(((i % 3) * 2 * 128 * 16) >> 4), db1 + (uint64_t)(((i % 3) * 2 * 64 * 16) >> 4), i1, true);
                if (GRAM) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) mma_tf32(tmem + 384u, da1 + (uint64_t)(((i % 8) * 2 * 128 * 16) >> 4), dbg + (uint64_t)(((i % 8) * 2 * 80 * 16) >> 4), ig, true);
                }
#pragma unroll
                for (int i = 0; i < 24; ++i) mma_tf32(tmem + (uint32_t)((r & 1) * 128), da1 + (uint64_t)(((i % 8) * 2 * 64 * 16) >> 4), db2 + (uint64_t)(((i % 8) * 2 * TE * 16) >> 4), i2, true);
            }
            __syncwarp();
        }
        if (elect_one_sync()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0 && tid == 0) out[0] = clock64() - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int TE, bool GRAM>
void run_mix(const char* label, double ideal) {
    long long* out;
    cudaMalloc(&out, 8);
    const int reps = 200;
    const size_t sm = 128 * 64 * 4 + 256 * 64 * 4 + 1024;
    cudaFuncSetAttribute(mix_kernel<TE, GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int w = 0; w < 2; ++w) mix_kernel<TE, GRAM><<<148, 128, sm>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-40s : %7.1f cycles per tile, sum of the parts %.0f  %s\n", label, (double)h / reps, ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    run<64, 24, false, 1>("T accumulator (backward)");
    run<64, 80, false, 1>("second layer, Gram variant");
    run<64, 120, false, 1>("second layer, inference");
    run<64, 120, true, 1>("second layer, W2 in TMEM");
    run<64, 120, false, 2>("second layer, 2 accumulators");
    run<64, 160, false, 1>("N = 160");
    run<64, 240, false, 1>("N = 240");
    run<64, 240, true, 1>("N = 240, A in TMEM");
    run<128, 64, false, 1>("first layer");
    run<128, 80, false, 1>("Gram");
    run<128, 128, false, 1>("M 128 N 128");
    run<128, 240, false, 1>("M 128 N 240");
    run<128, 240, false, 2>("M 128 N 240, 2 accumulators");
    run_mix<120, false>("tile mix, inference (9 + 24 of N 120)", 9 * 48.1 + 24 * 60.1);
    run_mix<80, true>("tile mix, training (9 + 10 + 24 of N 80)", 9 * 48.1 + 10 * 52.1 + 24 * 40.1);
    return 0;
}

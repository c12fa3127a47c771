// tcgen05 / TMEM / mbarrier building blocks (sm_100a inline PTX) shared by the tensor-core kernels.
//
// Operand convention used everywhere in this library: fp32 data is contracted as TF32 x 3
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = x with the low 13 mantissa bits cleared, x_lo = x - x_hi (exact)
// which keeps ~21 mantissa bits per product (fp32 accumulation in TMEM) — the parity bar of this path is 1e-4
// relative and discrete merge decisions hang off these features, so plain TF32 (10 bits) is not enough.
//
// Shared-memory operand tiles use the NO-SWIZZLE canonical layout of the UMMA descriptors: 16-byte chunks (4 fp32),
// core matrix = 8 rows x 16 B stored contiguously (128 B).  For a tile of R rows x C columns (C % 4 == 0, R % 8 == 0):
//     offset(r, c) = (c / 4) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + (c % 4) * 4          [bytes]
// Read as a K-major operand (rows = M or N, columns = K):  SBO = 128 (next 8-row group), LBO = R*16 (next 16-byte K chunk).
// Read as an MN-major operand (rows = K, columns = M or N): SBO = R*16 (next 4 MN elements), core matrix = 8 K-rows.
// The same physical tile therefore serves `X W^T` (K-major) and the Gram product `X^T X` (MN-major).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgb_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// byte offset of element (r, c) in a canonical no-swizzle tile with R rows
__device__ __forceinline__ uint32_t tile_off(int r, int c, int R) {
    return (uint32_t)((c >> 2) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (c & 3) * 4);
}

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1 = Blackwell, no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version
    return d;                                   // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// swizzled K-major variant: layout_type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B (the swizzle is a function of the
// absolute shared-memory address, so tile bases must be aligned to the swizzle repeat: 1024 / 512 / 256 bytes)
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return make_desc(smem_addr, lbo_bytes, sbo_bytes) | ((uint64_t)layout_type << 61);
}

// ---- instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                            // c_format = F32
         | (2u << 7) | (2u << 10)               // a_format = b_format = TF32
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// A operand read from tensor memory (rows on the lanes in the layout of an accumulator of the same M, K along the columns: 8 columns per
// kind::tf32 instruction), B from shared memory.  Constant A operands (weight tiles) are parked in TMEM once per CTA and stop costing
// shared-memory bandwidth on every instruction.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// One lane of a converged warp (cute::elect_one_sync).  Guarding the MMA issue with THIS predicate instead of `lane == 0` lets the
// compiler treat the branch as warp-uniform: with `lane == 0` every tcgen05.mma / commit was wrapped in its own ELECT + BRA.U.ANY loop
// (4 extra dependent instructions per MMA in SASS), which made the single issuing thread co-critical (ncu: 80 % busy).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, %2;\n\t"
        "@px mov.s32 %1, 1;\n\t"
        "mov.s32 %0, rx;\n\t"
        "}\n"
        : "+r"(laneid), "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

// all previously issued tcgen05.mma of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n"
        :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}

// ---- TMEM allocation (one full warp executes; the base address lands in *slot in shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: 32 lanes (this warp's quarter) x 16 consecutive columns; v[i] = D[lane][col0 + i]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// Split issue / wait (several loads in flight behind ONE wait): the registers are passed through the wait statement so that no use
// of them can be scheduled ahead of it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// M = 64 accumulators keep their 64 rows on lanes 0..15 of each TMEM quadrant.  The .16x32bx2 shape reads those 16 lanes twice: threads
// 0..15 get columns [col, col + n), threads 16..31 the SAME lanes at columns [col + SPLIT, col + SPLIT + n) — so the upper half-warp works
// on a second group of columns instead of idling.
template <int SPLIT>
__device__ __forceinline__ void tmem_ld16_halves_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x32bx2.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16], %17;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr), "n"(SPLIT) : "memory");
}
template <int SPLIT>
__device__ __forceinline__ void tmem_ld4_halves_issue(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr), "n"(SPLIT) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// compiler-level dependency of 20 loaded registers on the preceding wait (emits no instruction)
__device__ __forceinline__ void tmem_ld_pin20(uint32_t (&a)[16], uint32_t (&b)[4]) {
    asm volatile("" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                      "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]) :: "memory");
}

__device__ __forceinline__ void tmem_ld_pin16(uint32_t (&a)[16]) {
    asm volatile("" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                      "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]) :: "memory");
}

// byte offset of (row r, column c) in a K-major SWIZZLE_128B fp32 tile with R rows: 32-column blocks of R x 128 B, atoms of
// 8 rows x 128 B whose 16-byte chunks are XOR-ed with the row (tile base 1024-byte aligned)
__host__ __device__ __forceinline__ uint32_t sw128_off(int r, int c, int R) {
    return (uint32_t)((c >> 5) * (R * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 31) >> 2) ^ (r & 7)) << 4) + (c & 3) * 4);
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-byte shared-memory load that the compiler may not move across other memory operations: keeps software-pipelined loads
// (next block) AHEAD of the stores of the current block instead of being sunk next to their first use
__device__ __forceinline__ float4 lds128_ordered(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ float lds32_ordered(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// packed fp32x2 fused multiply-add (Blackwell FFMA2): acc.{x,y} = a.{x,y} * b.{x,y} + acc.{x,y}, each lane IEEE fma.rn
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
    unsigned long long d = *reinterpret_cast<unsigned long long*>(&acc);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    acc = *reinterpret_cast<float2*>(&d);
}

}  // namespace sgb_tc

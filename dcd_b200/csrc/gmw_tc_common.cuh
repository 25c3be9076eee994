// tcgen05 / TMEM / mbarrier helpers and operand layouts shared by the tensor-core MLP kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include "gmw_mlp_tile.cuh"

namespace dcd {
namespace {

// ---- operand geometry
constexpr size_t B_PART_BYTES = (size_t)CH * TE * 2;   // 32 KB: one FP16 part (hi or lo) of a [128 k][128 edge] operand
// B operand: MN-major, 128-byte swizzle.  Atom = 8 k-rows x 128 B (64 consecutive edges of one channel per row,
// 16-byte chunks XOR-swizzled by the row); [16 k-atoms][2 n-atoms] atoms of 1 KB.
constexpr uint32_t B_LBO = 1024, B_SBO = 2048;         // stride between n-atoms, between k-atoms (bytes)
constexpr uint32_t B_KSTEP = 2 * B_SBO;                // one MMA consumes K = 16 = 2 k-atoms
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One lane of a converged warp.  Issuing tcgen05.mma under this predicate (with warp-uniform operands: derive
// warp indices and tensor-memory addresses through __shfl_sync(.., 0)) lets the compiler keep descriptors in
// uniform registers; under `if (threadIdx.x == 0)` it emits a 14-instruction R2UR "waterfall" per MMA.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem], FP16 inputs, FP32 accumulate, M = 128, N = 128, K = 16
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] . B[smem] (both operands from shared memory)
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of FP32 accumulators -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one lane moves the 8 FP32 edges of one operand chunk
struct F8 {
    float v[8];
};
__device__ __forceinline__ F8 ld256(const float* p) {
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): version 1, layout_type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor: D = F32, A = B = F16, A K-major (TMEM), B MN-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 16) | ((uint32_t)(TE >> 3) << 17) | ((uint32_t)(CH >> 4) << 24);

// instruction descriptor with BOTH operands K-major (weight-gradient GEMM: K = edges)
constexpr uint32_t kIdescKK = (1u << 4) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)(CH >> 4) << 24);
// byte offset of k-step s (16 edges) inside a K-major view of an operand image
__device__ __forceinline__ uint32_t kmajor_koff(int s) { return (uint32_t)(s >> 2) * B_LBO + (uint32_t)(s & 3) * 32u; }

// One 128x128x128 layer GEMM as 3 x 8 MMAs: Wh.Xl, Wl.Xh, Wh.Xh  (issued by one thread).
// A lives in tensor memory (K = 16 halfs = 8 columns per step), B in shared memory.
__device__ __forceinline__ void issue_layer_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint32_t a = (term == 1) ? a_lo : a_hi;
        const uint32_t b = (term == 0) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < CH / 16; ++ks) {
            umma_f16_ts(tmem_d, a + ks * 8, smem_desc(b + ks * B_KSTEP, B_LBO, B_SBO), kIdesc, acc);
            acc = 1;
        }
    }
}

// 8 consecutive edges (edge block `eblk`) of one input channel -> FP16 hi/lo, one 16-byte store each into the
// swizzled MN-major operand.  Conflict-free both for 8 lanes = 8 channels of one edge block (epilogue) and for
// 8 lanes = {4 edge blocks} x {channels c, c+4} (coalesced source loads).
__device__ __forceinline__ void store_b8(unsigned char* b_hi, unsigned char* b_lo, int ch, int eblk, const float (&v)[8]) {
    __half2 h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half h0 = __float2half_rn(v[2 * q]), h1 = __float2half_rn(v[2 * q + 1]);
        h[q] = __halves2half2(h0, h1);
        l[q] = __halves2half2(__float2half_rn(v[2 * q] - __half2float(h0)), __float2half_rn(v[2 * q + 1] - __half2float(h1)));
    }
    const uint32_t krow = (uint32_t)ch & 7u;
    const uint32_t off = (uint32_t)(ch >> 3) * B_SBO + (uint32_t)(eblk >> 3) * B_LBO + krow * 128u + ((((uint32_t)eblk & 7u) ^ krow) << 4);
    *reinterpret_cast<uint4*>(b_hi + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(b_lo + off) = *reinterpret_cast<const uint4*>(l);
}

// One output-channel row of a weight matrix -> FP16 hi/lo pairs -> tensor memory (A operand, K-major:
// lane = out channel, 32-bit column c holds input channels 2c (low half) and 2c+1).
__device__ __forceinline__ void load_weight_row_to_tmem(const float* __restrict__ Wt, float scale, int row,
                                                        uint32_t t_hi, uint32_t t_lo, int half) {
    {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int k = half * 64 + 2 * c;
            const float w0 = __ldg(Wt + (int64_t)k * CH + row) * scale;
            const float w1 = __ldg(Wt + (int64_t)(k + 1) * CH + row) * scale;
            const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
            const __half l0 = __float2half_rn(w0 - __half2float(h0)), l1 = __float2half_rn(w1 - __half2float(h1));
            hi[c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            lo[c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        tmem_st32(t_hi + half * 32, hi);
        tmem_st32(t_lo + half * 32, lo);
    }
    tc_wait_st();
}


// Same for the TRANSPOSED matrix (A[m = input channel][k = output channel] = W[k][m]): the A operand of the
// data-gradient GEMMs.  Wt is the blob's [in][out] matrix, so row m is contiguous.
__device__ __forceinline__ void load_weight_row_to_tmem_T(const float* __restrict__ Wt, float scale, int row,
                                                          uint32_t t_hi, uint32_t t_lo, int half) {
    uint32_t hi[32], lo[32];
    const float2* src = reinterpret_cast<const float2*>(Wt + (int64_t)row * CH + half * 64);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float2 w = __ldg(src + c);
        const float w0 = w.x * scale, w1 = w.y * scale;
        const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
        const __half l0 = __float2half_rn(w0 - __half2float(h0)), l1 = __float2half_rn(w1 - __half2float(h1));
        hi[c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        lo[c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    tmem_st32(t_hi + half * 32, hi);
    tmem_st32(t_lo + half * 32, lo);
    tc_wait_st();
}

}  // namespace
}  // namespace dcd

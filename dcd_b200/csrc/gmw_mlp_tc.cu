// GMW edge-feature MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same segment structure as the CUDA-core version (FIRST / B / CA between context-norm barriers), but
// the 128x128x128 layer GEMMs run as tcgen05.mma with FP32 accumulation in tensor memory:
//   D[out-channel (TMEM lane) x edge (TMEM column)] = W[out x in] . X[in x edge]
// FP32 fidelity comes from a two-term FP16 split of BOTH operands (x = hi + lo, 11 + 11 significant
// bits) and three MMAs per product:  D = Wh.Xl + Wl.Xh + Wh.Xh   (the dropped Wl.Xl term is 2^-22).
// Weights are pre-scaled per matrix by a power of two so that hi/lo stay in FP16's normal range; the
// epilogue multiplies by the exact inverse.  No tensor maps are needed:
//   * A (weights) is pre-split and pre-tiled in global memory in the canonical K-major no-swizzle UMMA
//     layout by a prep kernel and lands in shared memory with cp.async.bulk; it stays resident while a
//     persistent CTA loops over its tiles (one CTA per SM: 74 per net);
//   * B (activations) is written straight into the canonical MN-major no-swizzle layout by the threads
//     that produce it (thread = input channel, 16-byte stores of 8 edges: conflict free), both for the
//     tile source (context norm + ReLU + residual fused into the load) and for the chained preconv
//     output, which never leaves the SM;
//   * the epilogue reads D with tcgen05.ld (32 lanes x 32 columns per warp): a thread owns one output
//     channel, so bias, context-norm statistics and the split for the next GEMM need no shuffles.
#include <cuda_fp16.h>
#include "gmw_mlp_tile.cuh"

namespace dcd {

struct MlpArgs {
    const float* kpts2d;
    const float* kpts3d;
    const float* params[2];
    float* ws;
    WsLayout L;
};

namespace {

enum { MODE_FIRST = 0, MODE_B = 1, MODE_CA = 2 };

// ---- tensor-core weight image: [net][blk][which]{ hi[128*128] half, lo[128*128] half }, + unscale floats
constexpr int W_HALFS = CH * CH;                       // 16384 halfs = 32 KB per part
constexpr size_t W_PART_BYTES = (size_t)W_HALFS * 2;   // 32768
constexpr uint32_t A_LBO = 128, A_SBO = 2048;          // K-major: K-core stride, M-block stride (bytes)
constexpr uint32_t B_SBO = 128, B_LBO = 2048;          // MN-major: N-block stride, K-block stride (bytes)
constexpr uint32_t A_KSTEP = 2 * A_LBO;                // 16 k-elements = 2 K-cores
constexpr uint32_t B_KSTEP = 2 * B_LBO;
constexpr int TC_THREADS = 256;
constexpr int TMEM_COLS = 128;

// smem carve (bytes)
constexpr size_t SM_A = 0;                                   // 4 x 32 KB: W(a) hi, lo, W(b) hi, lo
constexpr size_t SM_B = SM_A + 4 * W_PART_BYTES;             // 2 x 32 KB: X hi, lo
constexpr size_t SM_STAT = SM_B + 2 * W_PART_BYTES;          // 128 float2
constexpr size_t SM_F = SM_STAT + CH * sizeof(float2);       // 128 x 8 floats edge features (FIRST)
constexpr size_t SM_HALF = SM_F + TE * 8 * sizeof(float);    // 128 x float4 half-tile statistics exchange
constexpr size_t SM_BAR = SM_HALF + CH * sizeof(float4);     // mbarriers + tmem pointer
constexpr size_t kTcSmem = SM_BAR + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
// D[tmem] (+)= A[smem] . B[smem], FP16 inputs, FP32 accumulate, M = 128, N = 128, K = 16
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: version 1, layout_type 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = F16, A K-major, B MN-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 16) | ((uint32_t)(TE >> 3) << 17) | ((uint32_t)(CH >> 4) << 24);

// One 128x128x128 layer GEMM as 3 x 8 MMAs: Wh.Xl, Wl.Xh, Wh.Xh  (issued by one thread)
__device__ __forceinline__ void issue_layer_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint32_t a = (term == 1) ? a_lo : a_hi;
        const uint32_t b = (term == 0) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < CH / 16; ++ks) {
            umma_f16(tmem_d, smem_desc(a + ks * A_KSTEP, A_LBO, A_SBO), smem_desc(b + ks * B_KSTEP, B_LBO, B_SBO), kIdesc, acc);
            acc = 1;
        }
    }
}

// 8 consecutive edges of one input channel -> FP16 hi/lo, one 16-byte store each (MN-major canonical layout)
__device__ __forceinline__ void store_b8(unsigned char* b_hi, unsigned char* b_lo, int ch, int eblk, const float (&v)[8]) {
    __half2 h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half h0 = __float2half_rn(v[2 * q]), h1 = __float2half_rn(v[2 * q + 1]);
        h[q] = __halves2half2(h0, h1);
        l[q] = __halves2half2(__float2half_rn(v[2 * q] - __half2float(h0)), __float2half_rn(v[2 * q + 1] - __half2float(h1)));
    }
    const uint32_t off = (uint32_t)eblk * B_SBO + (uint32_t)(ch >> 3) * B_LBO + (uint32_t)(ch & 7) * 16;
    *reinterpret_cast<uint4*>(b_hi + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(b_lo + off) = *reinterpret_cast<const uint4*>(l);
}

__global__ void __launch_bounds__(256) tc_prep_weights_kernel(const float* __restrict__ p4, const float* __restrict__ p6,
                                                              int depth, __half* __restrict__ wimg, float* __restrict__ unscale) {
    const int m = blockIdx.x;                               // (net, blk, which)
    const int which = m % 3, blk = (m / 3) % depth, net = m / (3 * depth);
    const int cin = net == 0 ? 4 : 6;
    const float* Wt = (net == 0 ? p4 : p6) + blob_w(cin, blk, which);     // [in k][out m]
    __shared__ float red[8];
    float mx = 0.f;
    for (int i = threadIdx.x; i < CH * CH; i += 256) mx = fmaxf(mx, fabsf(Wt[i]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.f;
    for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[w]);
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);                           // mx = f * 2^e, f in [0.5, 1)
    const float scale = ldexpf(1.f, 10 - e);                // scaled max in [512, 1024)
    if (threadIdx.x == 0) unscale[m] = ldexpf(1.f, e - 10);
    __half* hi = wimg + (size_t)m * 2 * W_HALFS;
    __half* lo = hi + W_HALFS;
    for (int i = threadIdx.x; i < CH * CH; i += 256) {
        const int k = i >> 7, mo = i & 127;                 // coalesced read of Wt[k][mo]
        const float w = Wt[i] * scale;
        const __half h = __float2half_rn(w);
        const int off = ((mo >> 3) * (int)A_SBO + (k >> 3) * (int)A_LBO + (mo & 7) * 16 + (k & 7) * 2) >> 1;
        hi[off] = h;
        lo[off] = __float2half_rn(w - __half2float(h));
    }
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(MlpArgs a, int blk, const __half* __restrict__ wimg, const float* __restrict__ unscale) {
    const WsLayout& L = a.L;
    const int net = blockIdx.y;
    const int cin = net == 0 ? 4 : 6;
    const float* __restrict__ prm = a.params[net];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = L.E, EP = L.EP, T = L.T;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* A0_hi = smem + SM_A;
    unsigned char* A0_lo = A0_hi + W_PART_BYTES;
    unsigned char* A1_hi = A0_lo + W_PART_BYTES;
    unsigned char* A1_lo = A1_hi + W_PART_BYTES;
    unsigned char* B_hi = smem + SM_B;
    unsigned char* B_lo = B_hi + W_PART_BYTES;
    float2* stat_s = reinterpret_cast<float2*>(smem + SM_STAT);
    float* f_s = reinterpret_cast<float*>(smem + SM_F);
    float4* half_s = reinterpret_cast<float4*>(smem + SM_HALF);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_w + 2);

    // ---- one-time setup: barriers, TMEM, resident weights
    const int w0 = (MODE == MODE_B) ? 2 : 0;                // first matrix of this segment (conv2 | preconv)
    const int mat0 = (net * L.depth + blk) * 3 + w0;
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;
    if (tid == 0) {
        const uint32_t bytes = (MODE == MODE_B ? 2u : 4u) * (uint32_t)W_PART_BYTES;
        mbar_expect_tx(bar_w, bytes);
        bulk_g2s(A0_hi, wimg + (size_t)mat0 * 2 * W_HALFS, 2 * W_PART_BYTES, bar_w);
        if (MODE != MODE_B) bulk_g2s(A1_hi, wimg + (size_t)(mat0 + 1) * 2 * W_HALFS, 2 * W_PART_BYTES, bar_w);
    }
    const float un0 = __ldg(unscale + mat0);
    const float un1 = (MODE != MODE_B) ? __ldg(unscale + mat0 + 1) : 0.f;
    // epilogue ownership: thread = output channel `ch`, column half `hsel` (64 edges)
    const int ch = 32 * (warp & 3) + lane;
    const int hsel = warp >> 2;
    const uint32_t t_lane = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(hsel * 64);
    const float bias0 = __ldg(prm + blob_b(cin, blk, w0) + ch);
    const float bias1 = (MODE != MODE_B) ? __ldg(prm + blob_b(cin, blk, 1) + ch) : 0.f;

    uint32_t mma_phase = 0;
    bool weights_ready = false;
    int64_t stat_obj = -1;
    const int64_t ntiles = L.N * (int64_t)T;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t obj = t / T;
        const int tile = (int)(t - obj * T);
        const int64_t obj_off = obj * (int64_t)CH * EP;
        const int valid = min(TE, E - tile * TE);

        // ---- source tile -> B operand (hi/lo)
        if (MODE == MODE_FIRST) {
            if (tid < TE) {
                const int e = tile * TE + tid;
                int i, j;
                decode_edge(e < E ? e : E - 1, L.n, i, j);
                float* f = f_s + tid * 8;
                if (net == 0) {
                    const float2 pi = __ldg(reinterpret_cast<const float2*>(a.kpts2d + (obj * L.n + i) * 2));
                    const float2 pj = __ldg(reinterpret_cast<const float2*>(a.kpts2d + (obj * L.n + j) * 2));
                    f[0] = pi.x; f[1] = pi.y; f[2] = pj.x; f[3] = pj.y; f[4] = 0.f; f[5] = 0.f;
                } else {
                    const float* pi = a.kpts3d + (obj * L.n + i) * 3;
                    const float* pj = a.kpts3d + (obj * L.n + j) * 3;
                    f[0] = __ldg(pi); f[1] = __ldg(pi + 1); f[2] = __ldg(pi + 2);
                    f[3] = __ldg(pj); f[4] = __ldg(pj + 1); f[5] = __ldg(pj + 2);
                }
            }
            __syncthreads();
        } else if (stat_obj != obj) {
            const int pb = (MODE == MODE_B) ? blk : blk - 1;
            const int which = (MODE == MODE_B) ? 0 : 1;
            __syncthreads();                                  // previous tile's readers of stat_s are done
            if (tid < CH)
                stat_s[tid] = merge_cn_stats(stat_ptr(a.ws, L, net, pb, which) + obj * (int64_t)T * CH, tid, T, E);
            __syncthreads();
            stat_obj = obj;
        }
        {
            const int pb = (MODE == MODE_B) ? blk : blk - 1;
            const float* Y = (MODE == MODE_FIRST) ? nullptr : act_ptr(a.ws, L, net, pb, MODE == MODE_B ? SLOT_Y1 : SLOT_Y2) + obj_off;
            const float* Xp = (MODE == MODE_CA) ? act_ptr(a.ws, L, net, blk - 1, SLOT_X) + obj_off : nullptr;
            float* Xn = (MODE == MODE_B) ? nullptr : act_ptr(a.ws, L, net, MODE == MODE_FIRST ? 0 : blk, SLOT_X) + obj_off;
            const int c_sub = lane & 7, j_sub = lane >> 3;
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int item = warp * 8 + it;              // 64 items: 16 channel groups x 4 quads of edge blocks
                const int c = (item >> 2) * 8 + c_sub;
                const int eblk = (item & 3) * 4 + j_sub;     // block of 8 edges
                const int e0 = tile * TE + eblk * 8;
                float v[8];
                if (MODE == MODE_FIRST) {
                    const float b = __ldg(prm + blob_in_b(cin) + c);
                    float wq[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) wq[q] = (q < cin) ? __ldg(prm + blob_in_w() + q * CH + c) : 0.f;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float* f = f_s + (eblk * 8 + q) * 8;
                        float x = b;
#pragma unroll
                        for (int r = 0; r < 6; ++r) x = fmaf(wq[r], f[r], x);
                        v[q] = x;
                    }
                } else {
                    const float2 st = stat_s[c];
                    const float4 y0 = *reinterpret_cast<const float4*>(Y + (int64_t)c * EP + e0);
                    const float4 y1 = *reinterpret_cast<const float4*>(Y + (int64_t)c * EP + e0 + 4);
                    const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = (yv[q] - st.x) * st.y;
                    if (MODE == MODE_CA) {
                        const float4 x0 = *reinterpret_cast<const float4*>(Xp + (int64_t)c * EP + e0);
                        const float4 x1 = *reinterpret_cast<const float4*>(Xp + (int64_t)c * EP + e0 + 4);
                        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f) + xv[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (eblk * 8 + q >= valid) v[q] = 0.f;
                if (MODE != MODE_B) {
                    *reinterpret_cast<float4*>(Xn + (int64_t)c * EP + e0) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(Xn + (int64_t)c * EP + e0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                store_b8(B_hi, B_lo, c, eblk, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (!weights_ready) {
            mbar_wait(bar_w, 0);
            weights_ready = true;
        }
        // ---- GEMM 1
        if (tid == 0) {
            tc_fence_after();
            issue_layer_gemm(tmem_d, smem_u32(A0_hi), smem_u32(A0_lo), smem_u32(B_hi), smem_u32(B_lo));
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        float vals[64];
        if (MODE != MODE_B) {
            // preconv output: + bias, kept on chip as the operand of conv1
            float* Pg = L.save ? act_ptr(a.ws, L, net, blk, SLOT_P) + obj_off + (int64_t)ch * EP + tile * TE + hsel * 64 : nullptr;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float v[32];
                tmem_ld32(t_lane + part * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], un0, bias0);
                if (Pg != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(Pg + part * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) {
                    const float w8[8] = {v[bq * 8], v[bq * 8 + 1], v[bq * 8 + 2], v[bq * 8 + 3],
                                         v[bq * 8 + 4], v[bq * 8 + 5], v[bq * 8 + 6], v[bq * 8 + 7]};
                    store_b8(B_hi, B_lo, ch, hsel * 8 + part * 4 + bq, w8);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- GEMM 2 (conv1)
            if (tid == 0) {
                tc_fence_after();
                issue_layer_gemm(tmem_d, smem_u32(A1_hi), smem_u32(A1_lo), smem_u32(B_hi), smem_u32(B_lo));
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
        }
        // ---- final epilogue of the segment: + bias, store, tile statistics
        {
            const float un = (MODE == MODE_B) ? un0 : un1;
            const float bias = (MODE == MODE_B) ? bias0 : bias1;
            float* Yo = act_ptr(a.ws, L, net, blk, MODE == MODE_B ? SLOT_Y2 : SLOT_Y1) + obj_off + (int64_t)ch * EP + tile * TE + hsel * 64;
            const int nv = max(0, min(64, valid - hsel * 64));
            float s = 0.f;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float v[32];
                tmem_ld32(t_lane + part * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = fmaf(v[i], un, bias);
                    vals[part * 32 + i] = v[i];
                    if (part * 32 + i < nv) s += v[i];
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(Yo + part * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            const float mean_h = nv > 0 ? s / (float)nv : 0.f;
            float m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i)
                if (i < nv) {
                    const float d = vals[i] - mean_h;
                    m2 = fmaf(d, d, m2);
                }
            if (hsel == 1) half_s[ch] = make_float4(mean_h, m2, (float)nv, 0.f);
            tc_fence_before();
            __syncthreads();                                  // also orders the TMEM reads before the next tile's MMAs
            if (hsel == 0) {
                const float4 o = half_s[ch];
                float mean = mean_h, M2 = m2;
                if (o.z > 0.f) {                              // Chan merge of the two 64-edge halves
                    const float na = (float)nv, nb = o.z, tot = na + nb;
                    const float delta = o.x - mean_h;
                    mean = mean_h + delta * (nb / tot);
                    M2 = m2 + o.y + delta * delta * (na * nb / tot);
                }
                stat_ptr(a.ws, L, net, blk, MODE == MODE_B ? 1 : 0)[(obj * T + tile) * (int64_t)CH + ch] = make_float2(mean, M2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// Final features of both nets -> reg_weights (same arithmetic as the CUDA-core path's kernel).
__global__ void __launch_bounds__(256) gmw_edge_weight_kernel(MlpArgs a, float* __restrict__ reg_w,
                                                              float* __restrict__ feat4, float* __restrict__ feat6) {
    const WsLayout& L = a.L;
    const int E = L.E, EP = L.EP, last = L.depth - 1;
    const int nb = (E + 255) / 256;
    const int64_t obj = blockIdx.x / nb;
    const int e = (blockIdx.x % nb) * 256 + threadIdx.x;
    __shared__ float2 stat_s[2][CH];
    {
        const int net = threadIdx.x >> 7, c = threadIdx.x & 127;
        stat_s[net][c] = merge_cn_stats(stat_ptr(a.ws, L, net, last, 1) + obj * (int64_t)L.T * CH, c, L.T, E);
    }
    __syncthreads();
    if (e >= E) return;
    const int64_t off = obj * (int64_t)CH * EP + e;
    const float* Y4 = act_ptr(a.ws, L, 0, last, SLOT_Y2) + off;
    const float* X4 = act_ptr(a.ws, L, 0, last, SLOT_X) + off;
    const float* Y6 = act_ptr(a.ws, L, 1, last, SLOT_Y2) + off;
    const float* X6 = act_ptr(a.ws, L, 1, last, SLOT_X) + off;
    float n4 = 0.f, n6 = 0.f;
#pragma unroll 4
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = fmaxf((Y4[(int64_t)c * EP] - s4.x) * s4.y, 0.f) + X4[(int64_t)c * EP];
        const float x6 = fmaxf((Y6[(int64_t)c * EP] - s6.x) * s6.y, 0.f) + X6[(int64_t)c * EP];
        n4 = fmaf(x4, x4, n4);
        n6 = fmaf(x6, x6, n6);
        if (feat4 != nullptr) feat4[(obj * CH + c) * (int64_t)E + e] = x4;
        if (feat6 != nullptr) feat6[(obj * CH + c) * (int64_t)E + e] = x6;
    }
    n4 = fmaxf(sqrtf(n4), 1e-12f);     // F.normalize: x / max(||x||, eps)   (model.py:176-177)
    n6 = fmaxf(sqrtf(n6), 1e-12f);
    float a2 = 0.f, c2 = 0.f, ac = 0.f;
#pragma unroll 4
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = fmaxf((Y4[(int64_t)c * EP] - s4.x) * s4.y, 0.f) + X4[(int64_t)c * EP];
        const float x6 = fmaxf((Y6[(int64_t)c * EP] - s6.x) * s6.y, 0.f) + X6[(int64_t)c * EP];
        const float av = __fdiv_rn(x4, n4), cv = __fdiv_rn(x6, n6);
        a2 = fmaf(av, av, a2);
        c2 = fmaf(cv, cv, c2);
        ac = fmaf(av, cv, ac);
    }
    // pairwiseL2Dist diagonal (model.py:28-35): ((|c|^2 - 2 a.c) + |a|^2).clamp_min(1e-30).sqrt(); graph_extract: 1/M
    const float s = __fadd_rn(__fadd_rn(c2, -2.f * ac), a2);
    reg_w[obj * (int64_t)E + e] = __fdiv_rn(1.f, sqrtf(fmaxf(s, 1e-30f)));
}

}  // namespace

// bytes appended to the MLP workspace for the tensor-core weight image
size_t tc_weight_image_bytes(int depth) {
    const size_t mats = (size_t)2 * depth * 3;
    return mats * 2 * W_PART_BYTES + ((mats * sizeof(float) + 255) / 256) * 256;
}

int launch_gmw_weights_fwd(const float* kpts2d, const float* kpts3d, const float* params4, const float* params6,
                           int64_t N, int n, int depth, int save, float* reg_w, float* feat4, float* feat6,
                           float* ws, cudaStream_t st) {
    MlpArgs a;
    a.kpts2d = kpts2d; a.kpts3d = kpts3d;
    a.params[0] = params4; a.params[1] = params6;
    a.ws = ws;
    a.L = make_layout(N, n, depth, save);
    if ((int64_t)a.L.T * N > 0x7fffffffLL) return DCD_E_UNSUPPORTED;
    // weight image lives right after the layout's own area (256-byte aligned)
    unsigned char* img = reinterpret_cast<unsigned char*>(ws) + (((size_t)a.L.total * sizeof(float) + 255) / 256) * 256;
    __half* wimg = reinterpret_cast<__half*>(img);
    float* unscale = reinterpret_cast<float*>(img + (size_t)2 * depth * 3 * 2 * W_PART_BYTES);
    tc_prep_weights_kernel<<<2 * depth * 3, 256, 0, st>>>(params4, params6, depth, wimg, unscale);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    const int64_t ntiles = a.L.T * N;
    const int per_net = max(1, device_sm_count() / 2);
    const dim3 grid((unsigned)(ntiles < per_net ? ntiles : per_net), 2);
    mlp_tc_kernel<MODE_FIRST><<<grid, TC_THREADS, kTcSmem, st>>>(a, 0, wimg, unscale);
    for (int blk = 0; blk < depth; ++blk) {
        mlp_tc_kernel<MODE_B><<<grid, TC_THREADS, kTcSmem, st>>>(a, blk, wimg, unscale);
        if (blk + 1 < depth) mlp_tc_kernel<MODE_CA><<<grid, TC_THREADS, kTcSmem, st>>>(a, blk + 1, wimg, unscale);
    }
    DCD_CHECK_LAUNCH();
    const unsigned g2 = (unsigned)(((a.L.E + 255) / 256) * N);
    gmw_edge_weight_kernel<<<g2, 256, 0, st>>>(a, reg_w, feat4, feat6);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

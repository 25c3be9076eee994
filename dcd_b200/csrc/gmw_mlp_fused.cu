// GMW edge-feature MLP forward, inference form: ONE kernel for the whole net (conv_in + 12 blocks; each block's
// preconv and conv1 folded into one layer: 1 + 24 layers instead of 1 + 36), the object's activations never leave
// the chip (sm_100a: tcgen05 + TMEM, persistent co-resident CTA groups).
//
// The layer-wise kernels (gmw_mlp_tc.cu) are bound by HBM: every context norm needs the statistics of the
// whole object, so each of the 24 normalised layers writes its output and reads it back (198 MB / object).
// Here a group of 8 CTAs owns one (object, net): CTA r holds the edge slice [r*ES, (r+1)*ES) of ALL 128
// channels on chip for the whole network, and the only thing the CTAs exchange per context norm is the
// per-channel (mean, M2) partial of their slice (1 KB per CTA, through L2 with self-validating words: no
// fences, no barriers).  Groups are plain consecutive CTAs of a cooperative launch (all CTAs resident, one per
// SM): 18 groups use 144 of the 148 SMs, where hardware clusters of 8 can only be placed on 120.
// Per CTA (ES <= 336 edges, E <= 2688, i.e. up to the reference's n = 73 keypoints):
//   * residual stream X        : FP32 in shared memory, [ES/4][128 ch] float4            (172 KB)
//   * layer output (P, Y1, Y2) : FP32 accumulators in tensor memory, columns [0, ES)     (336 of 512 columns);
//                                every GEMM overwrites its own input sub-tile in place
//   * weights                  : FP16 hi/lo pairs as the A operand in tensor memory, columns [384, 512),
//                                reloaded per layer from an L2-resident pre-split image (64 KB / layer)
//   * B operand                : two 24 KB buffers (hi + lo of a 48-edge sub-tile, MN-major, no swizzle),
//                                written by the threads that produce the values (thread = channel)
// Warp roles: 12 converter warps (thread = channel; accumulators -> context norm / ReLU / residual -> FP16
// hi/lo operand; statistics of the layer output), 4 service warps (weight image -> tensor memory; the first one
// issues the MMAs from an elected lane).  Hand-offs are mbarriers; there is no CTA-wide barrier per sub-tile.
// Arithmetic is the one of the layer-wise kernels (FP16x3 split, power-of-two weight scaling, FP32 statistics) up to
// the folded layer (Wf = W1.Wp formed in FP64, rounded once) and the order in which the statistics partials are merged.
// HBM traffic: keypoints in, final features out (2.75 MB / object instead of 198 MB).
#include <type_traits>
#include "gmw_tc_common.cuh"

#ifndef DCD_FUSED_SLEEP_NS
#define DCD_FUSED_SLEEP_NS 0        // back-off of the weight loaders' and statistics warps' barrier waits (0: spin)
#endif

namespace dcd {
// Optional in-kernel timeline (build with -DDCD_FUSED_TRACE, see profiles/trace_fused.py): lane 0 of converter warp 0
// and of the MMA warp of CTA 0 record (tag, clock64) pairs; read back with dcd_debug_fused_trace().
#ifdef DCD_FUSED_TRACE
__device__ long long g_trace[12288];
__device__ int g_trace_n[3];
#ifndef DCD_FUSED_TRACE_CTA
#define DCD_FUSED_TRACE_CTA 0
#endif
#ifndef DCD_FUSED_TRACE_WARP
#define DCD_FUSED_TRACE_WARP 0
#endif
#define TR_DECL const bool trace_on = blockIdx.x == DCD_FUSED_TRACE_CTA && (warp == DCD_FUSED_TRACE_WARP || warp == FCONV_WARPS || warp == FCONV_WARPS + 1); int tr_k = 0;
#define TR(slot, tag)                                                       \
    do {                                                                    \
        if (trace_on && lane == 0 && tr_k < 2048) {                         \
            g_trace[(slot) * 4096 + 2 * tr_k] = (tag);                      \
            g_trace[(slot) * 4096 + 2 * tr_k + 1] = clock64();              \
            ++tr_k;                                                         \
            g_trace_n[slot] = tr_k;                                         \
        }                                                                   \
    } while (0)
#else
#define TR_DECL
#define TR(slot, tag) do { } while (0)
#endif
namespace {

constexpr int FCS = 16;             // CTAs per group = edge slices per object; a group works on TWO objects at a time
constexpr int FCONV_WARPS = 12;     // converter warps: 4 TMEM lane quarters x 3 units (16 edges) of a 48-edge sub-tile
constexpr int FCONV_THREADS = 32 * FCONV_WARPS;
constexpr int FTHREADS = FCONV_THREADS + 128;   // + 4 service warps: weight loads (all 4), MMA issue (the first)
constexpr int FSUB = 48;            // edges per MMA sub-tile
constexpr int FES_MAX = 168;        // edges per CTA and object
constexpr uint32_t FT_SLOT = 192;   // tensor-memory columns: D of object half 0 = [0, 176), of half 1 = [192, 368),
constexpr uint32_t FT_W = 384;      //                        weights hi = [384, 448), lo = [448, 512)
constexpr uint32_t FB_SBO = 128;    // B operand: bytes between 8-edge blocks of one 8-channel block
constexpr uint32_t FB_LBO = 768;    //            bytes between 8-channel blocks (6 edge blocks)
constexpr uint32_t FB_PART = 16 * FB_LBO;   // 12 KB: hi or lo part of one sub-tile

constexpr size_t SMF_X = 0;
constexpr size_t SMF_B = SMF_X + (size_t)2 * FES_MAX * CH * sizeof(float);      // [2 buffers]{hi, lo}
constexpr size_t SMF_PART = SMF_B + 4 * FB_PART;                                // [2][3][128] float2 (mean, M2) + count tables + [2][128] statistics
constexpr size_t SMF_BAR = SMF_PART + (2 * 3 * CH + 32 + 2 * CH) * sizeof(float2);
constexpr size_t kFusedSmem = SMF_BAR + 128;
static_assert(kFusedSmem <= 232448, "shared memory budget");
static_assert(2 * 6 * FES_MAX * sizeof(float) <= 4 * FB_PART, "edge features are staged in the operand buffers");
static_assert(2 * FES_MAX <= FCONV_THREADS, "one converter thread per staged edge");
// mbarriers (8 bytes each, at SMF_BAR)
enum { BAR_FULL0 = 0, BAR_FULL1 = 1, BAR_DONE0 = 2, BAR_DONE1 = 3, BAR_PDONE = 4, BAR_WREADY = 5, BAR_COUNT = 6 };
// hardware named barriers (a blocked warp takes no issue slots): 0 = CTA, 1 = converters, 2..4 = the units of the epilogue,
// then the hand-offs between the converters (12 warps) and the statistics warps (3 warps): one side arrives, the other waits
enum { NB_STAT0 = 5, NB_STAT1 = 6, NB_PART = 7 };
constexpr int FSTAT_THREADS = 96;
constexpr int NB_THREADS = FCONV_THREADS + FSTAT_THREADS;

// Statistics exchange between the 8 CTAs of a group through global memory (L2): every published 8-byte word carries
// its own "ready" flag in the sign bit of its second float (an M2 is never negative), so there are no fences, no
// barriers and no ordering requirements: a reader re-reads a word until the flag matches the expected use count.
__device__ __forceinline__ void st_flagged(float2* dst, float a, float b, uint32_t flag) {
    const uint32_t bits = (__float_as_uint(b) & 0x7fffffffu) | (flag << 31);
    asm volatile("st.volatile.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(a), "f"(__uint_as_float(bits)) : "memory");
}
__device__ __forceinline__ float2 ld_volatile_f2(const float2* p) {
    float2 v;
    asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// weight image: read-only, re-read by every CTA for every layer -> keep it in L2 against the streamed outputs
__device__ __forceinline__ void ld_weights8(const uint32_t* p, uint32_t* r) {
    asm volatile("ld.global.nc.L2::evict_last.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait of a warp that has nothing else to do for a whole layer (weight loaders): back off instead of spinning in the issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#if DCD_FUSED_SLEEP_NS > 0
        if (!done) __nanosleep(DCD_FUSED_SLEEP_NS);
#endif
    } while (!done);
}
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void conv_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FCONV_THREADS) : "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load split into issue and wait, so that its latency can be covered by independent work
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// packed FP32 pair arithmetic (sm_100 FFMA2 / FADD2): one issue slot for two elements
__device__ __forceinline__ void ffma2_bc(float& d0, float& d1, float a0, float a1, float b, float c) {   // d = a * b + c
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\tmov.b64 c, {%5, %5};\n\t"
        "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fsq2_acc(float& s0, float& s1, float d0, float d1) {                      // s += d * d
    asm("{\n\t.reg .b64 d, s;\n\tmov.b64 d, {%2, %3};\n\tmov.b64 s, {%0, %1};\n\t"
        "fma.rn.f32x2 s, d, d, s;\n\tmov.b64 {%0, %1}, s;\n\t}"
        : "+f"(s0), "+f"(s1) : "f"(d0), "f"(d1));
}
__device__ __forceinline__ void fadd2_acc(float& s0, float& s1, float d0, float d1) {                     // s += d
    asm("{\n\t.reg .b64 d, s;\n\tmov.b64 d, {%2, %3};\n\tmov.b64 s, {%0, %1};\n\t"
        "add.rn.f32x2 s, s, d;\n\tmov.b64 {%0, %1}, s;\n\t}"
        : "+f"(s0), "+f"(s1) : "f"(d0), "f"(d1));
}

// no-swizzle (interleaved) shared-memory matrix descriptor
__device__ __forceinline__ uint64_t smem_desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// (v0, v1) -> packed FP16 pair hp = rn(v) and lp = rn(hp - v) = MINUS the low part (the MMA that consumes the
// low parts negates B): F2FP + 2 x FHADD (mixed f16/f32 add) + F2FP for two values.
__device__ __forceinline__ void split2_neg(float v0, float v1, uint32_t& hp, uint32_t& lp) {
    asm("{\n\t.reg .f16 h0, h1;\n\t.reg .f32 l0, l1;\n\t"
        "cvt.rn.f16x2.f32 %0, %3, %2;\n\t"
        "mov.b32 {h0, h1}, %0;\n\t"
        "sub.rn.f32.f16 l0, h0, %2;\n\t"
        "sub.rn.f32.f16 l1, h1, %3;\n\t"
        "cvt.rn.f16x2.f32 %1, l1, l0;\n\t}"
        : "=&r"(hp), "=r"(lp)
        : "f"(v0), "f"(v1));
}

// 16 edges (edge blocks 2u, 2u+1 of the sub-tile) of input channel ch -> FP16 hi / -lo, two 16-byte stores each
__device__ __forceinline__ void store_unit(unsigned char* b_hi, unsigned char* b_lo, int ch, int u, const float (&v)[16]) {
    uint32_t hp[8], lp[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split2_neg(v[2 * q], v[2 * q + 1], hp[q], lp[q]);
    const uint32_t off = (uint32_t)(2 * u) * FB_SBO + (uint32_t)(ch >> 3) * FB_LBO + (uint32_t)(ch & 7) * 16u;
    *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(b_hi + off + FB_SBO) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
    *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    *reinterpret_cast<uint4*>(b_lo + off + FB_SBO) = make_uint4(lp[4], lp[5], lp[6], lp[7]);
}

// one sub-tile GEMM D[128 x N] = W[128 x 128] . B[128 x N] as 3 x 8 MMAs: Wh.(-(-Bl)), Wl.Bh, Wh.Bh
__device__ __forceinline__ void issue_sub_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int N) {
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(CH >> 4) << 24);
    const uint64_t d_hi = smem_desc_ns(b_hi, FB_LBO, FB_SBO), d_lo = smem_desc_ns(b_lo, FB_LBO, FB_SBO);
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint32_t a = (term == 1) ? a_lo : a_hi;
        const uint64_t d = (term == 0) ? d_lo : d_hi;
        const uint32_t id = (term == 0) ? (idesc | (1u << 14)) : idesc;      // bit 14: negate B
#pragma unroll
        for (int ks = 0; ks < CH / 16; ++ks) {
            umma_f16_ts(tmem_d, a + ks * 8, d + (uint64_t)((ks * 2 * FB_LBO) >> 4), id, acc);
            acc = 1;
        }
    }
}

}  // namespace

// Layer matrices of the fused forward.  There is no non-linearity between a block's preconv and conv1
// (ops.py:125-131: x = preconv(x); x = conv1(x) -> context norm), so the two 128x128 layers are folded into one:
//     Wf = W1 . Wp,   bf = W1 . bp + b1        (FP64 accumulation, rounded once to FP32; tc_fold_prep_kernel)
// which removes a third of the GEMM layers (24 instead of 36 per net; SURVEY 8d: report F_m with 24).
// Per matrix m = (net, block, j) with j = 0: Wf, j = 1: W2 this kernel writes the power-of-two FP16 scale, the bias
// and the pre-split weight image [row = out channel][128 x u32]: columns 0-63 the FP16 pairs (k = 2c, 2c+1) of
// scale*W (hi part), columns 64-127 the lo parts — exactly the tensor-memory image of the A operand.
__global__ void __launch_bounds__(1024) fused_prep_kernel(const float* __restrict__ p4, const float* __restrict__ p6, int depth,
                                                         const float* __restrict__ fold, float2* __restrict__ scales2,
                                                         float* __restrict__ bias2, uint32_t* __restrict__ img) {
    extern __shared__ float wf_s[];                           // [in][out] FP32, 64 KB
    __shared__ float red[32];
    __shared__ float scale_s;
    const int m = blockIdx.x;
    const int j = m & 1, blk = (m >> 1) % depth, net = m / (2 * depth);
    const int cin = net == 0 ? 4 : 6;
    const float* prm = net == 0 ? p4 : p6;
    const int tid = threadIdx.x;
    if (j == 0) {                                            // folded layer, formed by tc_fold_prep_kernel (gmw_mlp_tc.cu)
        const float* F = fold + ((int64_t)net * depth + blk) * FOLD_STRIDE;
        for (int i = tid; i < CH * CH; i += 1024) wf_s[i] = F[i];
        if (tid < CH) bias2[m * CH + tid] = F[CH * CH + tid];
    } else {
        const float* W2 = prm + blob_w(cin, blk, 2);
        for (int i = tid; i < CH * CH; i += 1024) wf_s[i] = W2[i];
        if (tid < CH) bias2[m * CH + tid] = prm[blob_b(cin, blk, 2) + tid];
    }
    __syncthreads();
    // per-matrix power-of-two scale that puts max|W| in [512, 1024): FP16 hi/lo both stay normal
    float mx = 0.f;
    for (int i = tid; i < CH * CH; i += 1024) mx = fmaxf(mx, fabsf(wf_s[i]));
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        mx = 0.f;
        for (int w = 0; w < 32; ++w) mx = fmaxf(mx, red[w]);
        int e = 0;
        if (mx > 0.f) frexpf(mx, &e);
        scales2[m] = make_float2(ldexpf(1.f, 10 - e), ldexpf(1.f, e - 10));
        scale_s = ldexpf(1.f, 10 - e);
    }
    __syncthreads();
    const float scale = scale_s;
    uint32_t* out = img + (size_t)m * CH * CH;
    for (int idx = tid; idx < CH * 64; idx += 1024) {
        const int row = idx & 127, c = idx >> 7;
        const float w0 = wf_s[(2 * c) * CH + row] * scale, w1 = wf_s[(2 * c + 1) * CH + row] * scale;
        const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
        const __half l0 = __float2half_rn(w0 - __half2float(h0)), l1 = __float2half_rn(w1 - __half2float(h1));
        out[row * CH + c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        out[row * CH + 64 + c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
}

namespace {

__global__ void __launch_bounds__(FTHREADS, 1)
mlp_fused_kernel(MlpArgs a, const float2* __restrict__ scales, const float* __restrict__ bias2, const uint32_t* __restrict__ wimg,
                 float2* xg, float* __restrict__ reg_w, float4* park) {
    const WsLayout& L = a.L;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (uniform datapath)
    const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
    TR_DECL
    const int ch = 32 * quarter + lane;                      // this thread's channel = TMEM lane = weight row
    const uint32_t rank = blockIdx.x % FCS, group = blockIdx.x / FCS;   // 16 consecutive CTAs form a group (co-resident: cooperative launch)
    const int E = L.E, EP = L.EP, depth = L.depth;
    const int ES = 8 * ((E + 127) / 128);                    // edges per CTA and object (16 * ES == EP)
    const int xrows = ES >> 2;                               // float4 rows of the residual stream per object
    const int nsub = (ES + FSUB - 1) / FSUB;
    const int nphase = 2 * depth;                            // per block: folded preconv.conv1, conv2

    extern __shared__ __align__(1024) unsigned char smem[];
    float4* Xs = reinterpret_cast<float4*>(smem + SMF_X);
    unsigned char* Bbuf = smem + SMF_B;
    float2* part2_s = reinterpret_cast<float2*>(smem + SMF_PART);            // [2 (alternating)][3][128] (mean, M2) per unit stream
    float2* tab_s = part2_s + 2 * 3 * CH;                                    // counts, see below
    float2* stat_s = tab_s + 32;                                             // [2 object halves][128] (mean, rstd) of the latest context norm
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SMF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SMF_BAR + 8 * BAR_COUNT);

    if (tid == 0) {
        mbar_init(bar + BAR_FULL0, FCONV_WARPS);
        mbar_init(bar + BAR_FULL1, FCONV_WARPS);
        mbar_init(bar + BAR_DONE0, 1);
        mbar_init(bar + BAR_DONE1, 1);
        mbar_init(bar + BAR_PDONE, 1);
        mbar_init(bar + BAR_WREADY, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // statistics merge weights: tab_s[q] = (n_q, n_q / n_cta) for the 3 unit streams of this CTA (q = 0..2),
    // tab_s[4 + r] = (n_r, n_r / E) for the 16 slices of an object
    if (tid < 3) {
        const int vld = max(0, min(ES, E - (int)rank * ES));
        int nq = 0;
        for (int col0 = 16 * tid; col0 < ES; col0 += FSUB) nq += max(0, min(16, vld - col0));
        tab_s[tid] = make_float2((float)nq, vld > 0 ? (float)nq / (float)vld : 0.f);
    } else if (tid >= 4 && tid < 4 + FCS) {
        const int nr = max(0, min(ES, E - (tid - 4) * ES));
        tab_s[tid] = make_float2((float)nr, (float)nr / (float)E);
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * quarter) << 16);
    // Work list of a group: "duals" = (net, object pair 2p, 2p + 1) — the two objects share the layer weights and take turns
    // layer by layer.  PAIRED (reg_w != nullptr): pairs group, group + G, ... with both nets of a pair back to back (net 0 then
    // net 1), so that the second net's final pass finds the first net's features parked by the very same threads and emits
    // the edge weights itself (see the epilogue).  Otherwise (features requested, or too few objects): duals 2 p + net dealt
    // round-robin, final features to global memory.  An odd object count runs its last object in both halves (one output).
    const int64_t ngroups = gridDim.x / FCS;
    const bool paired = reg_w != nullptr;
    const int64_t npairs = (L.N + 1) >> 1;
    const int64_t nlist = paired ? npairs : 2 * npairs;      // entries dealt round-robin: pairs (paired) or duals
    const int64_t nmine = nlist > (int64_t)group ? (nlist - 1 - group) / ngroups + 1 : 0;
    const int64_t nd = paired ? 2 * nmine : nmine;           // duals of this group
    auto dual_pair = [&](int64_t d) { return paired ? (int64_t)group + (d >> 1) * ngroups : ((int64_t)group + d * ngroups) >> 1; };
    auto dual_net = [&](int64_t d) { return paired ? (int)(d & 1) : (int)(((int64_t)group + d * ngroups) & 1); };

    if (warp >= FCONV_WARPS) {
        // =====================================================================================================
        // service warps: stream the layers' weight images into tensor memory; the first one issues the MMAs
        // =====================================================================================================
        uint32_t wr[64];
        auto w_src = [&](int mat) { return wimg + ((size_t)mat * CH + ch) * CH; };
        auto w_load = [&](const uint32_t* src, int half) {
#pragma unroll
            for (int i = 0; i < 8; ++i) ld_weights8(src + 64 * half + 8 * i, wr + 8 * i);
        };
        auto w_store = [&](int half) {
            uint32_t t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = wr[i];
            tmem_st32(t_lane + FT_W + 64 * half, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = wr[32 + i];
            tmem_st32(t_lane + FT_W + 64 * half + 32, t);
        };
        auto w_publish = [&]() {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar + BAR_WREADY);
        };
        uint32_t g = 0;                                       // running sub-tile step: operand buffer = g & 1
        uint32_t fpar = 0, wpar = 0, ppar = 0;                // parities: full[2] (bits), wready, pdone
        // ---- context-norm statistics (service warps 1..3; the MMA warp takes no part).  The converters post one (mean, M2) per
        //      unit stream and channel when a layer output of an object half is complete (NB_PART).  A statistics thread owns a
        //      channel (the first warp two): it merges the 3 unit streams, publishes the slice's (mean, M2) in global memory for
        //      the 15 other CTAs of the group (flagged words, see st_flagged), collects theirs, merges in rank order and posts
        //      (mean, rstd) of the layer in shared memory (NB_STAT0/1).  All counts are constants of the launch: no divisions.
        //      The converters never wait for the exchange unless it is slower than the other half's layer.
        const bool stat_warp = warp > FCONV_WARPS;
        const int nrep = warp == FCONV_WARPS + 1 ? 2 : 1;
        const float inv_em1 = 1.0f / (float)(E - 1);
        uint32_t sxc0 = 0, sxc1 = 0;                          // exchanges done per object half: slot = count & 1, flag = (count >> 1) & 1
        uint32_t ecount = 0;                                  // events handled (alternates the partials buffer)
        auto stat_event = [&](int sig) {
            TR(2, 700 + sig);
            nbar_sync(NB_PART, NB_THREADS);
            TR(2, 710 + sig);
            const float2* pbuf = part2_s + (ecount & 1u) * 3 * CH;
            const uint32_t xc = sig ? sxc1 : sxc0;
            const uint32_t slot = xc & 1u, flag = (xc >> 1) & 1u;
            float2* xrow0 = xg + ((size_t)((group * 2 + sig) * 2 + slot) * FCS) * CH;      // [rank][128] of this group, half and slot
            for (int rep = 0; rep < nrep; ++rep) {
                const int c = rep ? 96 + lane : 32 * (warp - FCONV_WARPS - 1) + lane;
                float2 q[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) q[k] = pbuf[k * CH + c];
                float m = 0.f, M2 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) m = fmaf(tab_s[k].y, q[k].x, m);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float dd = q[k].x - m;
                    M2 += fmaf(tab_s[k].x * dd, dd, q[k].y);
                }
                st_flagged(xrow0 + rank * CH + c, m, M2, flag);
            }
            TR(2, 720 + sig);
            for (int rep = 0; rep < nrep; ++rep) {
                const int c = rep ? 96 + lane : 32 * (warp - FCONV_WARPS - 1) + lane;
                const float2* xrow = xrow0 + c;
                // all 16 slices (the own one included) in flight at once, re-read until every word carries this exchange's flag
                float2 v[FCS];
                for (;;) {
#pragma unroll
                    for (int r = 0; r < FCS; ++r) v[r] = ld_volatile_f2(xrow + r * CH);
                    uint32_t bad = 0;
#pragma unroll
                    for (int r = 0; r < FCS; ++r) bad |= (__float_as_uint(v[r].y) >> 31) ^ flag;
                    if (!bad) break;
                    __nanosleep(100);                                             // (slices still missing: poll at a low rate)
                }
                float m = 0.f, M2 = 0.f;
#pragma unroll
                for (int r = 0; r < FCS; ++r) m = fmaf(tab_s[4 + r].y, v[r].x, m);
#pragma unroll
                for (int r = 0; r < FCS; ++r) {
                    const float dd = v[r].x - m;
                    M2 += fmaf(tab_s[4 + r].x * dd, dd, __uint_as_float(__float_as_uint(v[r].y) & 0x7fffffffu));
                }
                const float var = M2 * inv_em1;
                stat_s[sig * CH + c] = make_float2(m, 1.0f / sqrtf(var + 1e-3f));
            }
            nbar_arrive(NB_STAT0 + sig, NB_THREADS);
            TR(2, 730 + sig);
            ++ecount;
            if (sig) ++sxc1; else ++sxc0;
        };
        if (nd > 0) {
            const uint32_t* src = w_src(dual_net(0) * nphase);
            w_load(src, 0);
            w_store(0);
            w_load(src, 1);
            w_store(1);
            w_publish();
        }
        for (int64_t d = 0; d < nd; ++d) {
            const int mat_base = dual_net(d) * nphase;
            for (int ph = 0; ph <= nphase; ++ph) {            // (ph == nphase: only the last layer's second statistics event)
                if (warp == FCONV_WARPS) {
                    if (ph == nphase) break;
                    TR(1, 10);
                    mbar_wait(bar + BAR_WREADY, wpar);
                    tc_fence_after();
                    TR(1, 11);
                    for (int sig = 0; sig < 2; ++sig) {
                        for (int s = 0; s < nsub; ++s, ++g) {
                            const uint32_t b = g & 1u;
                            TR(1, 100 + 10 * sig + s);
                            mbar_wait(bar + BAR_FULL0 + b, (fpar >> b) & 1u);
                            fpar ^= 1u << b;
                            tc_fence_after();
                            TR(1, 200 + 10 * sig + s);
                            if (elect_one()) {
                                const uint32_t b_hi = smem_u32(Bbuf) + b * 2 * FB_PART;
                                // (N is a multiple of 16: a last sub-tile of 8 or 24 edges computes 8 columns nobody reads)
                                issue_sub_gemm(tmem_base + FT_SLOT * sig + FSUB * s, tmem_base + FT_W, tmem_base + FT_W + 64, b_hi,
                                               b_hi + FB_PART, min(FSUB, (ES - FSUB * s + 15) & ~15));
                                umma_commit(bar + BAR_DONE0 + b);
                                if (sig == 1 && s == nsub - 1) umma_commit(bar + BAR_PDONE);
                            }
                            __syncwarp();
                        }
                    }
                } else {
                    // layer outputs in the order the converters complete them: (half 1, ph - 1) during this layer's first half,
                    // (half 0, ph) during its second half
                    for (int ev = ph > 0 ? 0 : 1; ev < (ph < nphase ? 2 : 1); ++ev) stat_event(ev ^ 1);
                    if (ph == nphase) break;
                }
                wpar ^= 1u;
                // next layer's weights: first half fetched while this layer's last MMAs drain
                const bool last_ph = ph + 1 == nphase;
                const bool more = !last_ph || d + 1 < nd;
                if (more) {
                    const uint32_t* src = w_src(last_ph ? dual_net(d + 1) * nphase : mat_base + ph + 1);
                    w_load(src, 0);
                    TR(warp == FCONV_WARPS ? 1 : 2, 20);
                    mbar_wait_relaxed(bar + BAR_PDONE, ppar);
                    tc_fence_after();
                    TR(warp == FCONV_WARPS ? 1 : 2, 21);
                    w_store(0);
                    w_load(src, 1);
                    w_store(1);
                    w_publish();
                    TR(warp == FCONV_WARPS ? 1 : 2, 22);
                } else {
                    mbar_wait_relaxed(bar + BAR_PDONE, ppar);
                }
                ppar ^= 1u;
            }
        }
    } else {
        // =====================================================================================================
        // converter warps: thread = channel; units of 16 edges; produce the B operands, consume the accumulators
        // =====================================================================================================
        const int wg = warp >> 2;                             // unit inside a sub-tile
        const int e_base = (int)rank * ES;
        const int valid = max(0, min(ES, E - e_base));
        uint32_t g = 0;                                       // running sub-tile step (same sequence as the MMA warp)
        uint32_t par = 0, pend = 0;                           // per operand buffer: next wait parity, MMA in flight
        auto wait_buf = [&](uint32_t b) {
            if ((pend >> b) & 1u) {
                mbar_wait(bar + BAR_DONE0 + b, (par >> b) & 1u);
                par ^= 1u << b;
                pend &= ~(1u << b);
                tc_fence_after();
            }
        };
        uint32_t fcount = 0;                                  // finalisations done (alternates the partials buffer)
        const float inv_cnt_wg = tab_s[wg].x > 0.f ? 1.0f / tab_s[wg].x : 0.f;

        for (int64_t d = 0; d < nd; ++d) {
            const int64_t pair = dual_pair(d);
            const int net = dual_net(d);
            const int64_t objA = 2 * pair;
            const bool validB = 2 * pair + 1 < L.N;
            const int64_t objB = validB ? objA + 1 : objA;
            const int cin = net == 0 ? 4 : 6;
            const float* __restrict__ prm = a.params[net];
            const int mat_base = net * nphase;

            // ---- edge features of the two slices, staged in the (idle) operand buffers: f_s[2][6][ES]
            wait_buf(0);
            wait_buf(1);
            conv_sync();
            float* f_s = reinterpret_cast<float*>(Bbuf);
            if (tid < 2 * ES) {
                const int sig = tid >= ES ? 1 : 0;
                const int el = tid - sig * ES;
                const int e = e_base + el;
                const int64_t obj = sig ? objB : objA;
                int i, j;
                decode_edge(e < E ? e : E - 1, L.n, i, j);
                float f[6];
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
#pragma unroll
                for (int r = 0; r < 6; ++r) f_s[(sig * 6 + r) * ES + el] = f[r];
            }
            conv_sync();
            // ---- conv_in: X0 = W_in . f + b_in   (the features are warp-wide broadcasts)
            {
                float wq[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) wq[q] = (q < cin) ? __ldg(prm + blob_in_w() + q * CH + ch) : 0.f;
                const float b = __ldg(prm + blob_in_b(cin) + ch);
                for (int sig = 0; sig < 2; ++sig) {
                    const float* fo = f_s + sig * 6 * ES;
                    for (int col0 = 16 * wg; col0 < ES; col0 += FSUB) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            if ((col0 >> 2) + q4 >= xrows) break;
                            float x[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int e = col0 + 4 * q4 + i;
                                float acc = b;
#pragma unroll
                                for (int r = 0; r < 6; ++r) acc = fmaf(wq[r], fo[r * ES + e], acc);
                                x[i] = (e < valid) ? acc : 0.f;
                            }
                            Xs[(sig * xrows + (col0 >> 2) + q4) * CH + ch] = make_float4(x[0], x[1], x[2], x[3]);
                        }
                    }
                }
            }
            conv_sync();                                      // the features are consumed: operand buffers free

            // ---- the layers.  Stream of steps (layer, half, sub-tile); the two halves alternate per layer, so the statistics
            //      exchange of one object's layer (drain of its last MMAs, L2 round trip between the 16 CTAs) is hidden behind
            //      the other object's steps.
            float un_out = 0.f, b_out = 0.f, un_prev = 0.f, b_prev = 0.f;   // scale / bias of the matrices of this and the previous layer
            int ph_now = 0;
            int qa = -1, qb = -1;                             // sub-tiles whose statistics are due (older, newer): s | half << 8 | (layer & 1) << 9 | buffer << 10
            int fin0 = 0, fin1 = 0;                           // layers whose statistics are final (published), per half
            float K = 0.f, s1 = 0.f, s2 = 0.f, bK = 0.f;      // shifted sums of the layer output being accumulated
            bool have_K = false;
            bool cv_ready = false;                            // the next unit's accumulators are already in flight
            uint32_t cv[16], sv[16];

            // statistics of one unit from its raw accumulators (un, bb: scale and bias of the producing matrix)
            auto stats_math = [&](const uint32_t (&raw)[16], int nv, float un, float bb) {
                if (!have_K) {                                // first unit: choose the shift K = mean of this unit
                    float v[16], sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = fmaf(__uint_as_float(raw[i]), un, bb);
                        if (i < nv) sum += v[i];
                    }
                    K = sum / (float)nv;
                    bK = bb - K;
                    have_K = true;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < nv) {
                            const float dd = v[i] - K;
                            s1 += dd;
                            s2 = fmaf(dd, dd, s2);
                        }
                } else if (nv == 16) {
                    float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float d0, d1, d2, d3;
                        ffma2_bc(d0, d1, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), un, bK);
                        ffma2_bc(d2, d3, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]), un, bK);
                        fadd2_acc(t1[0], t1[1], d0, d1);
                        fadd2_acc(t1[2], t1[3], d2, d3);
                        fsq2_acc(t2[0], t2[1], d0, d1);
                        fsq2_acc(t2[2], t2[3], d2, d3);
                    }
                    s1 += (t1[0] + t1[1]) + (t1[2] + t1[3]);
                    s2 += (t2[0] + t2[1]) + (t2[2] + t2[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < nv) {
                            const float dd = fmaf(__uint_as_float(raw[i]), un, bK);
                            s1 += dd;
                            s2 = fmaf(dd, dd, s2);
                        }
                }
            };
            // one layer output of half `sig` is complete: hand this unit stream's (mean, M2) to the statistics warps (which merge,
            // publish, collect the 15 other slices of the group and post (mean, rstd) of the layer: see the service warps)
            auto finalize = [&](int sig) {
                float2* pbuf = part2_s + (fcount & 1u) * 3 * CH;
                const float m = K + s1 * inv_cnt_wg;
                const float M2 = fmaxf(fmaf(-s1 * inv_cnt_wg, s1, s2), 0.f);
                pbuf[wg * CH + ch] = make_float2(m, M2);
                nbar_arrive(NB_PART, NB_THREADS);
                TR(0, 603);
                K = 0.f; s1 = 0.f; s2 = 0.f; bK = 0.f;
                have_K = false;
                ++fcount;
                if (sig) ++fin1; else ++fin0;
            };
            // statistics of the oldest pending sub-tile; `wait`: its MMA is not yet known to be complete; `loaded`: its accumulators
            // are already in sv (loaded and waited for by the step)
            auto process_oldest = [&](bool wait, bool loaded) {
                const int e = qa;
                qa = qb;
                qb = -1;
                const int s_e = e & 0xff, sig_e = (e >> 8) & 1, pp = (e >> 9) & 1;
                if (wait) wait_buf((uint32_t)(e >> 10) & 1u);
                const int col = FSUB * s_e + 16 * wg;
                const int nv = min(16, valid - col);
                if (nv > 0) {
                    if (!loaded) {
                        tmem_ld16_issue(t_lane + FT_SLOT * sig_e + col, sv);
                        tmem_ld16_wait(sv);
                    }
                    const bool cur = pp == (ph_now & 1);
                    stats_math(sv, nv, cur ? un_out : un_prev, cur ? b_out : b_prev);
                }
                if (s_e == nsub - 1) finalize(sig_e);
            };
            auto ensure_finalized = [&](int sig, int need) {
                while ((sig ? fin1 : fin0) < need && qa >= 0) process_oldest(true, false);
            };
            // (mean, rstd) of the latest finalised layer of half `sig`, posted by the statistics warps
            auto collect = [&](int sig) {
                TR(0, 606);
                nbar_sync(NB_STAT0 + sig, NB_THREADS);
                TR(0, 608);
                return stat_s[sig * CH + ch];
            };

            float fin_a0 = 0.f, fin_c0 = 0.f, fin_a1 = 0.f, fin_c1 = 0.f;    // the final features' transform per half
            for (int ph = 0; ph <= nphase; ++ph) {            // (ph == nphase: only the last layer's statistics are collected)
                const int kind = (ph & 1) ? 2 : 0;            // 0: (residual update ->) folded preconv.conv1, 2: conv2
                ph_now = ph;
                un_prev = un_out;
                b_prev = b_out;
                if (ph < nphase) {
                    un_out = __ldg(scales + mat_base + ph).y;
                    b_out = __ldg(bias2 + (mat_base + ph) * CH + ch);
                }
                const bool reads_d = ph > 0;
                for (int sig = 0; sig < 2; ++sig) {
                    // input transform of this layer as one FMA on the raw accumulator: the context norm of the producing
                    // layer folded in ((d*un + b - mean) * rstd)
                    float a_in = 0.f, c_in = 0.f;
                    if (reads_d) {
                        ensure_finalized(sig, ph);
                        const float2 st = collect(sig);
                        a_in = un_prev * st.y;
                        c_in = (b_prev - st.x) * st.y;
                    }
                    if (ph == nphase) {
                        if (sig) { fin_a1 = a_in; fin_c1 = c_in; } else { fin_a0 = a_in; fin_c0 = c_in; }
                        continue;
                    }
                    const uint32_t t_half = t_lane + FT_SLOT * sig;
                    float4* Xh = Xs + (size_t)(sig * xrows) * CH + ch;
                    // Software pipeline over the steps: the accumulators of the NEXT unit to convert (cv) and of the unit whose
                    // statistics are due (sv, two steps behind: its MMA is known complete when its operand buffer comes free)
                    // are in flight from tensor memory while the current unit is being processed.
                    for (int s = 0; s < nsub; ++s, ++g) {
                        const uint32_t b = g & 1u;
                        TR(0, 1000 * kind + 100 + 10 * sig + s);
                        wait_buf(b);
                        TR(0, 1000 * kind + 200 + 10 * sig + s);
                        unsigned char* b_hi = Bbuf + (size_t)b * 2 * FB_PART;
                        const int col0 = FSUB * s + 16 * wg;
                        const bool active = col0 < ES;
                        const int rows = min(4, xrows - (col0 >> 2));         // float4 rows of this unit inside the slice (last unit: 2)
                        if (reads_d && active && !cv_ready) tmem_ld16_issue(t_half + col0, cv);
                        // the sub-tile of two steps ago (same operand buffer: its MMA is complete) is due for its statistics
                        const bool due = qb >= 0;
                        bool st_load = false;
                        if (due) {
                            const int col = FSUB * (qa & 0xff) + 16 * wg;
                            st_load = valid > col;
                        }
                        float v[16];
                        float4* Xp = Xh + (col0 >> 2) * CH;
                        float4 x4[4];
                        if (active && kind == 0) {
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) x4[q4] = (q4 < rows) ? Xp[q4 * CH] : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if (reads_d && active) tmem_ld16_wait(cv);
                        cv_ready = false;
                        TR(0, 1000 * kind + 300 + 10 * sig + s);
                        if (st_load) tmem_ld16_issue(t_lane + FT_SLOT * ((qa >> 8) & 1) + FSUB * (qa & 0xff) + 16 * wg, sv);
                        if (active) {
                            if (kind == 0) {
                                if (reads_d) {
#pragma unroll
                                    for (int i = 0; i < 16; i += 2) {
                                        ffma2_bc(v[i], v[i + 1], __uint_as_float(cv[i]), __uint_as_float(cv[i + 1]), a_in, c_in);
                                        v[i] = fmaxf(v[i], 0.f);
                                        v[i + 1] = fmaxf(v[i + 1], 0.f);
                                    }
#pragma unroll
                                    for (int q4 = 0; q4 < 4; ++q4) {
                                        fadd2_acc(v[4 * q4], v[4 * q4 + 1], x4[q4].x, x4[q4].y);
                                        fadd2_acc(v[4 * q4 + 2], v[4 * q4 + 3], x4[q4].z, x4[q4].w);
                                    }
                                    if (col0 + 16 > valid) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i)
                                            if (col0 + i >= valid) v[i] = 0.f;
                                    }
#pragma unroll
                                    for (int q4 = 0; q4 < 4; ++q4)
                                        if (q4 < rows) Xp[q4 * CH] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                                } else {
#pragma unroll
                                    for (int q4 = 0; q4 < 4; ++q4) {
                                        v[4 * q4] = x4[q4].x; v[4 * q4 + 1] = x4[q4].y; v[4 * q4 + 2] = x4[q4].z; v[4 * q4 + 3] = x4[q4].w;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; i += 2)
                                    ffma2_bc(v[i], v[i + 1], __uint_as_float(cv[i]), __uint_as_float(cv[i + 1]), a_in, c_in);
                            }
                            store_unit(b_hi, b_hi + FB_PART, ch, wg, v);
                        }
                        TR(0, 1000 * kind + 400 + 10 * sig + s);
                        fence_async_smem();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar + BAR_FULL0 + b);
                        pend |= 1u << b;
                        TR(0, 1000 * kind + 500 + 10 * sig + s);
                        // the next unit to convert: next sub-tile of this half, else the other half's first (whose producing MMAs
                        // are complete when this segment has at least two steps)
                        {
                            const bool cross = s + 1 == nsub;
                            const int nsig = cross ? (sig ^ 1) : sig;
                            const int nph = (cross && sig == 1) ? ph + 1 : ph;
                            const int ncol = cross ? 16 * wg : col0 + FSUB;
                            if (st_load) tmem_ld16_wait(sv);          // (before the next load is issued: the wait covers all loads in flight)
                            TR(0, 1000 * kind + 550 + 10 * sig + s);
                            if (nph > 0 && nph < nphase && ncol < ES && (!cross || nsub >= 2)) {
                                tmem_ld16_issue(t_lane + FT_SLOT * nsig + ncol, cv);
                                cv_ready = true;
                            }
                        }
                        if (due) process_oldest(false, true);
                        TR(0, 1000 * kind + 570 + 10 * sig + s);
                        const int ent = s | (sig << 8) | ((ph & 1) << 9) | ((int)b << 10);
                        if (qa < 0) qa = ent; else qb = ent;
                    }
                }
            }
            TR(0, 600);

            // ---- final features x = relu(cn(Y2)) + X of both halves (all MMAs are complete: the operand buffers are free)
            TR(0, 601);
            int itn = 0;
            for (int sig = 0; sig < 2; ++sig) {
                const float a_fin = sig ? fin_a1 : fin_a0, c_fin = sig ? fin_c1 : fin_c0;
                const int64_t obj = sig ? objB : objA;
                const bool emit = sig == 0 || validB;
                const uint32_t t_half = t_lane + FT_SLOT * sig;
                const float4* Xh = Xs + (size_t)(sig * xrows) * CH + ch;
                auto final_unit = [&](int col0, float (&v)[16]) {
                    const float4* Xp = Xh + (col0 >> 2) * CH;
                    const int rows = min(4, xrows - (col0 >> 2));
                    float4 x4[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) x4[q4] = (q4 < rows) ? Xp[q4 * CH] : make_float4(0.f, 0.f, 0.f, 0.f);
                    tmem_ld16(t_half + col0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(fmaf(v[i], a_fin, c_fin), 0.f);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        v[4 * q4] += x4[q4].x; v[4 * q4 + 1] += x4[q4].y; v[4 * q4 + 2] += x4[q4].z; v[4 * q4 + 3] += x4[q4].w;
                    }
                    if (col0 + 16 > valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (col0 + i >= valid) v[i] = 0.f;
                    }
                };
                if (!paired) {
                    // -> global, channel-major [obj][128][EP] (consumed by gmw_edge_weight_kernel / the correspondence branch)
                    float* G = act_ptr(a.ws, L, net, 0, SLOT_X) + obj * (int64_t)CH * EP + (int64_t)ch * EP + e_base;
                    for (int col0 = 16 * wg; col0 < ES; col0 += FSUB) {
                        float v[16];
                        final_unit(col0, v);
                        const int rows = min(4, xrows - (col0 >> 2));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            if (emit && q4 < rows)
                                __stcs(reinterpret_cast<float4*>(G + col0 + 4 * q4), make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]));   // streamed: keep the weight image in L2
                    }
                } else if (net == 0) {
                    // park the 4-d net's features in this CTA's slice of an L2-resident scratch ([2][ES/4][128] float4 like Xs; the
                    // 16 CTAs of a group fill exactly two objects' worth of the workspace slot, EP = 16 ES): the thread that
                    // writes a word is the one that reads it back after the 6-d net, so no fence or barrier is needed
                    float4* Pk = park + ((size_t)blockIdx.x * 2 * xrows + (size_t)sig * xrows) * CH + ch;
                    for (int col0 = 16 * wg; col0 < ES; col0 += FSUB) {
                        float v[16];
                        final_unit(col0, v);
                        const int rows = min(4, xrows - (col0 >> 2));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            if (q4 < rows) Pk[((col0 >> 2) + q4) * CH] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                    }
                } else {
                    // edge weights straight from the two nets' final features (GMW/model/model.py:176-181, diagonal of pairwiseL2Dist):
                    // per edge the three channel sums |a|^2, |c|^2, a.c — 32 channels by a halving shuffle tree, the 4 lane quarters
                    // through shared memory (the operand buffers are idle here) — then
                    //   w = 1 / sqrt(max((|c^|^2 - 2 a^.c^) + |a^|^2, 1e-30)),  a^ = a / max(|a|, 1e-12)
                    const float4* Pk = park + ((size_t)blockIdx.x * 2 * xrows + (size_t)sig * xrows) * CH + ch;
                    float* red = reinterpret_cast<float*>(Bbuf) + wg * (2 * 4 * 48);      // [2 buffers][4 quarters][16 edges][3]
                    float* W = reg_w + obj * (int64_t)E + e_base;
                    for (int col0 = 16 * wg; col0 < ES; col0 += FSUB, ++itn) {
                        float cvv[16], aa[16], cc[16], ac[16];
                        final_unit(col0, cvv);
                        const int rows = min(4, xrows - (col0 >> 2));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const float4 p = (q4 < rows) ? Pk[((col0 >> 2) + q4) * CH] : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float av[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float c = cvv[4 * q4 + i];
                                aa[4 * q4 + i] = __fmul_rn(av[i], av[i]);          // (explicit roundings: gmw_edge_weight_kernel<true>
                                cc[4 * q4 + i] = __fmul_rn(c, c);                  //  replays this exact sequence)
                                ac[4 * q4 + i] = __fmul_rn(av[i], c);
                            }
                        }
                        // halving tree over the 32 lanes: afterwards lane l holds the sums of edge l >> 1
#pragma unroll
                        for (int h = 8; h >= 1; h >>= 1) {
                            const bool up = (lane & (2 * h)) != 0;
#pragma unroll
                            for (int k = 0; k < h; ++k) {
                                const float sa = up ? aa[k] : aa[k + h], sc = up ? cc[k] : cc[k + h], sx = up ? ac[k] : ac[k + h];
                                const float ka = up ? aa[k + h] : aa[k], kc = up ? cc[k + h] : cc[k], kx = up ? ac[k + h] : ac[k];
                                aa[k] = __fadd_rn(ka, __shfl_xor_sync(0xffffffffu, sa, 2 * h));
                                cc[k] = __fadd_rn(kc, __shfl_xor_sync(0xffffffffu, sc, 2 * h));
                                ac[k] = __fadd_rn(kx, __shfl_xor_sync(0xffffffffu, sx, 2 * h));
                            }
                        }
                        aa[0] = __fadd_rn(aa[0], __shfl_xor_sync(0xffffffffu, aa[0], 1));
                        cc[0] = __fadd_rn(cc[0], __shfl_xor_sync(0xffffffffu, cc[0], 1));
                        ac[0] = __fadd_rn(ac[0], __shfl_xor_sync(0xffffffffu, ac[0], 1));
                        float* rb = red + (itn & 1) * (4 * 48);
                        if ((lane & 1) == 0) {
                            float* o = rb + quarter * 48 + (lane >> 1) * 3;
                            o[0] = aa[0]; o[1] = cc[0]; o[2] = ac[0];
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory");       // the 4 warps (lane quarters) of this unit
                        if (quarter == 0 && lane < 16) {
                            const int e = col0 + lane;
                            if (e < valid && emit) {
                                const float* o = rb + lane * 3;
                                const float saa = __fadd_rn(__fadd_rn(o[0], o[48]), __fadd_rn(o[96], o[144]));
                                const float scc = __fadd_rn(__fadd_rn(o[1], o[49]), __fadd_rn(o[97], o[145]));
                                const float sac = __fadd_rn(__fadd_rn(o[2], o[50]), __fadd_rn(o[98], o[146]));
                                const float n4 = fmaxf(sqrtf(saa), 1e-12f), n6 = fmaxf(sqrtf(scc), 1e-12f);
                                const float a2 = __fdiv_rn(saa, __fmul_rn(n4, n4)), c2 = __fdiv_rn(scc, __fmul_rn(n6, n6));
                                const float acn = __fdiv_rn(sac, __fmul_rn(n4, n6));
                                const float s2w = __fadd_rn(__fadd_rn(c2, -2.f * acn), a2);
                                W[e] = __fdiv_rn(1.f, sqrtf(fmaxf(s2w, 1e-30f)));
                            }
                        }
                    }
                }
            }
            tc_fence_before();                                // the next dual's MMAs overwrite these columns
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

#ifdef DCD_FUSED_TRACE
}  // namespace dcd
extern "C" __attribute__((visibility("default"))) int dcd_debug_fused_trace(long long* dst, int* n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, dcd::g_trace, sizeof(long long) * 12288);
    cudaMemcpyFromSymbol(n, dcd::g_trace_n, sizeof(int) * 3);
    return 0;
}
namespace dcd {
#endif

bool gmw_fused_supported(int n) {
    const int E = n * (n - 1) / 2;
    return 8 * ((E + 127) / 128) <= FES_MAX;
}

constexpr int FMAX_GROUPS = 16;                               // exchange buffer sized for up to 256 SMs
constexpr size_t kExchangeBytes = (size_t)FMAX_GROUPS * 2 * 2 * FCS * CH * sizeof(float2);    // [group][object half][slot][rank][128]

// Tail of the workspace used by the fused forward, per matrix m = (net, block, {folded preconv.conv1, conv2}):
//   scales2 [4*depth] float2 (256-byte padded) | bias2 [4*depth][128] | weight image [4*depth][128][128] u32 | exchange buffer
static size_t fused_scales_bytes(int depth) { return (((size_t)4 * depth * sizeof(float2)) + 255) / 256 * 256; }
static size_t fused_bias_bytes(int depth) { return (size_t)4 * depth * CH * sizeof(float); }
static size_t fused_image_only_bytes(int depth) { return (size_t)4 * depth * CH * CH * sizeof(uint32_t); }
size_t gmw_fused_image_bytes(int depth) {
    return fused_scales_bytes(depth) + fused_bias_bytes(depth) + fused_image_only_bytes(depth) + kExchangeBytes;
}

// Runs both nets of all objects.  reg_w != nullptr and at least as many objects as groups: PAIRED schedule, the kernel emits
// the edge weights itself (*emitted = true; nothing but reg_w is written to HBM; the first net's features are parked in the
// otherwise unused SLOT_Y1 area of the inference workspace, one object's worth per group, which stays in L2).  Otherwise the
// final features land in SLOT_X of the workspace (*emitted = false) for gmw_edge_weight_kernel / the correspondence branch.
// `tail` points to gmw_fused_image_bytes(depth) bytes (256-byte aligned).
int launch_gmw_fused_fwd(const MlpArgs& a, const float* params4, const float* params6, void* tail, float* reg_w, bool* emitted,
                         cudaStream_t st) {
    const int depth = a.L.depth;
    unsigned char* base = reinterpret_cast<unsigned char*>(tail);
    float2* scales2 = reinterpret_cast<float2*>(base);
    float* bias2 = reinterpret_cast<float*>(base + fused_scales_bytes(depth));
    uint32_t* wimg = reinterpret_cast<uint32_t*>(base + fused_scales_bytes(depth) + fused_bias_bytes(depth));
    float2* xg = reinterpret_cast<float2*>(base + fused_scales_bytes(depth) + fused_bias_bytes(depth) + fused_image_only_bytes(depth));
    // The CTAs of a group wait for each other, so all of them must be resident: cooperative launch, one CTA per SM.
    // (queried on every call: the library keeps no state between calls; these are host-side lookups, no device work)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DCD_E_DEVICE;
    cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
    cudaFuncSetAttribute(fused_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CH * CH * sizeof(float)));
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mlp_fused_kernel, FTHREADS, kFusedSmem);
    if (!coop || per_sm < 1) return DCD_E_UNSUPPORTED;
    int max_groups = device_sm_count() * per_sm / FCS;
    if (max_groups > FMAX_GROUPS) max_groups = FMAX_GROUPS;
    if (max_groups < 1) return DCD_E_UNSUPPORTED;
    fused_prep_kernel<<<4 * depth, 1024, CH * CH * sizeof(float), st>>>(params4, params6, depth, a.fold, scales2, bias2, wimg);
    cudaMemsetAsync(xg, 0x80, kExchangeBytes, st);            // every word starts with the flag its first use does not expect
    // work is dealt to the groups as object pairs (paired schedule: both nets of a pair back to back) or (pair, net) duals
    const bool paired = reg_w != nullptr && a.L.N >= 2 * (int64_t)max_groups;
    const int64_t npairs = (a.L.N + 1) / 2;
    const int64_t nitems = paired ? npairs : npairs * 2;
    const int ngroups = (int)(nitems < max_groups ? nitems : max_groups);
    MlpArgs args = a;
    const float2* scales_arg = scales2;
    const float* bias_arg = bias2;
    const uint32_t* wimg_arg = wimg;
    float* regw_arg = paired ? reg_w : nullptr;
    float4* park_arg = reinterpret_cast<float4*>(act_ptr(a.ws, a.L, 0, 0, SLOT_Y1));
    void* kargs[] = {&args, &scales_arg, &bias_arg, &wimg_arg, &xg, &regw_arg, &park_arg};
    if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(mlp_fused_kernel), dim3(FCS * ngroups), dim3(FTHREADS), kargs, kFusedSmem,
                                    st) != cudaSuccess) {
        cudaGetLastError();
        return DCD_E_LAUNCH;
    }
    *emitted = paired;
    return DCD_OK;
}

}  // namespace dcd

// GMW edge-feature MLP, forward (FP32 CUDA-core path) + edge weights, for sm_100a.
//
// Replaces GMW.edge_expand / graph_matching / graph_extract (GMW/model/model.py:153-181) and the
// context-normalised 1x1-conv ResNet (GMW/model/yi2018cvpr/model.py:63-67, ops.py:7-19,72-131).
//
// The n x n edge expansion and the E x E distance matrix are never built.  Activations are kept
// channel-major [128][E] per object; a CTA owns a tile of 128 edges and runs the segment between two
// context-norm barriers on it (the norm couples all E edges of an object, so it is a grid-level barrier):
//   FIRST : conv_in from the keypoints          -> preconv -> conv1 -> Y1 + tile statistics
//   B     : CN(Y1)                              -> conv2   -> Y2 + tile statistics
//   CA    : x = ReLU(CN(Y2)) + x (residual)     -> preconv -> conv1 -> Y1 + tile statistics
// and a last elementwise kernel turns the final features of both nets into 1 / ||a - c||.
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

constexpr size_t kFwdSmem = (size_t)(2 * CH * LD + 2 * KC * CH) * sizeof(float) + CH * sizeof(float2);

// accumulators (+bias) -> global [128][EP] tile + per-channel tile statistics (mean, M2)
__device__ __forceinline__ void store_tile_with_stats(float (&acc)[8][8], const float* __restrict__ bias,
                                                      float* __restrict__ out, int EP, int tile, int E,
                                                      float2* __restrict__ part) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int valid = min(TE, E - tile * TE);
    const float cnt = (float)valid;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int ch = own4(ty, r);
        const float b = __ldg(bias + ch);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc[r][q] += b;
            if (own4(tx, q) < valid) s += acc[r][q];
        }
        float* row = out + (int64_t)ch * EP + tile * TE;
        *reinterpret_cast<float4*>(row + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(row + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        const float mean = half_warp_sum(s) / cnt;
        float m2 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (own4(tx, q) < valid) {
                const float d = acc[r][q] - mean;
                m2 += d * d;
            }
        m2 = half_warp_sum(m2);
        if (tx == 0) part[ch] = make_float2(mean, m2);
    }
}

template <int MODE>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_fwd_kernel(MlpArgs a, int blk) {
    const WsLayout& L = a.L;
    const int tile = blockIdx.x % L.T;
    const int64_t obj = blockIdx.x / L.T;
    const int net = blockIdx.y;
    const int cin = net == 0 ? 4 : 6;
    const float* __restrict__ prm = a.params[net];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int E = L.E, EP = L.EP;
    const int64_t obj_off = obj * (int64_t)CH * EP;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* A_s = reinterpret_cast<float*>(smem_raw);
    float* B_s = A_s + CH * LD;
    float* Wc_s = B_s + CH * LD;
    float2* stat_s = reinterpret_cast<float2*>(Wc_s + 2 * KC * CH);

    // ---- source tile -> A_s
    if (MODE == MODE_FIRST) {
        const int el = tid & 127, half = tid >> 7;
        const int e = tile * TE + el;
        const bool ok = e < E;
        int i, j;
        decode_edge(ok ? e : E - 1, L.n, i, j);
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
        const float* Win = prm + blob_in_w();
        const float* bin = prm + blob_in_b(cin);
        float* X0 = act_ptr(a.ws, L, net, 0, SLOT_X) + obj_off;
        for (int c = half * 64; c < half * 64 + 64; ++c) {
            float x = __ldg(bin + c);
            for (int q = 0; q < cin; ++q) x = fmaf(__ldg(Win + q * CH + c), f[q], x);
            if (!ok) x = 0.f;
            A_s[c * LD + el] = x;
            X0[(int64_t)c * EP + e] = x;
        }
    } else {
        const int pb = (MODE == MODE_B) ? blk : blk - 1;            // block whose statistics are consumed
        const int which = (MODE == MODE_B) ? 0 : 1;
        if (tid < CH)
            stat_s[tid] = merge_cn_stats(stat_ptr(a.ws, L, net, pb, which) + obj * (int64_t)L.T * CH, tid, L.T, E);
        __syncthreads();
        const float* Y = act_ptr(a.ws, L, net, pb, MODE == MODE_B ? SLOT_Y1 : SLOT_Y2) + obj_off;
        const float* Xp = (MODE == MODE_CA) ? act_ptr(a.ws, L, net, blk - 1, SLOT_X) + obj_off : nullptr;
        float* Xn = (MODE == MODE_CA) ? act_ptr(a.ws, L, net, blk, SLOT_X) + obj_off : nullptr;
#pragma unroll 4
        for (int it = 0; it < (CH * TE / 4) / MLP_THREADS; ++it) {
            const int id = it * MLP_THREADS + tid;
            const int row = id >> 5, c4 = (id & 31) * 4;
            const int e0 = tile * TE + c4;
            const float2 st = stat_s[row];
            const float4 y = *reinterpret_cast<const float4*>(Y + (int64_t)row * EP + e0);
            float v[4] = {(y.x - st.x) * st.y, (y.y - st.x) * st.y, (y.z - st.x) * st.y, (y.w - st.x) * st.y};
            if (MODE == MODE_CA) {
                const float4 xp = *reinterpret_cast<const float4*>(Xp + (int64_t)row * EP + e0);
                v[0] = fmaxf(v[0], 0.f) + xp.x;
                v[1] = fmaxf(v[1], 0.f) + xp.y;
                v[2] = fmaxf(v[2], 0.f) + xp.z;
                v[3] = fmaxf(v[3], 0.f) + xp.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (e0 + q >= E) v[q] = 0.f;
            const float4 o = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(A_s + row * LD + c4) = o;
            if (MODE == MODE_CA) *reinterpret_cast<float4*>(Xn + (int64_t)row * EP + e0) = o;
        }
    }

    float acc[8][8];
    zero_acc(acc);
    if (MODE == MODE_B) {
        tile_gemm(prm + blob_w(cin, blk, 2), A_s, Wc_s, acc);
        store_tile_with_stats(acc, prm + blob_b(cin, blk, 2), act_ptr(a.ws, L, net, blk, SLOT_Y2) + obj_off, EP, tile, E,
                              stat_ptr(a.ws, L, net, blk, 1) + (obj * L.T + tile) * (int64_t)CH);
        return;
    }
    // preconv: P = Wp x + bp  (kept in shared memory as the operand of conv1)
    tile_gemm(prm + blob_w(cin, blk, 0), A_s, Wc_s, acc);
    {
        const float* bp = prm + blob_b(cin, blk, 0);
        float* Pg = L.save ? act_ptr(a.ws, L, net, blk, SLOT_P) + obj_off : nullptr;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ch = own4(ty, r);
            const float b = __ldg(bp + ch);
            const float4 lo4 = make_float4(acc[r][0] + b, acc[r][1] + b, acc[r][2] + b, acc[r][3] + b);
            const float4 hi4 = make_float4(acc[r][4] + b, acc[r][5] + b, acc[r][6] + b, acc[r][7] + b);
            *reinterpret_cast<float4*>(B_s + ch * LD + tx * 4) = lo4;
            *reinterpret_cast<float4*>(B_s + ch * LD + 64 + tx * 4) = hi4;
            if (Pg != nullptr) {
                float* row = Pg + (int64_t)ch * EP + tile * TE;
                *reinterpret_cast<float4*>(row + tx * 4) = lo4;
                *reinterpret_cast<float4*>(row + 64 + tx * 4) = hi4;
            }
        }
    }
    zero_acc(acc);
    tile_gemm(prm + blob_w(cin, blk, 1), B_s, Wc_s, acc);
    store_tile_with_stats(acc, prm + blob_b(cin, blk, 1), act_ptr(a.ws, L, net, blk, SLOT_Y1) + obj_off, EP, tile, E,
                          stat_ptr(a.ws, L, net, blk, 0) + (obj * L.T + tile) * (int64_t)CH);
}

// Final features of both nets -> reg_weights.  One thread per edge, channel loop with coalesced rows.
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

int launch_gmw_weights_fwd(const float* kpts2d, const float* kpts3d, const float* params4, const float* params6,
                           int64_t N, int n, int depth, int save, float* reg_w, float* feat4, float* feat6,
                           float* ws, cudaStream_t st) {
    MlpArgs a;
    a.kpts2d = kpts2d; a.kpts3d = kpts3d;
    a.params[0] = params4; a.params[1] = params6;
    a.ws = ws;
    a.L = make_layout(N, n, depth, save);
    if ((int64_t)a.L.T * N > 0x7fffffffLL) return DCD_E_UNSUPPORTED;
    cudaFuncSetAttribute(mlp_fwd_kernel<MODE_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    cudaFuncSetAttribute(mlp_fwd_kernel<MODE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    cudaFuncSetAttribute(mlp_fwd_kernel<MODE_CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    const dim3 grid((unsigned)(a.L.T * N), 2);
    mlp_fwd_kernel<MODE_FIRST><<<grid, MLP_THREADS, kFwdSmem, st>>>(a, 0);
    for (int blk = 0; blk < depth; ++blk) {
        mlp_fwd_kernel<MODE_B><<<grid, MLP_THREADS, kFwdSmem, st>>>(a, blk);
        if (blk + 1 < depth) mlp_fwd_kernel<MODE_CA><<<grid, MLP_THREADS, kFwdSmem, st>>>(a, blk + 1);
    }
    DCD_CHECK_LAUNCH();
    const unsigned g2 = (unsigned)(((a.L.E + 255) / 256) * N);
    gmw_edge_weight_kernel<<<g2, 256, 0, st>>>(a, reg_w, feat4, feat6);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

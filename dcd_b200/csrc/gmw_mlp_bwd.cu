// GMW edge-feature MLP, backward w.r.t. the parameters (FP32 CUDA-core path), for sm_100a.
//
// This is the autograd of GMW/main.py:465 restricted to the regression path
// (reg_weights -> 1/M -> normalise -> 12 residual blocks with context norm -> conv_in), written as
// explicit kernels over the activations saved by the forward (block input x, preconv output P, Y1, Y2):
//   BW0  d reg_weights -> d final features (both nets)
//   per block, last to first:
//     BW1  per-channel sums for the backward of the second context norm (through the ReLU mask)
//     R2   dy2 = CN'(.) ; dW2 += dy2 . yhat1^T ; d yhat1 = W2^T dy2 (+ sums for the first context norm)
//     R3   dy1 = CN'(.) ; dW1 += dy1 . P^T ; dP = W1^T dy1 ; dWp += dP . x^T ; dx = Wp^T dP + residual
//   BW4  conv_in gradient from the edge features rebuilt from the keypoints.
// Weight-gradient tiles are written as per-CTA partials and summed in a fixed order (deterministic).
#include "gmw_mlp_tile.cuh"

namespace dcd {

struct MlpBwdArgs {
    const float* kpts2d;
    const float* kpts3d;
    const float* params[2];
    float* ws;          // forward workspace (save = 1)
    WsLayout L;
    float* Wn;          // [2][depth][3][128][128] weights in native [out][in] layout
    float* G;           // [2][N][128][EP]   gradient w.r.t. the current block output
    float* D1;          // [2][N][128][EP]   gradient w.r.t. yhat1
    float2* bstat;      // [2][2][N][T][128] partial sums of the context-norm backward
    float* wpart;       // [3][2*N*T][128*128 + 128] partial weight/bias gradients
    float* inpart;      // [2*N*T][128][8] partial conv_in gradients
};

namespace {

constexpr int WP = CH * CH + CH;   // floats per partial (weights + bias)
constexpr size_t kBwdSmem = (size_t)(2 * CH * LD + 2 * KC * CH) * sizeof(float) + 3 * CH * sizeof(float2);

__device__ __forceinline__ float* Gp(const MlpBwdArgs& a, int net, int64_t obj) {
    return a.G + ((int64_t)net * a.L.N + obj) * CH * a.L.EP;
}
__device__ __forceinline__ float* D1p(const MlpBwdArgs& a, int net, int64_t obj) {
    return a.D1 + ((int64_t)net * a.L.N + obj) * CH * a.L.EP;
}
__device__ __forceinline__ float2* bstatp(const MlpBwdArgs& a, int net, int which, int64_t obj) {
    return a.bstat + (((int64_t)net * 2 + which) * a.L.N + obj) * a.L.T * CH;
}
__device__ __forceinline__ const float* Wnp(const MlpBwdArgs& a, int net, int blk, int which) {
    return a.Wn + (((int64_t)net * a.L.depth + blk) * 3 + which) * CH * CH;
}

// native [out][in] copies of the 128x128 matrices (the blob stores them transposed for the forward)
__global__ void transpose_weights_kernel(MlpBwdArgs a) {
    const int m = blockIdx.x;                         // (net, blk, which)
    const int which = m % 3, blk = (m / 3) % a.L.depth, net = m / (3 * a.L.depth);
    const int cin = net == 0 ? 4 : 6;
    const float* src = a.params[net] + blob_w(cin, blk, which);      // [in][out]
    float* dst = a.Wn + (int64_t)m * CH * CH;                        // [out][in]
    __shared__ float t[32][33];
    for (int bi = 0; bi < CH; bi += 32)
        for (int bo = 0; bo < CH; bo += 32) {
            for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = src[(bi + r) * CH + bo + threadIdx.x];
            __syncthreads();
            for (int r = threadIdx.y; r < 32; r += blockDim.y) dst[(bo + r) * CH + bi + threadIdx.x] = t[threadIdx.x][r];
            __syncthreads();
        }
}

// BW0: d reg_weights -> d final features.  One thread per edge (mirrors gmw_edge_weight_kernel).
__global__ void __launch_bounds__(256) edge_weight_bwd_kernel(MlpBwdArgs a, const float* __restrict__ grad_w) {
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
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = fmaxf((Y4[(int64_t)c * EP] - s4.x) * s4.y, 0.f) + X4[(int64_t)c * EP];
        const float x6 = fmaxf((Y6[(int64_t)c * EP] - s6.x) * s6.y, 0.f) + X6[(int64_t)c * EP];
        n4 = fmaf(x4, x4, n4);
        n6 = fmaf(x6, x6, n6);
    }
    const float r4 = sqrtf(n4), r6 = sqrtf(n6);
    n4 = fmaxf(r4, 1e-12f);
    n6 = fmaxf(r6, 1e-12f);
    float a2 = 0.f, c2 = 0.f, ac = 0.f;
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = fmaxf((Y4[(int64_t)c * EP] - s4.x) * s4.y, 0.f) + X4[(int64_t)c * EP];
        const float x6 = fmaxf((Y6[(int64_t)c * EP] - s6.x) * s6.y, 0.f) + X6[(int64_t)c * EP];
        const float av = __fdiv_rn(x4, n4), cv = __fdiv_rn(x6, n6);
        a2 = fmaf(av, av, a2);
        c2 = fmaf(cv, cv, c2);
        ac = fmaf(av, cv, ac);
    }
    const float s = __fadd_rn(__fadd_rn(c2, -2.f * ac), a2);
    const float w = __fdiv_rn(1.f, sqrtf(fmaxf(s, 1e-30f)));
    // w = s^(-1/2): dL/ds = -gw * w^3 / 2 (zero where the clamp is active)
    const float gw = __ldg(grad_w + obj * (int64_t)E + e);
    const float q = (s > 1e-30f) ? -0.5f * gw * w * w * w : 0.f;
    const float ada = q * (2.f * a2 - 2.f * ac);     // a . da
    const float cdc = q * (2.f * c2 - 2.f * ac);     // c . dc
    float* G4 = Gp(a, 0, obj) + e;
    float* G6 = Gp(a, 1, obj) + e;
    const bool live4 = r4 > 1e-12f, live6 = r6 > 1e-12f;   // clamp_min(eps) passes no gradient below eps
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = fmaxf((Y4[(int64_t)c * EP] - s4.x) * s4.y, 0.f) + X4[(int64_t)c * EP];
        const float x6 = fmaxf((Y6[(int64_t)c * EP] - s6.x) * s6.y, 0.f) + X6[(int64_t)c * EP];
        const float av = __fdiv_rn(x4, n4), cv = __fdiv_rn(x6, n6);
        const float da = q * (2.f * av - 2.f * cv);
        const float dc = q * (2.f * cv - 2.f * av);
        G4[(int64_t)c * EP] = live4 ? (da - av * ada) / n4 : da / n4;
        G6[(int64_t)c * EP] = live6 ? (dc - cv * cdc) / n6 : dc / n6;
    }
}

// BW1: partial sums (sum dyh, sum dyh*yh) of the second context norm's backward, dyh = G * (yh > 0).
__global__ void __launch_bounds__(256) cn2_bwd_sums_kernel(MlpBwdArgs a, int blk) {
    const WsLayout& L = a.L;
    const int tile = blockIdx.x % L.T;
    const int64_t obj = blockIdx.x / L.T;
    const int net = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int E = L.E, EP = L.EP;
    __shared__ float2 stat_s[CH];
    if (threadIdx.x < CH)
        stat_s[threadIdx.x] = merge_cn_stats(stat_ptr(a.ws, L, net, blk, 1) + obj * (int64_t)L.T * CH, threadIdx.x, L.T, E);
    __syncthreads();
    const float* Y2 = act_ptr(a.ws, L, net, blk, SLOT_Y2) + obj * (int64_t)CH * EP;
    const float* G = Gp(a, net, obj);
    float2* out = bstatp(a, net, 0, obj) + (int64_t)tile * CH;
    const int e0 = tile * TE + lane * 4;
    for (int row = warp * 16; row < warp * 16 + 16; ++row) {
        const float2 st = stat_s[row];
        const float4 y = *reinterpret_cast<const float4*>(Y2 + (int64_t)row * EP + e0);
        const float4 g = *reinterpret_cast<const float4*>(G + (int64_t)row * EP + e0);
        const float yv[4] = {y.x, y.y, y.z, y.w}, gv[4] = {g.x, g.y, g.z, g.w};
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float yh = (yv[q] - st.x) * st.y;
            if (e0 + q < E && yh > 0.f) {
                s1 += gv[q];
                s2 = fmaf(gv[q], yh, s2);
            }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) out[row] = make_float2(s1, s2);
    }
}

// row sums of a shared tile -> partial bias gradient
__device__ __forceinline__ void row_sums(const float* T_s, float* __restrict__ out) {
    if (threadIdx.x < CH) {
        const float* r = T_s + threadIdx.x * LD;
        float s = 0.f;
#pragma unroll 8
        for (int e = 0; e < TE; e += 4) {
            const float4 v = *reinterpret_cast<const float4*>(r + e);
            s += (v.x + v.y) + (v.z + v.w);
        }
        out[threadIdx.x] = s;
    }
}

__device__ __forceinline__ void store_wgrad_partial(const float (&acc)[8][8], float* __restrict__ part) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) part[(ty + 16 * r) * CH + tx + 16 * c] = acc[r][c];
}

// load a [128][TE] tile of a channel-major activation into shared memory, zeroing invalid edges
template <typename F>
__device__ __forceinline__ void load_tile(float* T_s, int tile, int E, F&& f) {
#pragma unroll 4
    for (int it = 0; it < (CH * TE / 4) / MLP_THREADS; ++it) {
        const int id = it * MLP_THREADS + threadIdx.x;
        const int row = id >> 5, c4 = (id & 31) * 4;
        const int e0 = tile * TE + c4;
        float4 v = f(row, e0);
        if (e0 + 0 >= E) v.x = 0.f;
        if (e0 + 1 >= E) v.y = 0.f;
        if (e0 + 2 >= E) v.z = 0.f;
        if (e0 + 3 >= E) v.w = 0.f;
        *reinterpret_cast<float4*>(T_s + row * LD + c4) = v;
    }
}

enum { MODE_R2 = 0, MODE_R3 = 1 };

template <int MODE>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_bwd_kernel(MlpBwdArgs a, int blk) {
    const WsLayout& L = a.L;
    const int tile = blockIdx.x % L.T;
    const int64_t obj = blockIdx.x / L.T;
    const int net = blockIdx.y;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int E = L.E, EP = L.EP;
    const int64_t obj_off = obj * (int64_t)CH * EP;
    const int64_t cta = ((int64_t)net * L.N + obj) * L.T + tile;
    const int64_t ncta = 2 * L.N * L.T;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* D_s = reinterpret_cast<float*>(smem_raw);
    float* H_s = D_s + CH * LD;
    float* Wc_s = H_s + CH * LD;
    float2* st1_s = reinterpret_cast<float2*>(Wc_s + 2 * KC * CH);   // (mean1, inv1)
    float2* st2_s = st1_s + CH;                                      // (mean2, inv2)           [R2]
    float2* sb_s = st2_s + CH;                                       // (S1/E, S2/(E-1)) of the norm being inverted

    const float* Y1 = act_ptr(a.ws, L, net, blk, SLOT_Y1) + obj_off;
    if (tid < CH) {
        st1_s[tid] = merge_cn_stats(stat_ptr(a.ws, L, net, blk, 0) + obj * (int64_t)L.T * CH, tid, L.T, E);
        if (MODE == MODE_R2)
            st2_s[tid] = merge_cn_stats(stat_ptr(a.ws, L, net, blk, 1) + obj * (int64_t)L.T * CH, tid, L.T, E);
        const float2 sb = merge_sums(bstatp(a, net, MODE == MODE_R2 ? 0 : 1, obj), tid, L.T);
        sb_s[tid] = make_float2(sb.x / (float)E, sb.y / (float)(E - 1));
    }
    __syncthreads();

    float acc[8][8];
    if (MODE == MODE_R2) {
        const float* Y2 = act_ptr(a.ws, L, net, blk, SLOT_Y2) + obj_off;
        const float* G = Gp(a, net, obj);
        // D_s = dy2 = inv2 * (dyh - mean(dyh) - yh * sum(dyh*yh)/(E-1)),  dyh = G * (yh > 0)
        load_tile(D_s, tile, E, [&](int row, int e0) {
            const float2 st = st2_s[row], sb = sb_s[row];
            const float4 y = *reinterpret_cast<const float4*>(Y2 + (int64_t)row * EP + e0);
            const float4 g = *reinterpret_cast<const float4*>(G + (int64_t)row * EP + e0);
            auto one = [&](float yv, float gv) {
                const float yh = (yv - st.x) * st.y;
                const float dyh = yh > 0.f ? gv : 0.f;
                return st.y * (dyh - sb.x - yh * sb.y);
            };
            return make_float4(one(y.x, g.x), one(y.y, g.y), one(y.z, g.z), one(y.w, g.w));
        });
        // H_s = yhat1
        load_tile(H_s, tile, E, [&](int row, int e0) {
            const float2 st = st1_s[row];
            const float4 y = *reinterpret_cast<const float4*>(Y1 + (int64_t)row * EP + e0);
            return make_float4((y.x - st.x) * st.y, (y.y - st.x) * st.y, (y.z - st.x) * st.y, (y.w - st.x) * st.y);
        });
        __syncthreads();
        zero_acc(acc);
        tile_wgrad(D_s, H_s, acc);
        float* part = a.wpart + ((int64_t)2 * ncta + cta) * WP;
        store_wgrad_partial(acc, part);
        row_sums(D_s, part + CH * CH);
        // d yhat1 = W2^T dy2
        zero_acc(acc);
        tile_gemm(Wnp(a, net, blk, 2), D_s, Wc_s, acc);
        float* D1 = D1p(a, net, obj);
        float2* bs = bstatp(a, net, 1, obj) + (int64_t)tile * CH;
        const int valid = min(TE, E - tile * TE);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ch = own4(ty, r);
            float* row = D1 + (int64_t)ch * EP + tile * TE;
            *reinterpret_cast<float4*>(row + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *reinterpret_cast<float4*>(row + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int el = own4(tx, q);
                if (el < valid) {
                    s1 += acc[r][q];
                    s2 = fmaf(acc[r][q], H_s[ch * LD + el], s2);
                }
            }
            s1 = half_warp_sum(s1);
            s2 = half_warp_sum(s2);
            if (tx == 0) bs[ch] = make_float2(s1, s2);
        }
    } else {
        const float* D1 = D1p(a, net, obj);
        const float* P = act_ptr(a.ws, L, net, blk, SLOT_P) + obj_off;
        const float* X = act_ptr(a.ws, L, net, blk, SLOT_X) + obj_off;
        float* G = Gp(a, net, obj);
        // D_s = dy1
        load_tile(D_s, tile, E, [&](int row, int e0) {
            const float2 st = st1_s[row], sb = sb_s[row];
            const float4 y = *reinterpret_cast<const float4*>(Y1 + (int64_t)row * EP + e0);
            const float4 d = *reinterpret_cast<const float4*>(D1 + (int64_t)row * EP + e0);
            auto one = [&](float yv, float dv) {
                const float yh = (yv - st.x) * st.y;
                return st.y * (dv - sb.x - yh * sb.y);
            };
            return make_float4(one(y.x, d.x), one(y.y, d.y), one(y.z, d.z), one(y.w, d.w));
        });
        load_tile(H_s, tile, E, [&](int row, int e0) {
            return *reinterpret_cast<const float4*>(P + (int64_t)row * EP + e0);
        });
        __syncthreads();
        zero_acc(acc);
        tile_wgrad(D_s, H_s, acc);                                   // dW1 = dy1 . P^T
        float* part1 = a.wpart + ((int64_t)1 * ncta + cta) * WP;
        store_wgrad_partial(acc, part1);
        row_sums(D_s, part1 + CH * CH);
        zero_acc(acc);
        tile_gemm(Wnp(a, net, blk, 1), D_s, Wc_s, acc);              // dP = W1^T dy1
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ch = own4(ty, r);
            *reinterpret_cast<float4*>(H_s + ch * LD + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *reinterpret_cast<float4*>(H_s + ch * LD + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
        load_tile(D_s, tile, E, [&](int row, int e0) {
            return *reinterpret_cast<const float4*>(X + (int64_t)row * EP + e0);
        });
        __syncthreads();
        zero_acc(acc);
        tile_wgrad(H_s, D_s, acc);                                   // dWp = dP . x^T
        float* part0 = a.wpart + ((int64_t)0 * ncta + cta) * WP;
        store_wgrad_partial(acc, part0);
        row_sums(H_s, part0 + CH * CH);
        zero_acc(acc);
        tile_gemm(Wnp(a, net, blk, 0), H_s, Wc_s, acc);              // dx = Wp^T dP (+ residual path)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ch = own4(ty, r);
            float* row = G + (int64_t)ch * EP + tile * TE;
            float4 g0 = *reinterpret_cast<float4*>(row + tx * 4);
            float4 g1 = *reinterpret_cast<float4*>(row + 64 + tx * 4);
            g0.x += acc[r][0]; g0.y += acc[r][1]; g0.z += acc[r][2]; g0.w += acc[r][3];
            g1.x += acc[r][4]; g1.y += acc[r][5]; g1.z += acc[r][6]; g1.w += acc[r][7];
            *reinterpret_cast<float4*>(row + tx * 4) = g0;
            *reinterpret_cast<float4*>(row + 64 + tx * 4) = g1;
        }
    }
}

// BW4: conv_in gradient partials.  Warp w owns channels 16w..16w+15, lanes own 4 edges of the tile each.
__global__ void __launch_bounds__(256) conv_in_bwd_kernel(MlpBwdArgs a) {
    const WsLayout& L = a.L;
    const int tile = blockIdx.x % L.T;
    const int64_t obj = blockIdx.x / L.T;
    const int net = blockIdx.y;
    const int cin = net == 0 ? 4 : 6;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int E = L.E, EP = L.EP;
    const int64_t cta = ((int64_t)net * L.N + obj) * L.T + tile;
    float f[4][6];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = tile * TE + lane * 4 + q;
        int i, j;
        decode_edge(e < E ? e : E - 1, L.n, i, j);
        if (net == 0) {
            const float* pi = a.kpts2d + (obj * L.n + i) * 2;
            const float* pj = a.kpts2d + (obj * L.n + j) * 2;
            f[q][0] = __ldg(pi); f[q][1] = __ldg(pi + 1); f[q][2] = __ldg(pj); f[q][3] = __ldg(pj + 1);
            f[q][4] = 0.f; f[q][5] = 0.f;
        } else {
            const float* pi = a.kpts3d + (obj * L.n + i) * 3;
            const float* pj = a.kpts3d + (obj * L.n + j) * 3;
            f[q][0] = __ldg(pi); f[q][1] = __ldg(pi + 1); f[q][2] = __ldg(pi + 2);
            f[q][3] = __ldg(pj); f[q][4] = __ldg(pj + 1); f[q][5] = __ldg(pj + 2);
        }
    }
    const float* G = Gp(a, net, obj);
    float* out = a.inpart + cta * (CH * 8);
    const int e0 = tile * TE + lane * 4;
    for (int c = warp * 16; c < warp * 16 + 16; ++c) {
        const float4 g4 = *reinterpret_cast<const float4*>(G + (int64_t)c * EP + e0);
        const float g[4] = {e0 + 0 < E ? g4.x : 0.f, e0 + 1 < E ? g4.y : 0.f, e0 + 2 < E ? g4.z : 0.f, e0 + 3 < E ? g4.w : 0.f};
        float s[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) s[k] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < 6; ++k) s[k] = fmaf(g[q], f[q][k], s[k]);
            s[6] += g[q];
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) s[k] = warp_sum(s[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) out[c * 8 + k] = (k < cin) ? s[k] : 0.f;
            out[c * 8 + 6] = s[6];
            out[c * 8 + 7] = 0.f;
        }
    }
}

// Sum the per-CTA partials of one net in CTA order and write them into the gradient blob
// (weights transposed back to the blob's [in][out] layout).
__global__ void __launch_bounds__(256) reduce_wgrad_kernel(MlpBwdArgs a, int blk, float* __restrict__ g4, float* __restrict__ g6) {
    const WsLayout& L = a.L;
    const int which = blockIdx.y, net = blockIdx.z;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= WP) return;
    const int64_t per_net = L.N * L.T, ncta = 2 * per_net;
    const float* p = a.wpart + ((int64_t)which * ncta + net * per_net) * WP + t;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t i = 0;
    for (; i + 4 <= per_net; i += 4) {
        s0 += p[(i + 0) * WP];
        s1 += p[(i + 1) * WP];
        s2 += p[(i + 2) * WP];
        s3 += p[(i + 3) * WP];
    }
    for (; i < per_net; ++i) s0 += p[i * WP];
    const float s = (s0 + s1) + (s2 + s3);
    const int cin = net == 0 ? 4 : 6;
    float* g = net == 0 ? g4 : g6;
    if (t < CH * CH) {
        const int o = t / CH, in = t % CH;
        g[blob_w(cin, blk, which) + (int64_t)in * CH + o] = s;
    } else {
        g[blob_b(cin, blk, which) + (t - CH * CH)] = s;
    }
}

__global__ void __launch_bounds__(128) reduce_conv_in_kernel(MlpBwdArgs a, float* __restrict__ g4, float* __restrict__ g6) {
    const WsLayout& L = a.L;
    const int net = blockIdx.y, k = blockIdx.x, c = threadIdx.x;     // k in 0..6
    const int cin = net == 0 ? 4 : 6;
    if (k < 6 && k >= cin) return;
    const int64_t per_net = L.N * L.T;
    const float* p = a.inpart + (int64_t)net * per_net * (CH * 8) + c * 8 + k;
    float s = 0.f;
    for (int64_t i = 0; i < per_net; ++i) s += p[i * (CH * 8)];
    float* g = net == 0 ? g4 : g6;
    if (k == 6) g[blob_in_b(cin) + c] = s;
    else g[blob_in_w() + (int64_t)k * CH + c] = s;
}

}  // namespace

struct BwdScratch {
    int64_t Wn, G, D1, bstat, wpart, inpart, total;   // float offsets
};

BwdScratch bwd_scratch_layout(const WsLayout& L) {
    BwdScratch s;
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += (n + 63) & ~(int64_t)63; return r; };
    s.Wn = take((int64_t)2 * L.depth * 3 * CH * CH);
    s.G = take(2 * L.act);
    s.D1 = take(2 * L.act);
    s.bstat = take((int64_t)2 * 2 * L.stat * 2);
    s.wpart = take((int64_t)3 * 2 * L.N * L.T * WP);
    s.inpart = take((int64_t)2 * L.N * L.T * CH * 8);
    s.total = o;
    return s;
}

size_t gmw_bwd_scratch_floats(int64_t N, int n, int depth) {
    return (size_t)bwd_scratch_layout(make_layout(N, n, depth, 1)).total;
}

int launch_gmw_weights_bwd(const float* kpts2d, const float* kpts3d, const float* params4, const float* params6,
                           int64_t N, int n, int depth, const float* grad_reg_w, float* grad4, float* grad6,
                           float* ws, float* scratch, cudaStream_t st) {
    MlpBwdArgs a;
    a.kpts2d = kpts2d; a.kpts3d = kpts3d;
    a.params[0] = params4; a.params[1] = params6;
    a.ws = ws;
    a.L = make_layout(N, n, depth, 1);
    if ((int64_t)a.L.T * N > 0x7fffffffLL) return DCD_E_UNSUPPORTED;
    const BwdScratch S = bwd_scratch_layout(a.L);
    a.Wn = scratch + S.Wn;
    a.G = scratch + S.G;
    a.D1 = scratch + S.D1;
    a.bstat = reinterpret_cast<float2*>(scratch + S.bstat);
    a.wpart = scratch + S.wpart;
    a.inpart = scratch + S.inpart;
    cudaFuncSetAttribute(mlp_bwd_kernel<MODE_R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    cudaFuncSetAttribute(mlp_bwd_kernel<MODE_R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);

    transpose_weights_kernel<<<2 * depth * 3, dim3(32, 8), 0, st>>>(a);
    const unsigned g0 = (unsigned)(((a.L.E + 255) / 256) * N);
    edge_weight_bwd_kernel<<<g0, 256, 0, st>>>(a, grad_reg_w);
    const dim3 grid((unsigned)(a.L.T * N), 2);
    const dim3 rgrid((WP + 255) / 256, 3, 2);
    for (int blk = depth - 1; blk >= 0; --blk) {
        cn2_bwd_sums_kernel<<<grid, 256, 0, st>>>(a, blk);
        mlp_bwd_kernel<MODE_R2><<<grid, MLP_THREADS, kBwdSmem, st>>>(a, blk);
        mlp_bwd_kernel<MODE_R3><<<grid, MLP_THREADS, kBwdSmem, st>>>(a, blk);
        reduce_wgrad_kernel<<<rgrid, 256, 0, st>>>(a, blk, grad4, grad6);
    }
    conv_in_bwd_kernel<<<grid, 256, 0, st>>>(a);
    reduce_conv_in_kernel<<<dim3(7, 2), 128, 0, st>>>(a, grad4, grad6);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

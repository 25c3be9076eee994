// GMW edge-feature MLP forward, layer by layer, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a:
// the training forward (activations saved for the backward) and the inference forward for n > 73; inference at
// n <= 73 runs gmw_mlp_fused.cu instead.
//
// The net is cut at the context norms into segments, each one persistent launch over all (object, 128-edge tile) pairs:
// FIRST conv_in -> Wf, B CN -> conv2, CA CN + ReLU + residual -> Wf, where Wf = W1 . Wp is the block's preconv and conv1
// folded into one layer (nothing sits between them, ops.py:125-131; tc_fold_prep_kernel).  The 128x128x128 layer GEMMs
// run as tcgen05.mma with FP32 accumulation in tensor memory:
//   D[out-channel (TMEM lane) x edge (TMEM column)] = W[out x in] . X[in x edge]
// FP32 fidelity comes from a two-term FP16 split of BOTH operands (x = hi + lo, 11 + 11 significant
// bits) and three MMAs per product:  D = Wh.Xl + Wl.Xh + Wh.Xh   (the dropped Wl.Xl term is 2^-22).
// Weights are pre-scaled per matrix by a power of two so that hi/lo stay in FP16's normal range; the
// epilogue multiplies by the exact inverse.  No tensor maps are needed:
//   * A (weights) is converted once per CTA (FP32 row -> scaled FP16 hi/lo pairs) and stays RESIDENT IN
//     TENSOR MEMORY as the A operand while the persistent CTA loops over its tiles (one CTA per SM,
//     74 per net).  With A in shared memory an M = N = 128 MMA would need all 128 B/clk of shared-memory
//     bandwidth; from tensor memory only B is fetched;
//   * B (activations) is written straight into the canonical MN-major 128B-swizzled layout by the threads
//     that load the tile (context norm + ReLU + residual fused into the load);
//   * the epilogue reads D with tcgen05.ld (32 lanes x 32 columns per warp): a thread owns one output
//     channel, so bias and context-norm statistics need no shuffles.
#include <cstdlib>
#include "gmw_tc_common.cuh"

namespace dcd {

namespace {

enum { MODE_FIRST = 0, MODE_B = 1, MODE_CA = 2 };

constexpr int TC_THREADS = 512;                        // two groups of 8 warps, each with its own tile stream
constexpr int GROUP_THREADS = 256;
// tensor memory map (512 columns): resident weights as the A operand, one accumulator per group
constexpr uint32_t TM_W0_HI = 0, TM_W0_LO = 64, TM_W1_HI = 128, TM_W1_LO = 192, TM_D0 = 256, TM_D1 = 384;
constexpr uint32_t TMEM_COLS = 512;

// smem carve (bytes)
constexpr size_t SM_B = 0;                                       // [group]{hi, lo} : 4 x 32 KB
constexpr size_t SM_STAT = SM_B + 4 * B_PART_BYTES;              // [group][128] float2
constexpr size_t SM_F = SM_STAT + 2 * CH * sizeof(float2);       // [group][128 edges][8] floats (FIRST)
constexpr size_t SM_HALF = SM_F + 2 * TE * 8 * sizeof(float);    // [group][128] float4: statistics of the upper column half
constexpr size_t SM_BAR = SM_HALF + 2 * CH * sizeof(float4);     // mbarriers + tmem pointer
constexpr size_t kTcSmem = SM_BAR + 64;

__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GROUP_THREADS) : "memory");
}

// per-matrix power-of-two scale that puts max|W| in [512, 1024): FP16 hi/lo both stay normal
__global__ void __launch_bounds__(256) tc_weight_scales_kernel(const float* __restrict__ p4, const float* __restrict__ p6,
                                                               int depth, float2* __restrict__ scales) {
    const int m = blockIdx.x;                               // (net, blk, which)
    const int which = m % 3, blk = (m / 3) % depth, net = m / (3 * depth);
    const int cin = net == 0 ? 4 : 6;
    const float* Wt = (net == 0 ? p4 : p6) + blob_w(cin, blk, which);
    __shared__ float red[8];
    float mx = 0.f;
    for (int i = threadIdx.x; i < CH * CH; i += 256) mx = fmaxf(mx, fabsf(Wt[i]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        mx = 0.f;
        for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[w]);
        int e = 0;
        if (mx > 0.f) frexpf(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
        scales[m] = make_float2(ldexpf(1.f, 10 - e), ldexpf(1.f, e - 10));
    }
}

// Folded layer of every block (see FOLD_STRIDE): Wf^T = Wp^T . W1^T and bf = W1 . bp + b1 with FP64 accumulation, rounded
// once to FP32.  grid (8, 2 * depth): block (bx, m) forms rows [16 bx, 16 bx + 16) of matrix m = (net, blk) and raises the
// matrix' running maximum (float bits, >= 0) in wmax[m]; wmax must be zero before the launch.
__global__ void __launch_bounds__(1024) tc_fold_prep_kernel(const float* __restrict__ p4, const float* __restrict__ p6, int depth,
                                                            float* __restrict__ fold, int* __restrict__ wmax) {
    const int m = blockIdx.y;                               // (net, blk)
    const int blk = m % depth, net = m / depth;
    const int cin = net == 0 ? 4 : 6;
    const float* prm = net == 0 ? p4 : p6;
    const float* Wp = prm + blob_w(cin, blk, 0);            // [in][mid]
    const float* W1 = prm + blob_w(cin, blk, 1);            // [mid][out]
    float* out = fold + (int64_t)m * FOLD_STRIDE;
    const int tid = threadIdx.x, o = tid & 127;
    float mx = 0.f;
    for (int i = 16 * blockIdx.x + (tid >> 7); i < 16 * blockIdx.x + 16; i += 8) {
        double acc = 0.0;
#pragma unroll 8
        for (int k = 0; k < CH; ++k) acc = fma((double)Wp[i * CH + k], (double)W1[k * CH + o], acc);
        const float w = (float)acc;
        out[i * CH + o] = w;
        mx = fmaxf(mx, fabsf(w));
    }
    if (blockIdx.x == 0 && tid < CH) {
        const float* bp = prm + blob_b(cin, blk, 0);
        double acc = (double)prm[blob_b(cin, blk, 1) + tid];
        for (int k = 0; k < CH; ++k) acc = fma((double)bp[k], (double)W1[k * CH + tid], acc);
        out[CH * CH + tid] = (float)acc;
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) atomicMax(wmax + m, __float_as_int(mx));
}
// FP16 scale of every folded layer into the scales slot of conv1 (the separate preconv / conv1 scales are unused)
__global__ void tc_fold_scale_kernel(const int* __restrict__ wmax, int nmat, float2* __restrict__ scales) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmat) return;
    const float mx = __int_as_float(wmax[m]);
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    scales[m * 3 + 1] = make_float2(ldexpf(1.f, 10 - e), ldexpf(1.f, e - 10));
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(MlpArgs a, int blk, const float2* __restrict__ scales) {
    const WsLayout& L = a.L;
    const int net = blockIdx.y;
    const int cin = net == 0 ? 4 : 6;
    const float* __restrict__ prm = a.params[net];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (MMA issue on the uniform datapath)
    const int group = warp >> 3, gwarp = warp & 7, gtid = tid & (GROUP_THREADS - 1);
    const int quarter = gwarp & 3, hsel = gwarp >> 2;      // TMEM lane quarter / column half owned in the epilogues
    const int E = L.E, EP = L.EP, T = L.T;

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* B_hi = smem + SM_B + (size_t)group * 2 * B_PART_BYTES;
    unsigned char* B_lo = B_hi + B_PART_BYTES;
    float2* stat_s = reinterpret_cast<float2*>(smem + SM_STAT) + group * CH;
    float* f_s = reinterpret_cast<float*>(smem + SM_F) + group * TE * 8;
    float4* half_s = reinterpret_cast<float4*>(smem + SM_HALF) + group * CH;
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + SM_BAR) + group;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SM_BAR + 16);

    // ---- one-time setup: barriers, tensor memory, resident weights
    const int w0 = (MODE == MODE_B) ? 2 : 0;                // scales slot of this segment's matrix: conv2 | (slot 0 + 1 =) folded layer
    const int mat0 = (net * L.depth + blk) * 3 + w0;
    if (tid == 0) {
        mbar_init(reinterpret_cast<uint64_t*>(smem + SM_BAR), 1);
        mbar_init(reinterpret_cast<uint64_t*>(smem + SM_BAR) + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    const int ch = 32 * quarter + lane;                     // this thread's TMEM lane = output channel
    const uint32_t lane_off = (uint32_t)(32 * quarter) << 16;
    const float2 sc0 = __ldg(scales + mat0);
    const float2 sc1 = (MODE != MODE_B) ? __ldg(scales + mat0 + 1) : make_float2(0.f, 0.f);
    // FIRST / CA run the block's folded preconv.conv1 layer (Wf, bf from the fold area), B runs conv2
    const float* foldp = a.fold + ((int64_t)net * L.depth + blk) * FOLD_STRIDE;
    if (group == 0) {
        if (MODE == MODE_B)
            load_weight_row_to_tmem(prm + blob_w(cin, blk, 2), sc0.x, ch, tmem_base + lane_off + TM_W0_HI, tmem_base + lane_off + TM_W0_LO, hsel);
        else
            load_weight_row_to_tmem(foldp, sc1.x, ch, tmem_base + lane_off + TM_W0_HI, tmem_base + lane_off + TM_W0_LO, hsel);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tmem_d = tmem_base + (group == 0 ? TM_D0 : TM_D1);
    const uint32_t t_lane = tmem_d + lane_off + (uint32_t)(hsel * 64);
    const float un0 = sc0.y, un1 = sc1.y;
    const float bias0 = (MODE == MODE_B) ? __ldg(prm + blob_b(cin, blk, 2) + ch) : 0.f;
    const float bias1 = (MODE != MODE_B) ? __ldg(foldp + CH * CH + ch) : 0.f;

    // contiguous tile range of this (CTA, group): consecutive tiles share the object's statistics
    const int64_t ntiles = L.N * (int64_t)T;
    const int64_t nworkers = (int64_t)gridDim.x * 2;
    const int64_t wid = (int64_t)blockIdx.x * 2 + group;
    const int64_t t_begin = ntiles * wid / nworkers, t_end = ntiles * (wid + 1) / nworkers;

    uint32_t mma_phase = 0;
    int64_t stat_obj = -1;
    // Source-tile mapping: a warp owns 16 channels = 8 row pairs (c, c+1); per pair lanes 0-15 take row c and lanes
    // 16-31 row c+1, each lane one 256-bit load of the 8 edges of block (lane & 15): two fully used 512-byte rows
    // per request, and the lane already holds exactly one 16-byte chunk of the FP16 operand (no shuffles).
    // Prefetch: the first PRE row pairs of the NEXT tile are loaded into registers right after the current tile's
    // operand is handed to the tensor core, so their HBM latency is covered by the MMA waits and both epilogues.
    constexpr int NPAIR = 8;                                // row pairs per warp (16 channels)
    constexpr int PRE = (MODE == MODE_FIRST) ? 0 : ((MODE == MODE_CA) ? 2 : 8);
    F8 ypre[PRE > 0 ? PRE : 1], xpre[(MODE == MODE_CA) ? PRE : 1];
    const int eblk = lane & 15;                             // edge block (8 edges) this lane converts
    auto lane_row = [&](int it) { return gwarp * 16 + it * 2 + (lane >> 4); };
    auto prefetch = [&](int64_t tt) {
        if (PRE == 0) return;
        const int64_t o = tt / T;
        const int tl = (int)(tt - o * T);
        const int pb = (MODE == MODE_B) ? blk : blk - 1;
        const float* Y = act_ptr(a.ws, L, net, pb, MODE == MODE_B ? SLOT_Y1 : SLOT_Y2) + o * (int64_t)CH * EP + tl * TE + eblk * 8;
        const float* Xp = (MODE == MODE_CA) ? act_ptr(a.ws, L, net, blk - 1, SLOT_X) + o * (int64_t)CH * EP + tl * TE + eblk * 8 : nullptr;
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
            const int c = lane_row(u);
            ypre[u] = ld256(Y + (int64_t)c * EP);
            if (MODE == MODE_CA) xpre[u] = ld256(Xp + (int64_t)c * EP);
        }
    };
    if (t_begin < t_end) prefetch(t_begin);
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int64_t obj = t / T;
        const int tile = (int)(t - obj * T);
        const int64_t obj_off = obj * (int64_t)CH * EP;
        const int valid = min(TE, E - tile * TE);

        // ---- source tile -> B operand (hi/lo)
        if (MODE == MODE_FIRST) {
            if (gtid < TE) {
                const int e = tile * TE + gtid;
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
                for (int r = 0; r < 6; ++r) f_s[r * TE + gtid] = f[r];     // feature-major: conflict-free both ways
            }
            group_sync(group);
        } else if (stat_obj != obj) {
            const int pb = (MODE == MODE_B) ? blk : blk - 1;
            const int which = (MODE == MODE_B) ? 0 : 1;
            if (gtid < CH)
                stat_s[gtid] = merge_cn_stats(stat_ptr(a.ws, L, net, pb, which) + obj * (int64_t)T * CH, gtid, T, E);
            group_sync(group);
            stat_obj = obj;
        }
        {
            const int pb = (MODE == MODE_B) ? blk : blk - 1;
            const float* Y = (MODE == MODE_FIRST) ? nullptr : act_ptr(a.ws, L, net, pb, MODE == MODE_B ? SLOT_Y1 : SLOT_Y2) + obj_off;
            const float* Xp = (MODE == MODE_CA) ? act_ptr(a.ws, L, net, blk - 1, SLOT_X) + obj_off : nullptr;
            float* Xn = (MODE == MODE_B) ? nullptr : act_ptr(a.ws, L, net, MODE == MODE_FIRST ? 0 : blk, SLOT_X) + obj_off;
            constexpr int UNR = (MODE == MODE_CA) ? 2 : 8;
            const bool full_tile = valid == TE;
            float fe[6][8];                                  // FIRST: the 6 edge features of this lane's 8 edges
            if (MODE == MODE_FIRST) {
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const float4 f0 = *reinterpret_cast<const float4*>(f_s + r * TE + eblk * 8);
                    const float4 f1 = *reinterpret_cast<const float4*>(f_s + r * TE + eblk * 8 + 4);
                    fe[r][0] = f0.x; fe[r][1] = f0.y; fe[r][2] = f0.z; fe[r][3] = f0.w;
                    fe[r][4] = f1.x; fe[r][5] = f1.y; fe[r][6] = f1.z; fe[r][7] = f1.w;
                }
            }
            // one operand chunk: 8 edges of channel c
            auto convert = [&](int it, const F8& yy, const F8& xx) {
                const int c = lane_row(it);
                float v[8];
                if (MODE == MODE_FIRST) {
                    const float b = __ldg(prm + blob_in_b(cin) + c);
                    float wq[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) wq[q] = (q < cin) ? __ldg(prm + blob_in_w() + q * CH + c) : 0.f;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float x = b;
#pragma unroll
                        for (int r = 0; r < 6; ++r) x = fmaf(wq[r], fe[r][q], x);
                        v[q] = x;
                    }
                } else {
                    const float2 st = stat_s[c];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = (yy.v[q] - st.x) * st.y;
                    if (MODE == MODE_CA) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f) + xx.v[q];
                    }
                }
                if (!full_tile) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (eblk * 8 + q >= valid) v[q] = 0.f;
                }
                if (MODE != MODE_B) st256(Xn + (int64_t)c * EP + tile * TE + eblk * 8, v);   // residual stream
                store_b8(B_hi, B_lo, c, eblk, v);
            };
            // row pairs not covered by the prefetch are loaded here; the prefetched ones are converted while they fly
            F8 ybuf[UNR], xbuf[(MODE == MODE_CA) ? UNR : 1];
#pragma unroll 1
            for (int it0 = PRE; it0 < NPAIR; it0 += UNR) {
                if (MODE != MODE_FIRST) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int c = lane_row(it0 + u);
                        ybuf[u] = ld256(Y + (int64_t)c * EP + tile * TE + eblk * 8);
                        if (MODE == MODE_CA) xbuf[u] = ld256(Xp + (int64_t)c * EP + tile * TE + eblk * 8);
                    }
                }
                if (it0 == PRE) {
#pragma unroll
                    for (int u = 0; u < PRE; ++u) convert(u, ypre[u], xpre[MODE == MODE_CA ? u : 0]);
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) convert(it0 + u, ybuf[u], xbuf[MODE == MODE_CA ? u : 0]);
            }
            if (PRE == NPAIR) {                              // everything came from the prefetch
#pragma unroll
                for (int u = 0; u < PRE; ++u) convert(u, ypre[u], xpre[0]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);
        // ---- GEMM 1
        if (gwarp == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue_layer_gemm(tmem_d, tmem_base + TM_W0_HI, tmem_base + TM_W0_LO, smem_u32(B_hi), smem_u32(B_lo));
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        if (t + 1 < t_end) prefetch(t + 1);
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        // ---- final epilogue of the segment: + bias, tile statistics (thread = channel, all 128 edges), then a
        //      coalesced store: the FP32 tile is staged in this group's (now idle) operand buffer, one 512-byte
        //      row per channel with the 16-byte slots XOR-swizzled by the row so that both the row-owner writes
        //      and the row-wise reads are bank-conflict free.
        {
            const float un = (MODE == MODE_B) ? un0 : un1;
            const float bias = (MODE == MODE_B) ? bias0 : bias1;
            float4* stage = reinterpret_cast<float4*>(B_hi);            // [128 rows][32 slots] = 64 KB (B_hi + B_lo)
            float mean = 0.f, M2 = 0.f, cnt = 0.f;
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                float v[32];
                tmem_ld32(t_lane + part * 32, v);
                const int nv = max(0, min(32, valid - hsel * 64 - part * 32));
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], un, bias);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    stage[ch * 32 + ((hsel * 16 + part * 8 + q) ^ (ch & 31))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                if (nv > 0) {
                    float pm, pm2 = 0.f;
                    if (nv == 32) {                          // full part (all but the object's last tile): no predication
                        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < 32; ++i) s4[i & 3] += v[i];
                        pm = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 32.0f);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float d = v[i] - pm;
                            pm2 = fmaf(d, d, pm2);
                        }
                    } else {
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nv) sum += v[i];
                        pm = sum / (float)nv;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nv) {
                                const float d = v[i] - pm;
                                pm2 = fmaf(d, d, pm2);
                            }
                    }
                    const float nb = (float)nv, tot = cnt + nb;       // Chan merge of the 32-edge parts
                    const float delta = pm - mean;
                    mean += delta * (nb / tot);
                    M2 += pm2 + delta * delta * (cnt * nb / tot);
                    cnt = tot;
                }
            }
            if (hsel == 1) half_s[ch] = make_float4(mean, M2, cnt, 0.f);
            tc_fence_before();
            group_sync(group);                                // staging complete; TMEM reads done
            if (hsel == 0) {
                const float4 o = half_s[ch];                  // Chan merge with the upper 64-edge half
                if (o.z > 0.f) {
                    const float tot = cnt + o.z, delta = o.x - mean;
                    mean += delta * (o.z / tot);
                    M2 += o.y + delta * delta * (cnt * o.z / tot);
                }
                stat_ptr(a.ws, L, net, blk, MODE == MODE_B ? 1 : 0)[(obj * T + tile) * (int64_t)CH + ch] = make_float2(mean, M2);
            }
            float* Yo = act_ptr(a.ws, L, net, blk, MODE == MODE_B ? SLOT_Y2 : SLOT_Y1) + obj_off + tile * TE;
#pragma unroll 8
            for (int i = 0; i < 16; ++i) {
                const int r = gwarp * 16 + i;
                const float4 val = stage[r * 32 + (lane ^ (r & 31))];
                *reinterpret_cast<float4*>(Yo + (int64_t)r * EP + lane * 4) = val;
            }
            group_sync(group);                                // staging buffer free for the next tile's operand
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Final features of both nets -> reg_weights.  FINAL: SLOT_X already holds the final features (fused forward);
// otherwise the last block's context norm + ReLU + residual is applied on the fly.
template <bool FINAL>
__global__ void __launch_bounds__(256) gmw_edge_weight_kernel(MlpArgs a, float* __restrict__ reg_w,
                                                              float* __restrict__ feat4, float* __restrict__ feat6) {
    const WsLayout& L = a.L;
    const int E = L.E, EP = L.EP, last = L.depth - 1;
    const int nb = (E + 255) / 256;
    const int64_t obj = blockIdx.x / nb;
    const int e = (blockIdx.x % nb) * 256 + threadIdx.x;
    __shared__ float2 stat_s[2][CH];
    if (!FINAL) {
        const int net = threadIdx.x >> 7, c = threadIdx.x & 127;
        stat_s[net][c] = merge_cn_stats(stat_ptr(a.ws, L, net, last, 1) + obj * (int64_t)L.T * CH, c, L.T, E);
        __syncthreads();
    }
    if (e >= E) return;
    const int64_t off = obj * (int64_t)CH * EP + e;
    const float* Y4 = act_ptr(a.ws, L, 0, last, SLOT_Y2) + off;
    const float* X4 = act_ptr(a.ws, L, 0, FINAL ? 0 : last, SLOT_X) + off;
    const float* Y6 = act_ptr(a.ws, L, 1, last, SLOT_Y2) + off;
    const float* X6 = act_ptr(a.ws, L, 1, FINAL ? 0 : last, SLOT_X) + off;
    auto feature = [&](const float* Y, const float* X, const float2 s, int c) {
        if (FINAL) return X[(int64_t)c * EP];
        return fmaxf((Y[(int64_t)c * EP] - s.x) * s.y, 0.f) + X[(int64_t)c * EP];
    };
    if (FINAL) {
        // Same arithmetic, in the same order, as the paired epilogue of mlp_fused_kernel (gmw_mlp_fused.cu): the three channel
        // sums |a|^2, |c|^2, a.c per lane quarter by the halving tree (c, c^16), (.., c^8), ... then the four quarters, and the
        // weight from the sums — so a batch gives bit-identical weights whichever of the two schedules its chunks take.
        float saa = 0.f, scc = 0.f, sac = 0.f;
        float qa[4], qc[4], qx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float ta[16], tc[16], tx[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int c0 = 32 * q + k, c1 = c0 + 16;
                const float a0 = X4[(int64_t)c0 * EP], a1 = X4[(int64_t)c1 * EP], b0 = X6[(int64_t)c0 * EP], b1 = X6[(int64_t)c1 * EP];
                if (feat4 != nullptr) { feat4[(obj * CH + c0) * (int64_t)E + e] = a0; feat4[(obj * CH + c1) * (int64_t)E + e] = a1; }
                if (feat6 != nullptr) { feat6[(obj * CH + c0) * (int64_t)E + e] = b0; feat6[(obj * CH + c1) * (int64_t)E + e] = b1; }
                ta[k] = __fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1));
                tc[k] = __fadd_rn(__fmul_rn(b0, b0), __fmul_rn(b1, b1));
                tx[k] = __fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
            }
#pragma unroll
            for (int h = 8; h >= 1; h >>= 1)
#pragma unroll
                for (int k = 0; k < h; ++k) {
                    ta[k] = __fadd_rn(ta[k], ta[k + h]);
                    tc[k] = __fadd_rn(tc[k], tc[k + h]);
                    tx[k] = __fadd_rn(tx[k], tx[k + h]);
                }
            qa[q] = ta[0]; qc[q] = tc[0]; qx[q] = tx[0];
        }
        saa = __fadd_rn(__fadd_rn(qa[0], qa[1]), __fadd_rn(qa[2], qa[3]));
        scc = __fadd_rn(__fadd_rn(qc[0], qc[1]), __fadd_rn(qc[2], qc[3]));
        sac = __fadd_rn(__fadd_rn(qx[0], qx[1]), __fadd_rn(qx[2], qx[3]));
        const float n4 = fmaxf(sqrtf(saa), 1e-12f), n6 = fmaxf(sqrtf(scc), 1e-12f);
        const float a2 = __fdiv_rn(saa, __fmul_rn(n4, n4)), c2 = __fdiv_rn(scc, __fmul_rn(n6, n6)), acn = __fdiv_rn(sac, __fmul_rn(n4, n6));
        const float s2 = __fadd_rn(__fadd_rn(c2, -2.f * acn), a2);
        reg_w[obj * (int64_t)E + e] = __fdiv_rn(1.f, sqrtf(fmaxf(s2, 1e-30f)));
        return;
    }
    float n4 = 0.f, n6 = 0.f;
#pragma unroll 4
    for (int c = 0; c < CH; ++c) {
        const float2 s4 = stat_s[0][c], s6 = stat_s[1][c];
        const float x4 = feature(Y4, X4, s4, c);
        const float x6 = feature(Y6, X6, s6, c);
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
        const float x4 = feature(Y4, X4, s4, c);
        const float x6 = feature(Y6, X6, s6, c);
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

// bytes appended to the MLP workspace for the per-matrix FP16 scales (scale, 1/scale)
bool gmw_fused_supported(int n);
size_t gmw_fused_image_bytes(int depth);
int launch_gmw_fused_fwd(const MlpArgs& a, const float* params4, const float* params6, void* tail, float* reg_w, bool* emitted,
                         cudaStream_t st);

static size_t tc_scales_bytes(int depth) { return (((size_t)2 * depth * 3 * sizeof(float2)) + 255) / 256 * 256; }
static size_t tc_fold_bytes(int depth) {                  // folded layers + their running maxima (one int per matrix)
    return (((size_t)2 * depth * FOLD_STRIDE * sizeof(float)) + 255) / 256 * 256 + (((size_t)2 * depth * sizeof(int)) + 255) / 256 * 256;
}
// forms the folded layers in the workspace tail (used by the layer-wise kernels and, as input, by the fused forward's prep)
float* launch_fold_prep(const float* params4, const float* params6, int depth, float2* scales, bool want_scales, cudaStream_t st) {
    float* fold = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scales) + tc_scales_bytes(depth) + gmw_fused_image_bytes(depth));
    int* wmax = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(fold) + (((size_t)2 * depth * FOLD_STRIDE * sizeof(float)) + 255) / 256 * 256);
    cudaMemsetAsync(wmax, 0, (size_t)2 * depth * sizeof(int), st);
    tc_fold_prep_kernel<<<dim3(8, 2 * depth), 1024, 0, st>>>(params4, params6, depth, fold, wmax);
    if (want_scales) tc_fold_scale_kernel<<<1, 64, 0, st>>>(wmax, 2 * depth, scales);
    return fold;
}
// byte offset (from the scales) of the folded layers of the layer-wise kernels
size_t tc_fold_offset_bytes(int depth) { return tc_scales_bytes(depth) + gmw_fused_image_bytes(depth); }
// bytes appended to the MLP workspace: per-matrix FP16 scales (scale, 1/scale), the tail of the fused forward (its
// own scales, biases, pre-split weight image, exchange buffer), the folded layers of the layer-wise kernels
size_t tc_weight_image_bytes(int depth) { return tc_fold_offset_bytes(depth) + tc_fold_bytes(depth); }

// DCD_B200_LAYERWISE=1 forces the layer-wise kernels also for inference (A/B measurements, cross-checks)
static bool force_layerwise() {
    const char* e = getenv("DCD_B200_LAYERWISE");           // read on every call: no cached library state
    return e != nullptr && e[0] == '1';
}

int launch_gmw_weights_fwd(const float* kpts2d, const float* kpts3d, const float* params4, const float* params6,
                           int64_t N, int n, int depth, int save, float* reg_w, float* feat4, float* feat6,
                           float* ws, cudaStream_t st) {
    MlpArgs a;
    a.kpts2d = kpts2d; a.kpts3d = kpts3d;
    a.params[0] = params4; a.params[1] = params6;
    a.ws = ws;
    a.L = make_layout(N, n, depth, save);
    a.fold = nullptr;
    if ((int64_t)a.L.T * N > 0x7fffffffLL) return DCD_E_UNSUPPORTED;
    // the scales live right after the layout's own area (256-byte aligned)
    float2* scales = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(ws) +
                                               (((size_t)a.L.total * sizeof(float) + 255) / 256) * 256);
    const unsigned g2 = (unsigned)(((a.L.E + 255) / 256) * N);
    if (!save && gmw_fused_supported(n) && !force_layerwise()) {
        // inference: whole network on chip, one kernel (gmw_mlp_fused.cu)
        a.fold = launch_fold_prep(params4, params6, depth, scales, false, st);
        // the edge weights come out of the fused kernel itself unless the final features are wanted too
        bool emitted = false;
        const int rc = launch_gmw_fused_fwd(a, params4, params6, reinterpret_cast<unsigned char*>(scales) + tc_scales_bytes(depth),
                                            (feat4 == nullptr && feat6 == nullptr) ? reg_w : nullptr, &emitted, st);
        if (rc != DCD_OK) return rc;
        if (!emitted) gmw_edge_weight_kernel<true><<<g2, 256, 0, st>>>(a, reg_w, feat4, feat6);
        DCD_CHECK_LAUNCH();
        return DCD_OK;
    }
    tc_weight_scales_kernel<<<2 * depth * 3, 256, 0, st>>>(params4, params6, depth, scales);
    a.fold = launch_fold_prep(params4, params6, depth, scales, true, st);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    cudaFuncSetAttribute(mlp_tc_kernel<MODE_CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    const int64_t ntiles = a.L.T * N;
    const int per_net = max(1, device_sm_count() / 2);
    const int64_t want = (ntiles + 1) / 2;                  // two tile streams (groups) per CTA
    const dim3 grid((unsigned)(want < per_net ? want : per_net), 2);
    mlp_tc_kernel<MODE_FIRST><<<grid, TC_THREADS, kTcSmem, st>>>(a, 0, scales);
    for (int blk = 0; blk < depth; ++blk) {
        mlp_tc_kernel<MODE_B><<<grid, TC_THREADS, kTcSmem, st>>>(a, blk, scales);
        if (blk + 1 < depth) mlp_tc_kernel<MODE_CA><<<grid, TC_THREADS, kTcSmem, st>>>(a, blk + 1, scales);
    }
    DCD_CHECK_LAUNCH();
    gmw_edge_weight_kernel<false><<<g2, 256, 0, st>>>(a, reg_w, feat4, feat6);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

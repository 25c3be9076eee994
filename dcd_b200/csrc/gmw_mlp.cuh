// Shared definitions of the GMW edge-MLP kernels: workspace layout, parameter-blob layout.
#pragma once
#include "dcd_common.cuh"

namespace dcd {

constexpr int CH = DCD_NET_CH;   // 128 channels
constexpr int TE = 128;          // edges per tile

// ---- parameter blob (per net) : W_in^T [Cin][128], b_in[128], then per block {Wp^T,bp,W1^T,b1,W2^T,b2}
__host__ __device__ inline int64_t blob_in_w() { return 0; }
__host__ __device__ inline int64_t blob_in_b(int cin) { return (int64_t)cin * CH; }
__host__ __device__ inline int64_t blob_block(int cin, int blk) {
    return (int64_t)cin * CH + CH + (int64_t)blk * 3 * (CH * CH + CH);
}
// which: 0 = preconv, 1 = conv1, 2 = conv2
__host__ __device__ inline int64_t blob_w(int cin, int blk, int which) {
    return blob_block(cin, blk) + (int64_t)which * (CH * CH + CH);
}
__host__ __device__ inline int64_t blob_b(int cin, int blk, int which) { return blob_w(cin, blk, which) + CH * CH; }
__host__ __device__ inline int64_t blob_size(int cin, int depth) { return blob_block(cin, depth); }

// ---- workspace: activations [net][slot][N][128][EP] (channel-major, EP = tiles*128) followed by the
//      context-norm partial statistics [net][blk][2][N][T][128] float2 (tile mean, tile M2)
struct WsLayout {
    int64_t N;
    int n, E, T, EP, depth, save, slots;
    int64_t act;        // floats per activation buffer
    int64_t stat;       // float2 per statistics buffer
    int64_t stats_off;  // float offset of the statistics area
    int64_t total;      // floats
};

__host__ __device__ inline WsLayout make_layout(int64_t N, int n, int depth, int save) {
    WsLayout L;
    L.N = N; L.n = n; L.depth = depth; L.save = save;
    L.E = n * (n - 1) / 2;
    L.T = (L.E + TE - 1) / TE;
    L.EP = L.T * TE;
    L.slots = save ? 4 * depth : 3;
    L.act = N * (int64_t)CH * L.EP;
    L.stat = N * (int64_t)L.T * CH;
    L.stats_off = 2 * (int64_t)L.slots * L.act;
    L.total = L.stats_off + 2 * (int64_t)depth * 2 * L.stat * 2;
    return L;
}
enum Slot { SLOT_X = 0, SLOT_P = 1, SLOT_Y1 = 2, SLOT_Y2 = 3 };
__host__ __device__ inline int slot_index(const WsLayout& L, int blk, int s) {
    if (L.save) return 4 * blk + s;
    return s == SLOT_X ? 0 : (s == SLOT_Y1 ? 1 : 2);   // SLOT_P is not stored in inference mode
}
__host__ __device__ inline float* act_ptr(float* ws, const WsLayout& L, int net, int blk, int s) {
    return ws + ((int64_t)net * L.slots + slot_index(L, blk, s)) * L.act;
}
__host__ __device__ inline float2* stat_ptr(float* ws, const WsLayout& L, int net, int blk, int which) {
    return reinterpret_cast<float2*>(ws + L.stats_off) + (((int64_t)net * L.depth + blk) * 2 + which) * L.stat;
}

// Folded layer of a block: there is no non-linearity between preconv and conv1 (ops.py:125-131), so the layer-wise
// kernels run  Wf = W1 . Wp,  bf = W1 . bp + b1  as ONE GEMM.  Fold area per (net, block): Wf^T [in][out], then bf [128].
constexpr int64_t FOLD_STRIDE = (int64_t)CH * CH + CH;

// arguments shared by the forward kernels
struct MlpArgs {
    const float* kpts2d;
    const float* kpts3d;
    const float* params[2];
    float* ws;
    WsLayout L;
    const float* fold;     // [2][depth] folded layers (FOLD_STRIDE floats each), written by tc_fold_prep_kernel
};

}  // namespace dcd

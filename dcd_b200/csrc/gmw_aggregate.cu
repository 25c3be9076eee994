// GMW weighted depth aggregation (forward + backward) for sm_100a.
//
// Replaces compute_reg_loss's gather + softmax + weighted sum (GMW/main.py:364-371) and its autograd.
// One warp per object: k (=1500) selected edges are gathered through idx, the softmax follows torch's
// three passes (max, sum of exp(x-max), exp(x-max)/sum) and every reduction is a warp-shuffle tree.
#include "dcd_common.cuh"

namespace dcd {
namespace {

constexpr int AGG_THREADS = 256;

struct Softmax {
    float mx, sum;
};

// an index outside [0, E) must not fault (torch.gather raises there; the Python wrapper validates on request)
__device__ __forceinline__ int64_t edge_id(const int64_t* __restrict__ id_row, int r, int64_t E) {
    const int64_t id = id_row[r];
    return id < 0 ? 0 : (id >= E ? E - 1 : id);
}

__device__ __forceinline__ Softmax warp_softmax_stats(const float* __restrict__ w_row,
                                                      const int64_t* __restrict__ id_row, int k, int64_t E, int lane) {
    float mx = -INFINITY;
    for (int r = lane; r < k; r += 32) mx = fmaxf(mx, __ldg(w_row + edge_id(id_row, r, E)));
    mx = warp_max(mx);
    float s = 0.f;
    for (int r = lane; r < k; r += 32) s += expf(__ldg(w_row + edge_id(id_row, r, E)) - mx);
    s = warp_sum(s);
    return {mx, s};
}

__global__ void __launch_bounds__(AGG_THREADS)
gmw_aggregate_fwd_kernel(const float* __restrict__ reg_w, const float* __restrict__ depths,
                         const int64_t* __restrict__ idx, int64_t N, int64_t E, int k, int sel,
                         float* __restrict__ depth_out, float* __restrict__ probs) {
    const int lane = threadIdx.x & 31;
    const int64_t obj = (int64_t)blockIdx.x * (AGG_THREADS / 32) + (threadIdx.x >> 5);
    if (obj >= N) return;
    const float* w_row = reg_w + obj * E;
    const int64_t* id_row = idx + obj * k;
    const float* z_row = depths + obj * (sel ? (int64_t)k : E);
    const Softmax sm = warp_softmax_stats(w_row, id_row, k, E, lane);
    float acc = 0.f;
    for (int r = lane; r < k; r += 32) {
        const int64_t id = edge_id(id_row, r, E);
        const float p = __fdiv_rn(expf(__ldg(w_row + id) - sm.mx), sm.sum);
        const float z = sel ? __ldg(z_row + r) : __ldg(z_row + id);
        if (probs != nullptr) probs[obj * k + r] = p;
        acc += z * p;
    }
    acc = warp_sum(acc);
    if (lane == 0) depth_out[obj] = acc;
}

__global__ void __launch_bounds__(AGG_THREADS)
gmw_aggregate_bwd_kernel(const float* __restrict__ reg_w, const float* __restrict__ depths,
                         const int64_t* __restrict__ idx, int64_t N, int64_t E, int k, int sel,
                         const float* __restrict__ grad_out, float* __restrict__ grad_w,
                         float* __restrict__ grad_z) {
    const int lane = threadIdx.x & 31;
    const int64_t obj = (int64_t)blockIdx.x * (AGG_THREADS / 32) + (threadIdx.x >> 5);
    if (obj >= N) return;
    const float* w_row = reg_w + obj * E;
    const int64_t* id_row = idx + obj * k;
    const float* z_row = depths + obj * (sel ? (int64_t)k : E);
    float* gw_row = grad_w + obj * E;
    for (int64_t e = lane; e < E; e += 32) gw_row[e] = 0.f;
    if (grad_z != nullptr && !sel)
        for (int64_t e = lane; e < E; e += 32) grad_z[obj * E + e] = 0.f;
    const Softmax sm = warp_softmax_stats(w_row, id_row, k, E, lane);
    float acc = 0.f;
    for (int r = lane; r < k; r += 32) {
        const int64_t id = edge_id(id_row, r, E);
        const float p = __fdiv_rn(expf(__ldg(w_row + id) - sm.mx), sm.sum);
        const float z = sel ? __ldg(z_row + r) : __ldg(z_row + id);
        acc += z * p;
    }
    const float Zbar = warp_sum(acc);
    const float g = __ldg(grad_out + obj);
    __syncwarp();   // orders the zero fill before the scatter
    for (int r = lane; r < k; r += 32) {
        const int64_t id = edge_id(id_row, r, E);
        const float p = __fdiv_rn(expf(__ldg(w_row + id) - sm.mx), sm.sum);
        const float z = sel ? __ldg(z_row + r) : __ldg(z_row + id);
        gw_row[id] = g * p * (z - Zbar);
        if (grad_z != nullptr) {
            if (sel) grad_z[obj * k + r] = g * p;
            else grad_z[obj * E + id] = g * p;
        }
    }
}

}  // namespace

int launch_gmw_aggregate_fwd(const float* reg_w, const float* depths, const int64_t* idx, int64_t N, int64_t E,
                             int k, int sel, float* depth_out, float* probs, cudaStream_t st) {
    const int64_t per = AGG_THREADS / 32;
    const unsigned grid = (unsigned)((N + per - 1) / per);
    gmw_aggregate_fwd_kernel<<<grid, AGG_THREADS, 0, st>>>(reg_w, depths, idx, N, E, k, sel, depth_out, probs);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

int launch_gmw_aggregate_bwd(const float* reg_w, const float* depths, const int64_t* idx, int64_t N, int64_t E,
                             int k, int sel, const float* grad_out, float* grad_w, float* grad_z, cudaStream_t st) {
    const int64_t per = AGG_THREADS / 32;
    const unsigned grid = (unsigned)((N + per - 1) / per);
    gmw_aggregate_bwd_kernel<<<grid, AGG_THREADS, 0, st>>>(reg_w, depths, idx, N, E, k, sel, grad_out, grad_w, grad_z);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

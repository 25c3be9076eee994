// Statistics helpers of the context normalisation shared by the edge-MLP kernels (forward and backward).
#pragma once
#include "gmw_mlp.cuh"

namespace dcd {

// Merge the per-tile (mean, M2) partials of one (object, channel) with Chan's formula and return
// (mean, 1/sqrt(unbiased var + 1e-3))  — GMW/model/yi2018cvpr/ops.py:12-19.
__device__ __forceinline__ float2 merge_cn_stats(const float2* __restrict__ part, int c, int T, int E) {
    float na = 0.f, mean = 0.f, M2 = 0.f;
    for (int t = 0; t < T; ++t) {
        const float cnt = (float)min(TE, E - t * TE);
        const float2 p = part[(int64_t)t * CH + c];
        const float delta = p.x - mean;
        const float tot = na + cnt;
        mean += delta * (cnt / tot);
        M2 += p.y + delta * delta * (na * cnt / tot);
        na = tot;
    }
    const float var = M2 / (float)(E - 1);
    return make_float2(mean, 1.0f / sqrtf(var + 1e-3f));
}

// Sum the per-tile partial sums (s1, s2) of one (object, channel) in tile order.
__device__ __forceinline__ float2 merge_sums(const float2* __restrict__ part, int c, int T) {
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < T; ++t) {
        const float2 p = part[(int64_t)t * CH + c];
        s1 += p.x;
        s2 += p.y;
    }
    return make_float2(s1, s2);
}

}  // namespace dcd

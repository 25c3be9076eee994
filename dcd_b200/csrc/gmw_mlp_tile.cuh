// Tile primitives of the edge MLP: a 128(out) x 128(edge) x 128(k) FP32 GEMM tile per CTA with the weight
// matrix streamed through a double-buffered cp.async pipeline, plus the statistics helpers of the
// context normalisation.  256 threads, thread (ty,tx) = (tid/16, tid%16).
#pragma once
#include "gmw_mlp.cuh"

namespace dcd {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
}

// Row (output channel) and column (edge in tile) owned by accumulator index q in 0..7 for the
// forward-type mapping: {t*4 .. t*4+3, 64 + t*4 .. 64 + t*4+3}.
__device__ __forceinline__ int own4(int t, int q) { return (q < 4 ? 0 : 60) + t * 4 + q; }

// acc[r][c] += sum_k Wt[k][own4(ty,r)] * In_s[k][own4(tx,c)]
// Wt: global [128][128] (k-major rows, contiguous).  In_s: shared [128][LD].  Wc_s: shared [2][KC][128].
// Starts with a __syncthreads (makes In_s written by the caller visible) and ends with one.
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Wt, const float* In_s, float* Wc_s,
                                          float (&acc)[8][8]) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    auto load_chunk = [&](int c, int b) {
        const float* src = Wt + c * KC * CH;
        float* dst = Wc_s + b * KC * CH;
        cp_async16(dst + tid * 4, src + tid * 4);
        cp_async16(dst + 1024 + tid * 4, src + 1024 + tid * 4);
        cp_async_commit();
    };
    load_chunk(0, 0);
#pragma unroll 1
    for (int c = 0; c < CH / KC; ++c) {
        if (c + 1 < CH / KC) {
            load_chunk(c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* W = Wc_s + (c & 1) * KC * CH;
        const float* I = In_s + c * KC * LD;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(W + kk * CH + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(W + kk * CH + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(I + kk * LD + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(I + kk * LD + 64 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[r][q] = fmaf(a[r], b[q], acc[r][q]);
        }
        __syncthreads();
    }
}

// Weight-gradient tile: acc[r][c] += sum_e A_s[ty + 16 r][e] * B_s[tx + 16 c][e]   (e over the 128 tile edges)
// The interleaved row ownership keeps the 128-bit shared loads conflict-free with LD = 132.
__device__ __forceinline__ void tile_wgrad(const float* A_s, const float* B_s, float (&acc)[8][8]) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll 2
    for (int e = 0; e < TE; e += 4) {
        float4 a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(A_s + (ty + 16 * r) * LD + e);
#pragma unroll
        for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const float4*>(B_s + (tx + 16 * c) * LD + e);
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
                acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
                acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
                acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
            }
    }
}

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

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dcd

// Edge selection for sm_100a: the k (= 1500) edges of an object with the largest |v_i - v_j|, sorted by
// (|V| descending, edge id ascending), with their depths / pair masks / mean.
//
// Replaces torch.topk + gather of the training branch of decode_pairs_kpts_depth (DGDE/model/anno_encoder.py:377-382)
// and of compute_z (GMW/main.py:413-414).  One CTA per object, three phases, everything in shared memory / registers:
//   1. keys: the bit patterns of |V| for all E edges (non-negative floats order like unsigned integers; a NaN key
//      sorts first, as in torch.topk), rows of the upper triangle spread over the warps, lanes along j;
//   2. radix select: four 8-bit histogram passes (MSB first) find the k-th largest key T and how many edges
//      equal to T belong to the selection; an ORDERED compaction (warp ballots, edge id order) collects the
//      edges above T and the first ties at T as (key, ~id) pairs  — so only k <= 2048 of the E
//      (2628 ... 32 640) candidates are ever sorted;
//   3. bitonic sort of the 2048-padded composites with 8 elements per thread IN REGISTERS: strides 1-4 are register
//      compare-exchanges, strides 8-128 warp shuffles, strides 256-1024 register compare-exchanges again after a
//      transposition through (bank-padded) shared memory: 6 shared-memory round trips per sort instead of one
//      barrier + bank-conflicting pass per stage (the r01 kernel sorted all 4096 padded keys that way: 78 passes).
// The winners' depths are then evaluated with the same per-edge arithmetic as the solve kernels.
#include "dcd_common.cuh"

namespace dcd {

int launch_edge_select_bitonic(const float*, const float*, const float*, const float*, const uint8_t*, int64_t, int, int,
                               float, float, int, int64_t*, float*, float*, float*, cudaStream_t);

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / 32;
constexpr int SEL_EPT = 8;                               // sort elements per thread
constexpr int SEL_P = SEL_THREADS * SEL_EPT;             // 2048 sorted slots (k <= SEL_P)
constexpr int SEL_BUF = SEL_P + SEL_P / 16;              // one pad slot per 16: both register layouts are conflict-free

__device__ __forceinline__ int sel_phys(int idx) { return idx + (idx >> 4); }

// histogram increment with warp aggregation: the keys of an object share a handful of exponent bytes, so plain
// shared-memory atomics would serialise up to 32-fold on one address.  Must be called by all 32 lanes.
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool valid, int lane) {
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? bin : 0xffffffffu);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

// A sort element is the pair (key, ~edge id) held as two 32-bit registers; elements are distinct (unique ids) except for the
// all-zero padding, so "not less" can stand for "greater" and every exchange is branch-free predicate logic + selects
// (written on 64-bit integers the compiler turned each exchange into divergent branches).
struct SortRegs {
    uint32_t k[SEL_EPT];     // key bits
    uint32_t v[SEL_EPT];     // 0xffffffff - edge id (larger = earlier edge)
};
__device__ __forceinline__ bool elem_less(uint32_t ka, uint32_t va, uint32_t kb, uint32_t vb) {
    return (ka < kb) | ((ka == kb) & (va < vb));
}
// compare-exchange inside a thread: afterwards slot a (the lower index) holds the larger element iff desc
__device__ __forceinline__ void ce_reg(SortRegs& x, int a, int b, bool desc) {
    const bool swap = elem_less(x.k[a], x.v[a], x.k[b], x.v[b]) == desc;
    const uint32_t ka = x.k[a], va = x.v[a], kb = x.k[b], vb = x.v[b];
    x.k[a] = swap ? kb : ka;
    x.v[a] = swap ? vb : va;
    x.k[b] = swap ? ka : kb;
    x.v[b] = swap ? va : vb;
}

// Everything below is indexed by compile-time constants (SIZE / FROM are template parameters) so that the 8 elements of a
// thread stay in registers and every compare-exchange is straight-line code.

// strides min(4, FROM), ..., 1 of the merge step SIZE on the thread's 8 consecutive elements (layout A: idx = 8 t + r)
template <int SIZE, int FROM>
__device__ __forceinline__ void merge_regs_A(SortRegs& x, int base_idx) {
#pragma unroll
    for (int stride = 4; stride >= 1; stride >>= 1) {
        if (stride <= FROM) {
#pragma unroll
            for (int r = 0; r < SEL_EPT; ++r) {
                if ((r & stride) == 0) {
                    const bool desc = ((base_idx + r) & SIZE) == 0;
                    ce_reg(x, r, r + stride, desc);
                }
            }
        }
    }
}

// strides min(128, FROM), ..., 8 of the merge step SIZE through warp shuffles (layout A: lane bit m <-> stride 8 << m)
template <int SIZE, int FROM>
__device__ __forceinline__ void merge_shfl_A(SortRegs& x, int base_idx, int lane) {
#pragma unroll
    for (int m = 4; m >= 0; --m) {
        if ((8 << m) <= FROM) {
            const bool lower = (lane & (1 << m)) == 0;
#pragma unroll
            for (int r = 0; r < SEL_EPT; ++r) {
                const uint32_t pk = __shfl_xor_sync(0xffffffffu, x.k[r], 1 << m), pv = __shfl_xor_sync(0xffffffffu, x.v[r], 1 << m);
                const bool desc = ((base_idx + r) & SIZE) == 0;
                const bool want_max = lower == desc;
                const bool take = elem_less(x.k[r], x.v[r], pk, pv) == want_max;      // the partner makes the mirrored choice
                x.k[r] = take ? pk : x.k[r];
                x.v[r] = take ? pv : x.v[r];
            }
        }
    }
}

// strides SIZE / 2, ..., 256 of the merge step SIZE on layout B (idx = 256 r + t): the stride is a register distance
template <int SIZE>
__device__ __forceinline__ void merge_regs_B(SortRegs& x, int tid) {
#pragma unroll
    for (int rs = SEL_EPT / 2; rs >= 1; rs >>= 1) {
        if (rs * SEL_THREADS <= SIZE / 2) {
#pragma unroll
            for (int r = 0; r < SEL_EPT; ++r) {
                if ((r & rs) == 0) {
                    const bool desc = ((r * SEL_THREADS + tid) & SIZE) == 0;
                    ce_reg(x, r, r + rs, desc);
                }
            }
        }
    }
}

// one merge step of size SIZE >= 512: transpose to layout B through shared memory, exchange in registers, transpose back
template <int SIZE>
__device__ __forceinline__ void merge_big(SortRegs& x, uint2* buf_s, int tid, int lane) {
    const int baseA = tid * SEL_EPT;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SEL_EPT; ++r) buf_s[sel_phys(baseA + r)] = make_uint2(x.v[r], x.k[r]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SEL_EPT; ++r) { const uint2 e = buf_s[sel_phys(r * SEL_THREADS + tid)]; x.v[r] = e.x; x.k[r] = e.y; }
    merge_regs_B<SIZE>(x, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SEL_EPT; ++r) buf_s[sel_phys(r * SEL_THREADS + tid)] = make_uint2(x.v[r], x.k[r]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SEL_EPT; ++r) { const uint2 e = buf_s[sel_phys(baseA + r)]; x.v[r] = e.x; x.k[r] = e.y; }
    merge_shfl_A<SIZE, 128>(x, baseA, lane);
    merge_regs_A<SIZE, 4>(x, baseA);
}

template <int SIZE>
__device__ __forceinline__ void merge_warp(SortRegs& x, int baseA, int lane) {
    merge_shfl_A<SIZE, SIZE / 2>(x, baseA, lane);
    merge_regs_A<SIZE, 4>(x, baseA);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
edge_select_radix_kernel(const float* __restrict__ kps, const float* __restrict__ kps3d,
                         const float* __restrict__ rot, const float* __restrict__ K,
                         const uint8_t* __restrict__ kpt_mask, int64_t N, int n, int k,
                         float lo, float hi, int flags,
                         int64_t* __restrict__ idx_out, float* __restrict__ depth_sel,
                         float* __restrict__ mask_sel, float* __restrict__ depth_mean) {
    static_assert(THREADS == SEL_THREADS, "layout constants assume 256 threads");
    const int E = n * (n - 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* buf_s = reinterpret_cast<uint2*>(smem_raw);                                     // [SEL_BUF] (.x = ~id, .y = key)
    float4* kp_s = reinterpret_cast<float4*>(buf_s + SEL_BUF);                             // [n] {v, Y, vC, C}
    float* v_s = reinterpret_cast<float*>(kp_s + n);                                       // [n]
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(v_s + ((n + 3) & ~3));                  // [256]
    uint32_t* misc_s = hist_s + 256;                                                       // [32]: warp counts, broadcast
    float* red_s = reinterpret_cast<float*>(misc_s + 32);                                  // [SEL_WARPS]
    uint32_t* key_s = reinterpret_cast<uint32_t*>(red_s + SEL_WARPS);                      // [E]
    uint8_t* m_s = reinterpret_cast<uint8_t*>(key_s + E);                                  // [n]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool normalise = (flags & DCD_NORMALISE_2D) != 0;

    for (int64_t obj = blockIdx.x; obj < N; obj += gridDim.x) {
        __syncthreads();                                     // previous object's readers are done with shared memory
        // ---- stage the keypoint terms
        float b3 = 0.f;
        {
            float cy = 0.f, fy = 1.f;
            if (K != nullptr) {
                const float* Ko = K + obj * 12;
                if (normalise) { cy = __ldg(Ko + 6); fy = __ldg(Ko + 5); }
                if (flags & DCD_SUB_B3) b3 = __ldg(Ko + 11);
            }
            if (tid < ((n + 31) & ~31)) {
                const float r = __ldg(rot + obj);
                const float sn = sinf(r), cs = cosf(r);
                if (tid < n) {
                    const float2 uv = __ldg(reinterpret_cast<const float2*>(kps + (obj * n + tid) * 2));
                    const float* p3 = kps3d + (obj * n + tid) * 3;
                    const float4 t = keypoint_terms(uv.y, __ldg(p3), __ldg(p3 + 1), __ldg(p3 + 2), sn, cs, normalise, cy, fy);
                    kp_s[tid] = t;
                    v_s[tid] = t.x;
                    if (kpt_mask != nullptr) m_s[tid] = kpt_mask[obj * n + tid];
                }
            }
            if (tid < 256) hist_s[tid] = 0u;
        }
        __syncthreads();
        // ---- phase 1: keys of all edges + first histogram (bits 31..24)
        for (int i = warp; i < n - 1; i += SEL_WARPS) {
            const float vi = v_s[i];
            const int e0 = row_offset(i, n) - i - 1;
            for (int j0 = i + 1; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                const bool valid = j < n;
                const uint32_t key = __float_as_uint(fabsf(__fsub_rn(vi, v_s[valid ? j : i])));
                if (valid) key_s[e0 + j] = key;
                hist_add(hist_s, key >> 24, valid, lane);
            }
        }
        __syncthreads();
        // ---- phase 2a: radix select of the k-th largest key
        uint32_t prefix = 0u, mask = 0u;
        int remaining = k;
#pragma unroll 1
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (shift != 24) {
                for (int e0 = 0; e0 < E; e0 += THREADS) {
                    const int e = e0 + tid;
                    const uint32_t key = e < E ? key_s[e] : 0u;
                    hist_add(hist_s, (key >> shift) & 255u, e < E && (key & mask) == prefix, lane);
                }
                __syncthreads();
            }
            if (warp == 0) {
                // lane l owns bins 255 - 8 l ... 248 - 8 l (descending); find the bin where the count from the top reaches `remaining`
                uint32_t c[8], sum = 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[q] = hist_s[255 - 8 * lane - q]; sum += c[q]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                uint32_t above = incl - sum;                 // keys in bins above this lane's
                const bool mine = above < (uint32_t)remaining && (uint32_t)remaining <= incl;
                if (mine) {
                    int bin = 255 - 8 * lane;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (above + c[q] >= (uint32_t)remaining) { bin = 255 - 8 * lane - q; break; }
                        above += c[q];
                    }
                    misc_s[0] = (uint32_t)bin;
                    misc_s[1] = above;
                }
            }
            __syncthreads();
            prefix |= misc_s[0] << shift;
            mask |= 255u << shift;
            remaining -= (int)misc_s[1];
            __syncthreads();
            if (tid < 256) hist_s[tid] = 0u;
            __syncthreads();
        }
        const uint32_t T = prefix;                           // the k-th largest key; `remaining` ties at T are selected
        // ---- phase 2b: ordered compaction of {key > T} and the first `remaining` {key == T} (edge id order)
        const int CW = (E + SEL_WARPS - 1) / SEL_WARPS;      // contiguous id range per warp
        const int w_lo = warp * CW, w_hi = min(E, w_lo + CW);
        {
            uint32_t ngt = 0u, neq = 0u;
            for (int e0 = w_lo; e0 < w_hi; e0 += 32) {
                const int e = e0 + lane;
                const uint32_t key = e < w_hi ? key_s[e] : 0u;
                const bool gt = e < w_hi && key > T, eq = e < w_hi && key == T;
                ngt += __popc(__ballot_sync(0xffffffffu, gt));
                neq += __popc(__ballot_sync(0xffffffffu, eq));
            }
            if (lane == 0) { misc_s[2 + warp] = ngt; misc_s[2 + SEL_WARPS + warp] = neq; }
        }
        for (int q = tid; q < SEL_P - k; q += THREADS) buf_s[sel_phys(k + q)] = make_uint2(0u, 0u);     // padding sorts last
        __syncthreads();
        {
            uint32_t gt_off = 0u, eq_off = 0u;
            for (int w = 0; w < warp; ++w) { gt_off += misc_s[2 + w]; eq_off += misc_s[2 + SEL_WARPS + w]; }
            const uint32_t c_gt = (uint32_t)(k - remaining);
            const uint32_t lt_mask = (1u << lane) - 1u;
            for (int e0 = w_lo; e0 < w_hi; e0 += 32) {
                const int e = e0 + lane;
                const uint32_t key = e < w_hi ? key_s[e] : 0u;
                const bool gt = e < w_hi && key > T, eq = e < w_hi && key == T;
                const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
                const uint2 comp = make_uint2(0xffffffffu - (uint32_t)e, key);
                if (gt) buf_s[sel_phys((int)(gt_off + __popc(bg & lt_mask)))] = comp;
                if (eq) {
                    const uint32_t rnk = eq_off + __popc(be & lt_mask);
                    if (rnk < (uint32_t)remaining) buf_s[sel_phys((int)(c_gt + rnk))] = comp;
                }
                gt_off += __popc(bg);
                eq_off += __popc(be);
            }
        }
        __syncthreads();
        // ---- phase 3: bitonic sort (descending) of SEL_P composites, 8 per thread in registers
        SortRegs x;
        const int baseA = tid * SEL_EPT;                     // layout A: idx = 8 t + r
#pragma unroll
        for (int r = 0; r < SEL_EPT; ++r) {                  // any bijection will do for unsorted input
            const uint2 e = buf_s[sel_phys(r * THREADS + tid)];
            x.v[r] = e.x; x.k[r] = e.y;
        }
        merge_regs_A<2, 1>(x, baseA);
        merge_regs_A<4, 2>(x, baseA);
        merge_regs_A<8, 4>(x, baseA);
        merge_warp<16>(x, baseA, lane);
        merge_warp<32>(x, baseA, lane);
        merge_warp<64>(x, baseA, lane);
        merge_warp<128>(x, baseA, lane);
        merge_warp<256>(x, baseA, lane);
        merge_big<512>(x, buf_s, tid, lane);
        merge_big<1024>(x, buf_s, tid, lane);
        merge_big<2048>(x, buf_s, tid, lane);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SEL_EPT; ++r) buf_s[sel_phys(baseA + r)] = make_uint2(x.v[r], x.k[r]);
        __syncthreads();
        // ---- outputs for the k winners
        float acc = 0.f;
        for (int r = tid; r < k; r += THREADS) {
            const int e = (int)(0xffffffffu - buf_s[sel_phys(r)].x);
            idx_out[obj * k + r] = (int64_t)e;
            if (depth_sel != nullptr || mask_sel != nullptr || depth_mean != nullptr) {
                int i, j;
                decode_edge(e, n, i, j);
                const float z = edge_depth(kp_s[i], kp_s[j], lo, hi, b3);
                if (depth_sel != nullptr) depth_sel[obj * k + r] = z;
                if (mask_sel != nullptr) mask_sel[obj * k + r] = (m_s[i] != 0 && m_s[j] != 0) ? 1.f : 0.f;
                acc += z;
            }
        }
        if (depth_mean != nullptr) {
            acc = warp_sum(acc);
            if (lane == 0) red_s[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < SEL_WARPS; ++w) t += red_s[w];
                depth_mean[obj] = __fdiv_rn(t, (float)k);
            }
        }
    }
}

}  // namespace

int launch_edge_select(const float* kps, const float* kps3d, const float* rot, const float* K,
                       const uint8_t* kpt_mask, int64_t N, int n, int k, float lo, float hi, int flags,
                       int64_t* idx_out, float* depth_sel, float* mask_sel, float* depth_mean, cudaStream_t st) {
    if (k > SEL_P)                                           // more winners than the register sort holds: full shared-memory sort
        return launch_edge_select_bitonic(kps, kps3d, rot, K, kpt_mask, N, n, k, lo, hi, flags, idx_out, depth_sel, mask_sel,
                                          depth_mean, st);
    const int E = n * (n - 1) / 2;
    const size_t smem = (size_t)SEL_BUF * 8 + (size_t)n * 16 + (size_t)((n + 3) & ~3) * 4 + 256 * 4 + 32 * 4 + SEL_WARPS * 4 +
                        (size_t)E * 4 + (size_t)((n + 15) & ~15);
    if (smem > 227 * 1024) return DCD_E_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(edge_select_radix_kernel<SEL_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    const int64_t max_grid = (int64_t)device_sm_count() * per_sm;
    const int grid = (int)(N < max_grid ? N : max_grid);
    edge_select_radix_kernel<SEL_THREADS><<<grid, SEL_THREADS, smem, st>>>(kps, kps3d, rot, K, kpt_mask, N, n, k, lo, hi, flags,
                                                                          idx_out, depth_sel, mask_sel, depth_mean);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

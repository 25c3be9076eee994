// Frame epilogue of the DGDE detector head around the edge solve, one kernel:
//   image-space keypoints   real_2d = (kpts_off + (points + offsets)) * down_ratio - pad       detector_infer.py:216-217
//   edge solve + mean       depth   = mean_e clamp(|H| / max(|V|, 1e-10), lo, hi) - b3          anno_encoder.py:326-390, :225
//   3D location             (u, v)  = (points + offsets) * down_ratio - pad                      anno_encoder.py:147-161
//                           x = ((u - c_u) * depth) / f_u + b_x,  y likewise,  z = depth         kitti_utils.py:239-244, 399-417
//                           y += h / 2                                                           detector_infer.py:188
// so the 2D keypoints never exist in global memory and the location needs no second launch.
// One warp per object, same staging as the throughput edge-solve kernel (edge_solve.cu): 16-byte per-keypoint
// terms in shared memory, (i, j) byte offsets from a CTA-shared pair table, per-lane strided accumulation.
// Every operation is rounded like the reference's FP32 torch sequence; the calibration scalars b_x, b_y are formed
// in FP32 from the FP32 matrix (the reference forms them in float64 numpy: a 1-ulp difference of a 0.06 m offset).
#include "dcd_common.cuh"

namespace dcd {
namespace {

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dgde_locate_warp_kernel(const float* __restrict__ kpts_off, const float* __restrict__ kps3d, const float* __restrict__ rot,
                        const float* __restrict__ K, const float* __restrict__ points, const float* __restrict__ offsets,
                        const float* __restrict__ pad, const float* __restrict__ dims, const float* __restrict__ depth_in,
                        int64_t N, int n, float lo, float hi, int flags, float down_ratio,
                        float* __restrict__ depth_out, float* __restrict__ loc_out) {
    constexpr int NWARP = THREADS / 32;
    const int E = n * (n - 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* kp_all = reinterpret_cast<float4*>(smem_raw);                     // [NWARP][n]
    uint32_t* tab_s = reinterpret_cast<uint32_t*>(kp_all + NWARP * n);        // [E] byte offsets (i*16) | (j*16) << 16
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool solve = kpts_off != nullptr;
    if (solve) {
        for (int e = tid; e < E; e += THREADS) {
            int i, j;
            decode_edge(e, n, i, j);
            tab_s[e] = (uint32_t)(i * 16) | ((uint32_t)(j * 16) << 16);
        }
    }
    __syncthreads();
    float4* kp = kp_all + warp * n;
    const unsigned char* kpb = reinterpret_cast<const unsigned char*>(kp);
    const bool normalise = (flags & DCD_NORMALISE_2D) != 0;
    for (int64_t obj = (int64_t)blockIdx.x * NWARP + warp; obj < N; obj += (int64_t)gridDim.x * NWARP) {
        const float* Ko = K + obj * 12;
        const float f_u = __ldg(Ko + 0), c_u = __ldg(Ko + 2), f_v = __ldg(Ko + 5), c_v = __ldg(Ko + 6);
        const float cx = __fadd_rn(__ldg(points + obj * 2), __ldg(offsets + obj * 2));
        const float cyy = __fadd_rn(__ldg(points + obj * 2 + 1), __ldg(offsets + obj * 2 + 1));
        const float pad_u = __ldg(pad + obj * 2), pad_v = __ldg(pad + obj * 2 + 1);
        float depth;
        if (solve) {
            const float b3 = (flags & DCD_SUB_B3) ? __ldg(Ko + 11) : 0.f;
            const float r = __ldg(rot + obj);
            const float sn = sinf(r), cs = cosf(r);
            for (int t = lane; t < n; t += 32) {
                const float off_v = __ldg(kpts_off + (obj * n + t) * 2 + 1);
                const float v_img = __fsub_rn(__fmul_rn(__fadd_rn(off_v, cyy), down_ratio), pad_v);
                const float* p3 = kps3d + (obj * n + t) * 3;
                kp[t] = keypoint_terms(v_img, __ldg(p3), __ldg(p3 + 1), __ldg(p3 + 2), sn, cs, normalise, c_v, f_v);
            }
            __syncwarp();
            float acc = 0.f;
            for (int e = lane; e < E; e += 32) {
                const uint32_t p = tab_s[e];
                const float4 a = *reinterpret_cast<const float4*>(kpb + (p & 0xffffu));
                const float4 b = *reinterpret_cast<const float4*>(kpb + (p >> 16));
                acc += edge_depth(a, b, lo, hi, b3);
            }
            acc = warp_sum(acc);
            depth = __fdiv_rn(acc, (float)E);
            __syncwarp();                                    // all lanes done with kp before it is restaged
        } else {
            depth = __ldg(depth_in + obj);
        }
        if (lane == 0) {
            if (depth_out != nullptr) depth_out[obj] = depth;
            if (loc_out != nullptr) {
                const float u = __fsub_rn(__fmul_rn(cx, down_ratio), pad_u);
                const float v = __fsub_rn(__fmul_rn(cyy, down_ratio), pad_v);
                const float b_x = __fdiv_rn(__ldg(Ko + 3), -f_u), b_y = __fdiv_rn(__ldg(Ko + 7), -f_v);
                float x = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(u, c_u), depth), f_u), b_x);
                float y = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(v, c_v), depth), f_v), b_y);
                if (dims != nullptr) y = __fadd_rn(y, __fmul_rn(__ldg(dims + obj * 3 + 1), 0.5f));
                loc_out[obj * 3 + 0] = x;
                loc_out[obj * 3 + 1] = y;
                loc_out[obj * 3 + 2] = depth;
            }
        }
    }
}

// The same epilogue reading the detector's regression map directly (row N2 fused into the load stage): for detection d at
// heat-map position index[d] of image batch_idx[d] the keypoint offsets, 3D template points and the sub-pixel offset are
// the values of their channel groups at that position of the [B,C,H,W] map (select_point_of_interest,
// DGDE/model/layers/utils.py:120-145, + the key2channel slices of detector_infer.py:133,216,219), so POI gather ->
// image keypoints -> edge solve -> mean -> location is ONE launch and the [B,K,C] gather never exists in memory.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dgde_frame_warp_kernel(const float* __restrict__ fmap, const int64_t* __restrict__ index, const int32_t* __restrict__ batch_idx,
                       int C, int H, int W, int ch_k2, int ch_k3, int ch_off,
                       const float* __restrict__ rot, const float* __restrict__ K, const float* __restrict__ pad,
                       const float* __restrict__ dims, int64_t N, int n, float lo, float hi, int flags, float down_ratio,
                       float* __restrict__ depth_out, float* __restrict__ loc_out, float* __restrict__ kimg_out,
                       float* __restrict__ k3_out) {
    constexpr int NWARP = THREADS / 32;
    const int E = n * (n - 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* kp_all = reinterpret_cast<float4*>(smem_raw);                     // [NWARP][n]
    uint32_t* tab_s = reinterpret_cast<uint32_t*>(kp_all + NWARP * n);        // [E]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < E; e += THREADS) {
        int i, j;
        decode_edge(e, n, i, j);
        tab_s[e] = (uint32_t)(i * 16) | ((uint32_t)(j * 16) << 16);
    }
    __syncthreads();
    float4* kp = kp_all + warp * n;
    const unsigned char* kpb = reinterpret_cast<const unsigned char*>(kp);
    const bool normalise = (flags & DCD_NORMALISE_2D) != 0;
    const int64_t HW = (int64_t)H * W;
    const float qnan = __int_as_float(0x7fc00000);
    for (int64_t obj = (int64_t)blockIdx.x * NWARP + warp; obj < N; obj += (int64_t)gridDim.x * NWARP) {
        const int64_t pos = index[obj];
        const bool inside = pos >= 0 && pos < HW;             // like dcd_poi_gather_fwd: NaN instead of a fault
        const float* base = fmap + (int64_t)(batch_idx != nullptr ? batch_idx[obj] : 0) * C * HW + (inside ? pos : 0);
        auto chan = [&](int c) { return inside ? __ldg(base + (int64_t)c * HW) : qnan; };
        const float* Ko = K + obj * 12;
        const float f_u = __ldg(Ko + 0), c_u = __ldg(Ko + 2), f_v = __ldg(Ko + 5), c_v = __ldg(Ko + 6);
        const float px = (float)(pos % W), py = (float)(pos / W);             // select_topk: xs = ind % W, ys = ind // W
        const float cx = __fadd_rn(px, chan(ch_off)), cyy = __fadd_rn(py, chan(ch_off + 1));
        const float pad_u = __ldg(pad + obj * 2), pad_v = __ldg(pad + obj * 2 + 1);
        const float b3 = (flags & DCD_SUB_B3) ? __ldg(Ko + 11) : 0.f;
        const float r = __ldg(rot + obj);
        const float sn = sinf(r), cs = cosf(r);
        for (int t = lane; t < n; t += 32) {
            const float off_v = chan(ch_k2 + 2 * t + 1);
            const float v_img = __fsub_rn(__fmul_rn(__fadd_rn(off_v, cyy), down_ratio), pad_v);
            const float X = chan(ch_k3 + 3 * t), Y = chan(ch_k3 + 3 * t + 1), Z = chan(ch_k3 + 3 * t + 2);
            kp[t] = keypoint_terms(v_img, X, Y, Z, sn, cs, normalise, c_v, f_v);
            if (kimg_out != nullptr) {
                const float u_img = __fsub_rn(__fmul_rn(__fadd_rn(chan(ch_k2 + 2 * t), cx), down_ratio), pad_u);
                kimg_out[(obj * n + t) * 2] = u_img;
                kimg_out[(obj * n + t) * 2 + 1] = v_img;
            }
            if (k3_out != nullptr) {
                float* o = k3_out + (obj * n + t) * 3;
                o[0] = X; o[1] = Y; o[2] = Z;
            }
        }
        __syncwarp();
        float acc = 0.f;
        for (int e = lane; e < E; e += 32) {
            const uint32_t p = tab_s[e];
            const float4 a = *reinterpret_cast<const float4*>(kpb + (p & 0xffffu));
            const float4 b = *reinterpret_cast<const float4*>(kpb + (p >> 16));
            acc += edge_depth(a, b, lo, hi, b3);
        }
        acc = warp_sum(acc);
        const float depth = __fdiv_rn(acc, (float)E);
        __syncwarp();
        if (lane == 0) {
            if (depth_out != nullptr) depth_out[obj] = depth;
            if (loc_out != nullptr) {
                const float u = __fsub_rn(__fmul_rn(cx, down_ratio), pad_u);
                const float v = __fsub_rn(__fmul_rn(cyy, down_ratio), pad_v);
                const float b_x = __fdiv_rn(__ldg(Ko + 3), -f_u), b_y = __fdiv_rn(__ldg(Ko + 7), -f_v);
                float x = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(u, c_u), depth), f_u), b_x);
                float y = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(v, c_v), depth), f_v), b_y);
                if (dims != nullptr) y = __fadd_rn(y, __fmul_rn(__ldg(dims + obj * 3 + 1), 0.5f));
                loc_out[obj * 3 + 0] = x;
                loc_out[obj * 3 + 1] = y;
                loc_out[obj * 3 + 2] = depth;
            }
        }
    }
}

}  // namespace

int launch_dgde_frame(const float* fmap, const int64_t* index, const int32_t* batch_idx, int C, int H, int W, int ch_k2, int ch_k3,
                      int ch_off, const float* rot, const float* K, const float* pad, const float* dims, int64_t N, int n, float lo,
                      float hi, int flags, float down_ratio, float* depth_out, float* loc_out, float* kimg_out, float* k3_out,
                      cudaStream_t st) {
    constexpr int T = 128;                                  // a frame has <= 50 detections: more, smaller CTAs
    const int E = n * (n - 1) / 2;
    const size_t smem = (size_t)(T / 32) * n * sizeof(float4) + (size_t)E * sizeof(uint32_t);
    if (smem > 227 * 1024) return DCD_E_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dgde_frame_warp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t want = (N + T / 32 - 1) / (T / 32);
    const int64_t cap = (int64_t)device_sm_count() * 8;
    const int grid = (int)(want < cap ? want : cap);
    dgde_frame_warp_kernel<T><<<grid, T, smem, st>>>(fmap, index, batch_idx, C, H, W, ch_k2, ch_k3, ch_off, rot, K, pad, dims, N, n,
                                                     lo, hi, flags, down_ratio, depth_out, loc_out, kimg_out, k3_out);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

int launch_dgde_locate(const float* kpts_off, const float* kps3d, const float* rot, const float* K, const float* points,
                       const float* offsets, const float* pad, const float* dims, const float* depth_in, int64_t N, int n,
                       float lo, float hi, int flags, float down_ratio, float* depth_out, float* loc_out, cudaStream_t st) {
    constexpr int T = 256;
    const int E = n * (n - 1) / 2;
    const size_t smem = (size_t)(T / 32) * n * sizeof(float4) + (size_t)E * sizeof(uint32_t);
    if (smem > 227 * 1024) return DCD_E_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dgde_locate_warp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t want = (N + T / 32 - 1) / (T / 32);
    const int64_t cap = (int64_t)device_sm_count() * 8;
    const int grid = (int)(want < cap ? want : cap);
    dgde_locate_warp_kernel<T><<<grid, T, smem, st>>>(kpts_off, kps3d, rot, K, points, offsets, pad, dims, depth_in, N, n, lo, hi,
                                                      flags, down_ratio, depth_out, loc_out);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

// ---------------------------------------------------------------------------------------------
// Rest of row N4: per-object depth ensemble of the detector head and the GMW-validation ray rescale.
// One thread per object; every operation rounded like the reference's FP32 torch sequence.
// ---------------------------------------------------------------------------------------------
namespace dcd {
namespace {

// depth of one projected height: f_u * h3d / (relu(height) * down_ratio + eps)          anno_encoder.py:209-211
__device__ __forceinline__ float height_depth(float fh, float height, float down_ratio, float eps) {
    return __fdiv_rn(fh, __fadd_rn(__fmul_rn(fmaxf(height, 0.f), down_ratio), eps));
}

__global__ void __launch_bounds__(256)
dgde_depth_ensemble_kernel(const float* __restrict__ kp10, const float* __restrict__ dims, const float* __restrict__ K,
                           const float* __restrict__ direct, const float* __restrict__ log_unc_direct,
                           const float* __restrict__ log_unc_kp, const float* __restrict__ scores, int64_t N, float down_ratio,
                           float eps, float lo, float hi, float* __restrict__ kp_depths, float* __restrict__ depth,
                           float* __restrict__ depth_error, int64_t* __restrict__ argmax, float* __restrict__ scores_out) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const float* kp = kp10 + o * 20;
    float v[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) v[t] = __ldg(kp + 2 * t + 1);
    const float fh = __fmul_rn(__ldg(K + o * 12), __ldg(dims + o * 3 + 1));          // calib.f_u * pred_height_3D
    // anno_encoder.py:199-201, 209-214, 221: centre line, corner pairs (0,4),(2,6) and (1,5),(3,7); mean of two; clamp
    float d[4];
    d[1] = height_depth(fh, __fsub_rn(v[8], v[9]), down_ratio, eps);
    d[2] = __fmul_rn(__fadd_rn(height_depth(fh, __fsub_rn(v[0], v[4]), down_ratio, eps),
                               height_depth(fh, __fsub_rn(v[2], v[6]), down_ratio, eps)), 0.5f);
    d[3] = __fmul_rn(__fadd_rn(height_depth(fh, __fsub_rn(v[1], v[5]), down_ratio, eps),
                               height_depth(fh, __fsub_rn(v[3], v[7]), down_ratio, eps)), 0.5f);
#pragma unroll
    for (int j = 1; j < 4; ++j) d[j] = fminf(fmaxf(d[j], lo), hi);
    if (kp_depths != nullptr) {
        kp_depths[o * 3 + 0] = d[1]; kp_depths[o * 3 + 1] = d[2]; kp_depths[o * 3 + 2] = d[3];
    }
    if (depth == nullptr && depth_error == nullptr && argmax == nullptr && scores_out == nullptr) return;
    // detector_infer.py:141,154,158-171: uncertainties = exp(channel); weights = (1/u) / sum(1/u)
    const int first = direct != nullptr ? 0 : 1;
    float u[4], w[4];
    if (first == 0) {
        d[0] = __ldg(direct + o);
        u[0] = expf(__ldg(log_unc_direct + o));
    }
#pragma unroll
    for (int j = 1; j < 4; ++j) u[j] = expf(__ldg(log_unc_kp + o * 3 + j - 1));
    float wsum = 0.f;
    int best = first;
    for (int j = first; j < 4; ++j) {
        w[j] = __fdiv_rn(1.f, u[j]);
        wsum = __fadd_rn(wsum, w[j]);
        if (w[j] > w[best]) best = j;                       // first maximum, like torch.argmax
    }
    float dsum = 0.f, esum = 0.f;
    for (int j = first; j < 4; ++j) {
        const float wn = __fdiv_rn(w[j], wsum);
        dsum = __fadd_rn(dsum, __fmul_rn(d[j], wn));
        esum = __fadd_rn(esum, __fmul_rn(wn, u[j]));
    }
    if (depth != nullptr) depth[o] = dsum;
    if (depth_error != nullptr) depth_error[o] = esum;
    if (argmax != nullptr) argmax[o] = best - first;
    if (scores_out != nullptr) {                             // detector_infer.py:197-203
        const float conf = __fsub_rn(1.f, fminf(fmaxf(esum, 0.01f), 1.f));
        const float s = __fmul_rn(__ldg(scores + o), conf);
        scores_out[o] = (esum != esum || s != s) ? 0.f : s;       // torch.clamp keeps NaN (fmaxf drops it): NaN -> 0
    }
}

// GMW/main.py:542-547
__global__ void __launch_bounds__(256)
gmw_ray_rescale_kernel(const float* __restrict__ raw_location, const float* __restrict__ pred_depth, const float* __restrict__ dim,
                       int64_t N, float* __restrict__ out) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const float x = __ldg(raw_location + o * 3), y = __ldg(raw_location + o * 3 + 1), z = __ldg(raw_location + o * 3 + 2);
    const float scale = __fdiv_rn(__ldg(pred_depth + o), z);
    const float hh = __fmul_rn(__ldg(dim + o * 3), 0.5f);
    out[o * 3 + 0] = __fmul_rn(scale, x);
    out[o * 3 + 1] = __fadd_rn(__fmul_rn(scale, __fsub_rn(y, hh)), hh);
    out[o * 3 + 2] = __fmul_rn(scale, z);
}

}  // namespace

int launch_dgde_depth_ensemble(const float* kp10, const float* dims, const float* K, const float* direct, const float* lud,
                               const float* luk, const float* scores, int64_t N, float down_ratio, float eps, float lo, float hi,
                               float* kp_depths, float* depth, float* depth_error, int64_t* argmax, float* scores_out,
                               cudaStream_t st) {
    dgde_depth_ensemble_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(kp10, dims, K, direct, lud, luk, scores, N, down_ratio, eps,
                                                                             lo, hi, kp_depths, depth, depth_error, argmax, scores_out);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

int launch_gmw_ray_rescale(const float* raw_location, const float* pred_depth, const float* dim, int64_t N, float* out,
                           cudaStream_t st) {
    gmw_ray_rescale_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(raw_location, pred_depth, dim, N, out);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

// ---------------------------------------------------------------------------------------------
// Row N2, upstream gather: select_point_of_interest (DGDE/model/layers/utils.py:120-145) without the NCHW -> NHWC copy of
// the whole regression map the reference makes to pick <= 50 points per image:  out[b, k, c] = feat[b, c, index[b, k]].
// ---------------------------------------------------------------------------------------------
namespace dcd {
namespace {

__global__ void __launch_bounds__(256)
poi_gather_kernel(const float* __restrict__ feat, const int64_t* __restrict__ index, int64_t B, int64_t Kp, int C, int64_t HW,
                  float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * Kp * C) return;
    const int c = (int)(t % C);
    const int64_t bk = t / C, b = bk / Kp;
    const int64_t p = index[bk];
    out[t] = (p >= 0 && p < HW) ? __ldg(feat + (b * C + c) * HW + p) : __int_as_float(0x7fc00000);   // out of range: NaN, no fault
}

}  // namespace

int launch_poi_gather(const float* feat, const int64_t* index, int64_t B, int64_t Kp, int C, int64_t HW, float* out, cudaStream_t st) {
    const int64_t total = B * Kp * C;
    poi_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, index, B, Kp, C, HW, out);
    DCD_CHECK_LAUNCH();
    return DCD_OK;
}

}  // namespace dcd

"""Host-side mirror of the reference's call surface for the densely-constrained-depth path.

Same names, argument meaning and return values as the reference functions they replace
(SURVEY.md section 8b); the arithmetic runs in libdcd_b200.so (CUDA, sm_100a).  There is no
CPU path: non-CUDA tensors raise.

    decode_pairs_kpts_depth   <- Anno_Encoder.decode_pairs_kpts_depth   DGDE/model/anno_encoder.py:326
    compute_z                 <- compute_z                              GMW/main.py:373
    GMW.forward               <- GMW.forward                            GMW/model/model.py:195
    compute_reg_loss          <- compute_reg_loss                       GMW/main.py:364
plus fused fast paths for callers that do not need the per-edge intermediates:
    edge_depth_mean           decode_pairs_kpts_depth(...)[0].mean(1)   detector_infer.py:222-225, detector_loss.py:388
    gmw_weighted_depth        compute_z -> GMW.forward -> weighted sum  GMW/main.py:524-533
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, require_cuda, stream_ptr
from .weights import NET_DEPTH, NET_NAMES, blob_size, pack_state_dict, unpack_blob

_CHECK_FINITE = os.environ.get("DCD_B200_CHECK_FINITE", "0") == "1"   # debug: synchronising range check of GMW.forward
K_SEL = 1500                 # anno_encoder.py:378, GMW/main.py:413
DGDE_CLAMP = (2.0, 80.0)     # anno_encoder.py:375
GMW_CLAMP = (0.1, 80.0)      # GMW/main.py:410
FLAG_NORMALISE_2D = 1
FLAG_SUB_B3 = 2
FLAG_FAST_QUOTIENT = 4
MAX_KPTS = 256


def _num_edges(n: int) -> int:
    return n * (n - 1) // 2


def _inference_only(name: str, *tensors) -> None:
    """The frame-epilogue ops have no backward: fail loudly instead of silently cutting the
    autograd graph when the reference would differentiate through them (detector_loss.py:391)."""
    if torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        raise RuntimeError("dcd_b200.%s is inference-only (no backward is implemented): call it under torch.no_grad() "
                           "or detach its inputs" % name)


def _check_shapes(kps, kps_3d, rot, K=None):
    if kps.dim() != 3 or kps.shape[-1] != 2:
        raise ValueError("kps must be [N,n,2], got %s" % (tuple(kps.shape),))
    N, n = kps.shape[0], kps.shape[1]
    if tuple(kps_3d.shape) != (N, n, 3):
        raise ValueError("kps_3d must be [N,n,3] matching kps, got %s" % (tuple(kps_3d.shape),))
    if rot.numel() != N:
        raise ValueError("rot_y must hold one yaw per object ([N,1] or [N])")
    if K is not None and tuple(K.shape) != (N, 3, 4):
        raise ValueError("K must be [N,3,4], got %s" % (tuple(K.shape),))
    if not 2 <= n <= MAX_KPTS:
        raise ValueError("keypoints per object must be in [2, %d]" % MAX_KPTS)
    return N, n


# ---------------------------------------------------------------------------------------------
# edge solve
# ---------------------------------------------------------------------------------------------
class _EdgeSolve(torch.autograd.Function):
    """All E edge depths (+ optional fused mean).  Gradients flow to kps and kps_3d only."""

    @staticmethod
    def forward(ctx, kps, kps_3d, rot, K, lo, hi, flags, want_edges, want_mean):
        N, n = kps.shape[0], kps.shape[1]
        E = _num_edges(n)
        dev = kps.device
        edges = torch.empty((N, E), dtype=torch.float32, device=dev) if want_edges else None
        mean = torch.empty((N,), dtype=torch.float32, device=dev) if want_mean else None
        if N:
            check(_lib.lib().dcd_edge_solve_fwd(ptr(kps), ptr(kps_3d), ptr(rot), ptr(K), N, n, lo, hi, flags,
                                                ptr(edges), ptr(mean), stream_ptr()), "dcd_edge_solve_fwd")
        ctx.save_for_backward(kps, kps_3d, rot, K)
        ctx.cfg = (lo, hi, flags)
        ctx.set_materialize_grads(False)
        if edges is None:
            edges = torch.empty((0,), dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(edges)
        if mean is None:
            mean = torch.empty((0,), dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(mean)
        return edges, mean

    @staticmethod
    def backward(ctx, g_edges, g_mean):
        kps, kps_3d, rot, K = ctx.saved_tensors
        lo, hi, flags = ctx.cfg
        N, n = kps.shape[0], kps.shape[1]
        g_kps = torch.zeros_like(kps)
        g_k3 = torch.zeros_like(kps_3d)
        if N and (g_edges is not None or g_mean is not None):
            ge = f32c(g_edges) if g_edges is not None else None
            gm = f32c(g_mean) if g_mean is not None else None
            check(_lib.lib().dcd_edge_solve_bwd(ptr(kps), ptr(kps_3d), ptr(rot), ptr(K), N, n, lo, hi, flags,
                                                0, 0, ptr(ge), ptr(gm), ptr(g_kps), ptr(g_k3), stream_ptr()),
                  "dcd_edge_solve_bwd")
        return g_kps, g_k3, None, None, None, None, None, None, None


class _EdgeSelectSolve(torch.autograd.Function):
    """Top-k edges by |V| (canonical order) with their depths, pair mask and mean."""

    @staticmethod
    def forward(ctx, kps, kps_3d, rot, K, mask_u8, lo, hi, flags, k, want_mean):
        N, n = kps.shape[0], kps.shape[1]
        dev = kps.device
        idx = torch.empty((N, k), dtype=torch.int64, device=dev)
        depth = torch.empty((N, k), dtype=torch.float32, device=dev)
        msel = torch.empty((N, k), dtype=torch.float32, device=dev) if mask_u8 is not None else None
        mean = torch.empty((N,), dtype=torch.float32, device=dev) if want_mean else None
        if N:
            check(_lib.lib().dcd_edge_select_fwd(ptr(kps), ptr(kps_3d), ptr(rot), ptr(K), ptr(mask_u8), N, n, k,
                                                 lo, hi, flags, ptr(idx), ptr(depth), ptr(msel), ptr(mean),
                                                 stream_ptr()), "dcd_edge_select_fwd")
        ctx.save_for_backward(kps, kps_3d, rot, K, idx)
        ctx.cfg = (lo, hi, flags, k)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(idx)
        if msel is None:
            msel = torch.empty((0,), dtype=torch.float32, device=dev)
        ctx.mark_non_differentiable(msel)
        if mean is None:
            mean = torch.empty((0,), dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(mean)
        return depth, msel, idx, mean

    @staticmethod
    def backward(ctx, g_depth, _g_mask, _g_idx, g_mean):
        kps, kps_3d, rot, K, idx = ctx.saved_tensors
        lo, hi, flags, k = ctx.cfg
        N, n = kps.shape[0], kps.shape[1]
        g_kps = torch.zeros_like(kps)
        g_k3 = torch.zeros_like(kps_3d)
        if N and (g_depth is not None or g_mean is not None):
            gd = f32c(g_depth) if g_depth is not None else None
            gm = f32c(g_mean) if g_mean is not None else None
            check(_lib.lib().dcd_edge_solve_bwd(ptr(kps), ptr(kps_3d), ptr(rot), ptr(K), N, n, lo, hi, flags,
                                                ptr(idx), k, ptr(gd), ptr(gm), ptr(g_kps), ptr(g_k3), stream_ptr()),
                  "dcd_edge_solve_bwd")
        return g_kps, g_k3, None, None, None, None, None, None, None, None


def _prep_dgde(kps, kps_3d, rot_y, K):
    require_cuda(kps, kps_3d, rot_y, K)
    kps, kps_3d = f32c(kps), f32c(kps_3d)
    rot = f32c(rot_y).reshape(-1)
    K = f32c(K)                       # float64 stride-0 calibration at inference (detector_infer.py:221)
    _check_shapes(kps, kps_3d, rot, K)
    return kps, kps_3d, rot, K


def decode_pairs_kpts_depth(kps, kps_3d, rot_y, K, training=False, kpts_2d_mask=None, gt_depth=None, weight=None,
                            num_k: int = K_SEL, return_idx: bool = False):
    """Drop-in for Anno_Encoder.decode_pairs_kpts_depth (DGDE/model/anno_encoder.py:326-390).

    Returns (depth_all, depth_mask): depth_all is [N,E] (all edges, row-major i<j order) when not
    training, else the [N,1500] depths of the edges with the largest |v_i - v_j| sorted descending
    (ties by ascending edge id); depth_mask is the float32 0/1 product of the keypoint masks of the
    selected pairs, or None when kpts_2d_mask is None.  `gt_depth` and `weight` are accepted and
    unused, as in the reference.  Differentiable w.r.t. kps (v column) and kps_3d.
    """
    kps, kps_3d, rot, K = _prep_dgde(kps, kps_3d, rot_y, K)
    n = kps.shape[1]
    flags = FLAG_NORMALISE_2D | FLAG_SUB_B3
    lo, hi = DGDE_CLAMP
    if not training:
        depth, _ = _EdgeSolve.apply(kps, kps_3d, rot, K, lo, hi, flags, True, False)
        mask = None
        if kpts_2d_mask is not None:   # the reference still returns the (un-gathered) pair mask in this case
            m = kpts_2d_mask.to(torch.float32)
            ii, jj = torch.triu_indices(n, n, 1, device=kps.device)
            mask = m[:, ii] * m[:, jj]
        return (depth, mask, None) if return_idx else (depth, mask)
    if _num_edges(n) < num_k:
        raise RuntimeError("selected index k out of range: %d edges < k=%d (n=%d keypoints)" % (_num_edges(n), num_k, n))
    mask_u8 = None
    if kpts_2d_mask is not None:
        require_cuda(kpts_2d_mask)
        mask_u8 = (kpts_2d_mask != 0).to(torch.uint8).contiguous()
    depth, msel, idx, _ = _EdgeSelectSolve.apply(kps, kps_3d, rot, K, mask_u8, lo, hi, flags, num_k, False)
    mask = msel if kpts_2d_mask is not None else None
    return (depth, mask, idx) if return_idx else (depth, mask)


def edge_depth_mean(kps, kps_3d, rot_y, K, training=False, num_k: int = K_SEL, fast: bool = False) -> torch.Tensor:
    """Fused `decode_pairs_kpts_depth(...)[0].mean(1)` (detector_infer.py:222-225, detector_loss.py:388):
    per-object depth [N] without materialising the per-edge depths.  `fast=True` (inference, large batches) trades the
    IEEE division for the hardware reciprocal: per-object depth within 1e-6 (relative) of the default, see DCD_FAST_QUOTIENT."""
    kps, kps_3d, rot, K = _prep_dgde(kps, kps_3d, rot_y, K)
    flags = FLAG_NORMALISE_2D | FLAG_SUB_B3 | (FLAG_FAST_QUOTIENT if fast and not training else 0)
    lo, hi = DGDE_CLAMP
    if not training:
        return _EdgeSolve.apply(kps, kps_3d, rot, K, lo, hi, flags, False, True)[1]
    if _num_edges(kps.shape[1]) < num_k:
        raise RuntimeError("selected index k out of range")
    return _EdgeSelectSolve.apply(kps, kps_3d, rot, K, None, lo, hi, flags, num_k, True)[3]


DOWN_RATIO = 4.0    # MODEL.BACKBONE.DOWN_RATIO of the reference; hard-coded `*4` at detector_infer.py:217


def _calib_per_object(P, N, dev):
    """[3,4] / [N,3,4] torch or numpy projection matrix (float64 in the reference) -> contiguous FP32 [N,3,4]."""
    P = torch.as_tensor(P)
    if P.dim() == 2:
        P = P.unsqueeze(0).expand(N, -1, -1)
    return f32c(P.to(dev))


def _pad_per_object(pad_size, batch_idxs, N, dev):
    pad = torch.as_tensor(pad_size, dtype=torch.float32, device=dev)
    if pad.dim() == 1:
        return pad.reshape(1, 2).expand(N, 2).contiguous()
    if batch_idxs is not None:
        return pad[batch_idxs.to(dev).long()].contiguous()
    if pad.shape[0] == 1:
        return pad.expand(N, 2).contiguous()
    return f32c(pad)


def compute_pairs_kpts_depth(pred_extra_kpts_2d, pred_bbox_points, pred_offset_3D, pad_size, pred_extra_kpts_3d, pred_rots, P,
                             batch_idxs=None, dims=None, return_locations: bool = False):
    """Fused form of PostProcessor.compute_pairs_kpts_depth (DGDE/model/head/detector_infer.py:215-227): the image-space
    keypoints `(kpts + (points + offsets)) * 4 - pad_size` are formed on the fly, solved over all keypoint pairs and
    averaged -> depth [N].  With `return_locations` also the object's 3D location (decode_location_flatten,
    anno_encoder.py:147-161, and `+ h / 2` on y when `dims` [N,3] is given, detector_infer.py:186-188) from the same launch.
    P: the image's 3x4 projection matrix (or one per object), pad_size: [2], [1,2], or [B,2] with batch_idxs.
    Inference-only (no backward; raises when an input requires grad)."""
    require_cuda(pred_extra_kpts_2d, pred_extra_kpts_3d, pred_rots, pred_bbox_points, pred_offset_3D)
    _inference_only("compute_pairs_kpts_depth", pred_extra_kpts_2d, pred_extra_kpts_3d, pred_rots, pred_bbox_points,
                    pred_offset_3D, dims)
    off, k3 = f32c(pred_extra_kpts_2d), f32c(pred_extra_kpts_3d)
    rot = f32c(pred_rots).reshape(-1)
    N, n = off.shape[0], off.shape[1]
    dev = off.device
    K = _calib_per_object(P, N, dev)
    _check_shapes(off, k3, rot, K)
    pts, ofs = f32c(pred_bbox_points), f32c(pred_offset_3D)
    pad = _pad_per_object(pad_size, batch_idxs, N, dev)
    d3 = f32c(dims) if dims is not None else None
    depth = torch.empty((N,), dtype=torch.float32, device=dev)
    loc = torch.empty((N, 3), dtype=torch.float32, device=dev) if return_locations else None
    lo, hi = DGDE_CLAMP
    if N:
        check(_lib.lib().dcd_dgde_locate_fwd(ptr(off), ptr(k3), ptr(rot), ptr(K), ptr(pts), ptr(ofs), ptr(pad),
                                             ptr(d3) if d3 is not None else 0, 0, N, n, lo, hi,
                                             FLAG_NORMALISE_2D | FLAG_SUB_B3, DOWN_RATIO, ptr(depth),
                                             ptr(loc) if loc is not None else 0, stream_ptr()), "dcd_dgde_locate_fwd")
    return (depth, loc) if return_locations else depth


# first channel of each regression group in the head's channel table of the reference configuration
# (DGDE/runs/DGDE.yaml:27-28: [4],[2],[20],[3],[3],[8,8],[1],[1],[146],[219]; layers/utils.py:22-38 Converter_key2channel)
DGDE_CHANNELS = {"3d_offset": 4, "extra_kpts_2d": 50, "extra_kpts_3d": 196}


def frame_depths_from_map(pred_regression, indexs, pred_rots, P, pad_size, dims=None, batch_idxs=None, n: int = 73,
                          channels=None, return_keypoints: bool = False):
    """POI gather + image keypoints + edge solve + mean + 3D location in ONE launch, reading the regression map itself
    (detector_infer.py:107 select_point_of_interest, :133 / :216-219 channel slices, :215-227 compute_pairs_kpts_depth,
    :186-188 decode_location_flatten): pred_regression [B,C,H,W], indexs [N] int64 heat-map positions y * W + x of the kept
    detections (select_topk), pred_rots [N] decoded yaw, P the 3x4 calibration ([3,4], per image [B,3,4] with batch_idxs, or
    per object), pad_size [2] / [B,2], dims [N,3] or None -> (depth [N], locations [N,3]) and, with return_keypoints, the
    image-space 2D keypoints [N,n,2] and template points [N,n,3] that generate_infer_data dumps (:228-236).  Inference-only."""
    require_cuda(pred_regression, indexs, pred_rots)
    _inference_only("frame_depths_from_map", pred_regression, pred_rots, dims)
    fm = f32c(pred_regression)
    B, C, H, W = fm.shape
    ch = dict(DGDE_CHANNELS)
    if channels:
        ch.update(channels)
    idx = indexs.reshape(-1).long().contiguous()
    N, dev = idx.shape[0], fm.device
    rot = f32c(pred_rots).reshape(-1)
    if rot.shape[0] != N:
        raise ValueError("pred_rots must hold one yaw per detection")
    bi = batch_idxs.to(dev).to(torch.int32).contiguous() if batch_idxs is not None else None
    if B > 1 and bi is None:
        raise ValueError("batch_idxs is required for a multi-image map")
    Pt = torch.as_tensor(P)
    if Pt.dim() == 3 and Pt.shape[0] != N and batch_idxs is not None:
        Pt = Pt.to(dev)[batch_idxs.to(dev).long()]
    K = _calib_per_object(Pt, N, dev)
    pad = _pad_per_object(pad_size, batch_idxs, N, dev)
    d3 = f32c(dims) if dims is not None else None
    depth = torch.empty((N,), dtype=torch.float32, device=dev)
    loc = torch.empty((N, 3), dtype=torch.float32, device=dev)
    kimg = torch.empty((N, n, 2), dtype=torch.float32, device=dev) if return_keypoints else None
    k3 = torch.empty((N, n, 3), dtype=torch.float32, device=dev) if return_keypoints else None
    lo, hi = DGDE_CLAMP
    if N:
        check(_lib.lib().dcd_dgde_frame_fwd(ptr(fm), ptr(idx), ptr(bi), B, C, H, W, ch["extra_kpts_2d"], ch["extra_kpts_3d"],
                                            ch["3d_offset"], ptr(rot), ptr(K), ptr(pad), ptr(d3), N, n, lo, hi,
                                            FLAG_NORMALISE_2D | FLAG_SUB_B3, DOWN_RATIO, ptr(depth), ptr(loc), ptr(kimg), ptr(k3),
                                            stream_ptr()), "dcd_dgde_frame_fwd")
    return (depth, loc, kimg, k3) if return_keypoints else (depth, loc)


def decode_location_flatten(points, offsets, depths, P, pad_size, batch_idxs=None):
    """Anno_Encoder.decode_location_flatten (DGDE/model/anno_encoder.py:147-161) with the calibration given as the
    3x4 matrix of the image ([3,4]) or per object ([N,3,4]) instead of Calibration objects -> locations [N,3].
    Inference-only (no backward; raises when an input requires grad)."""
    require_cuda(points, offsets, depths)
    _inference_only("decode_location_flatten", points, offsets, depths)
    pts, ofs, dep = f32c(points), f32c(offsets), f32c(depths).reshape(-1)
    N, dev = pts.shape[0], pts.device
    K = _calib_per_object(P, N, dev)
    pad = _pad_per_object(pad_size, batch_idxs, N, dev)
    loc = torch.empty((N, 3), dtype=torch.float32, device=dev)
    if N:
        check(_lib.lib().dcd_dgde_locate_fwd(0, 0, 0, ptr(K), ptr(pts), ptr(ofs), ptr(pad), 0, ptr(dep), N, 2, 0.0, 0.0, 0,
                                             DOWN_RATIO, 0, ptr(loc), stream_ptr()), "dcd_dgde_locate_fwd")
    return loc


def select_point_of_interest(batch, index, feature_maps, validate: bool = False):
    """Drop-in for DGDE/model/layers/utils.py:120-145: feature_maps [B,C,H,W], index [B,K] flattened positions (or
    [B,K,2] (x, y) points) -> [B,K,C], reading only the selected values (no NHWC copy of the map).  An index outside the
    map yields NaN; `validate=True` checks the range first (a host synchronisation) and raises like torch.gather.
    Inference-only (no backward; raises when feature_maps requires grad)."""
    require_cuda(feature_maps, index)
    _inference_only("select_point_of_interest", feature_maps)
    fm = f32c(feature_maps)
    B, C, H, W = fm.shape
    if index.dim() == 3:
        index = index[:, :, 1] * W + index[:, :, 0]
    idx = index.reshape(batch, -1).long().contiguous()
    if validate and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= H * W):
        raise RuntimeError("index out of range")
    out = torch.empty((B, idx.shape[1], C), dtype=torch.float32, device=fm.device)
    if out.numel():
        check(_lib.lib().dcd_poi_gather_fwd(ptr(fm), ptr(idx), B, idx.shape[1], C, H * W, ptr(out), stream_ptr()),
              "dcd_poi_gather_fwd")
    return out


DEPTH_RANGE = (0.1, 100.0)   # MODEL.HEAD.DEPTH_RANGE (DGDE/config/defaults.py:196)
KP_DEPTH_EPS = 1e-3          # Anno_Encoder.EPS (anno_encoder.py:17)


def decode_depth_from_keypoints_batch(pred_keypoints, pred_dimensions, P, batch_idxs=None):
    """Anno_Encoder.decode_depth_from_keypoints_batch (DGDE/model/anno_encoder.py:193-224) with the calibration as the
    image's 3x4 matrix ([3,4]; [B,3,4] with batch_idxs; or per object [N,3,4]) -> [N,3] (centre, corner_02, corner_13)."""
    require_cuda(pred_keypoints, pred_dimensions)
    _inference_only("decode_depth_from_keypoints_batch", pred_keypoints, pred_dimensions)
    kp, dm = f32c(pred_keypoints).reshape(-1, 10, 2), f32c(pred_dimensions)
    N, dev = kp.shape[0], kp.device
    P = torch.as_tensor(P)
    if P.dim() == 3 and batch_idxs is not None and P.shape[0] != N:
        P = P.to(dev)[batch_idxs.to(dev).long()]
    K = _calib_per_object(P, N, dev)
    out = torch.empty((N, 3), dtype=torch.float32, device=dev)
    if N:
        check(_lib.lib().dcd_dgde_depth_ensemble_fwd(ptr(kp), ptr(dm), ptr(K), 0, 0, 0, 0, N, DOWN_RATIO, KP_DEPTH_EPS,
                                                     DEPTH_RANGE[0], DEPTH_RANGE[1], ptr(out), 0, 0, 0, 0, stream_ptr()),
              "dcd_dgde_depth_ensemble_fwd")
    return out


def depth_ensemble(pred_keypoints, pred_dimensions, P, keypoint_log_uncertainty, direct_depths=None,
                   direct_log_uncertainty=None, scores=None):
    """Fused detector_infer.py:141-171,197-203: keypoint depths, uncertainties = exp(channels), inverse-uncertainty soft
    ensemble -> dict(keypoint_depths [N,3], depth [N], depth_error [N], min_uncertainty [N] int64, scores [N,1] or None)."""
    require_cuda(pred_keypoints, pred_dimensions, keypoint_log_uncertainty)
    _inference_only("depth_ensemble", pred_keypoints, pred_dimensions, keypoint_log_uncertainty, direct_depths,
                    direct_log_uncertainty, scores)
    kp, dm = f32c(pred_keypoints).reshape(-1, 10, 2), f32c(pred_dimensions)
    N, dev = kp.shape[0], kp.device
    K = _calib_per_object(P, N, dev)
    luk = f32c(keypoint_log_uncertainty).reshape(N, 3)
    dd = f32c(direct_depths).reshape(-1) if direct_depths is not None else None
    lud = f32c(direct_log_uncertainty).reshape(-1) if direct_log_uncertainty is not None else None
    if (dd is None) != (lud is None):
        raise RuntimeError("direct depth and its uncertainty channel go together")
    sc = f32c(scores).reshape(-1) if scores is not None else None
    kd = torch.empty((N, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((N,), dtype=torch.float32, device=dev)
    err = torch.empty((N,), dtype=torch.float32, device=dev)
    amax = torch.empty((N,), dtype=torch.int64, device=dev)
    so = torch.empty((N,), dtype=torch.float32, device=dev) if sc is not None else None
    if N:
        check(_lib.lib().dcd_dgde_depth_ensemble_fwd(ptr(kp), ptr(dm), ptr(K), ptr(dd) if dd is not None else 0,
                                                     ptr(lud) if lud is not None else 0, ptr(luk),
                                                     ptr(sc) if sc is not None else 0, N, DOWN_RATIO, KP_DEPTH_EPS,
                                                     DEPTH_RANGE[0], DEPTH_RANGE[1], ptr(kd), ptr(depth), ptr(err), ptr(amax),
                                                     ptr(so) if so is not None else 0, stream_ptr()),
              "dcd_dgde_depth_ensemble_fwd")
    return {"keypoint_depths": kd, "depth": depth, "depth_error": err, "min_uncertainty": amax,
            "scores": so.reshape(-1, 1) if so is not None else None}


def ray_rescale(raw_location, pred_depth, dim):
    """GMW/main.py:542-547: detector location [N,3] moved along its viewing ray to the GMW depth (dim [N,3] = (h,w,l))."""
    require_cuda(raw_location, pred_depth, dim)
    _inference_only("ray_rescale", raw_location, pred_depth, dim)
    rl, pd, dm = f32c(raw_location), f32c(pred_depth).reshape(-1), f32c(dim)
    out = torch.empty_like(rl)
    if rl.shape[0]:
        check(_lib.lib().dcd_gmw_ray_rescale_fwd(ptr(rl), ptr(pd), ptr(dm), rl.shape[0], ptr(out), stream_ptr()),
              "dcd_gmw_ray_rescale_fwd")
    return out


def compute_z(kpts_2d, kpts_3d, pred_rot, num_k: int = K_SEL) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for GMW/main.py:373-416: (Z_v_raw [b,E] clamped to [0.1,80], good_idx [b,1500] int64)."""
    require_cuda(kpts_2d, kpts_3d, pred_rot)
    kps, kps_3d = f32c(kpts_2d), f32c(kpts_3d)
    rot = f32c(pred_rot).reshape(-1)
    N, n = _check_shapes(kps, kps_3d, rot)
    if _num_edges(n) < num_k:
        raise RuntimeError("selected index k out of range")
    lo, hi = GMW_CLAMP
    with torch.no_grad():
        Z, _ = _EdgeSolve.apply(kps, kps_3d, rot, None, lo, hi, 0, True, False)
        idx = torch.empty((N, num_k), dtype=torch.int64, device=kps.device)
        if N:
            check(_lib.lib().dcd_edge_select_fwd(ptr(kps), ptr(kps_3d), ptr(rot), 0, 0, N, n, num_k, lo, hi, 0,
                                                 ptr(idx), 0, 0, 0, stream_ptr()), "dcd_edge_select_fwd")
    return Z, idx


# ---------------------------------------------------------------------------------------------
# GMW edge weights + aggregation
# ---------------------------------------------------------------------------------------------
POISON_WORKSPACES = os.environ.get("DCD_B200_POISON", "0") == "1"   # debug/tests: NaN-fill every workspace first


def _alloc_bytes(nbytes: int, dev) -> torch.Tensor:
    t = torch.empty(((nbytes + 255) // 256 * 64,), dtype=torch.float32, device=dev)
    if POISON_WORKSPACES:
        t.fill_(float("nan"))       # the kernels must never read what they did not write
    return t


class _GmwWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kpts_2d, kpts_3d, params4, params6, depth, need_grad):
        N, n = kpts_2d.shape[0], kpts_2d.shape[1]
        E = _num_edges(n)
        dev = kpts_2d.device
        L = _lib.lib()
        reg_w = torch.empty((N, E), dtype=torch.float32, device=dev)
        save = 1 if need_grad else 0
        ws = None
        if N:
            nbytes = L.dcd_gmw_workspace_bytes(N, n, depth, save)
            ws = _alloc_bytes(nbytes, dev)
            check(L.dcd_gmw_weights_fwd(ptr(kpts_2d), ptr(kpts_3d), ptr(params4), ptr(params6), N, n, depth, save,
                                        ptr(reg_w), 0, 0, ptr(ws), ws.numel() * 4, stream_ptr()), "dcd_gmw_weights_fwd")
        if need_grad:
            ctx.save_for_backward(kpts_2d, kpts_3d, params4, params6)
            ctx.ws = ws
            ctx.depth = depth
        return reg_w

    @staticmethod
    def backward(ctx, g_reg_w):
        kpts_2d, kpts_3d, params4, params6 = ctx.saved_tensors
        N, n = kpts_2d.shape[0], kpts_2d.shape[1]
        L = _lib.lib()
        g4 = torch.zeros_like(params4)
        g6 = torch.zeros_like(params6)
        if N:
            g = f32c(g_reg_w)
            scratch = _alloc_bytes(L.dcd_gmw_bwd_scratch_bytes(N, n, ctx.depth), g.device)
            check(L.dcd_gmw_weights_bwd(ptr(kpts_2d), ptr(kpts_3d), ptr(params4), ptr(params6), N, n, ctx.depth,
                                        ptr(g), 0, 0, ptr(g4), ptr(g6), ptr(ctx.ws), ctx.ws.numel() * 4,
                                        ptr(scratch), scratch.numel() * 4, stream_ptr()), "dcd_gmw_weights_bwd")
        ctx.ws = None
        return None, None, g4, g6, None, None


SINKHORN_CG_ITERATIONS, SINKHORN_CG_TOLERANCE = 32, 1e-7


class _GmwWeightsTransport(torch.autograd.Function):
    """GMW.forward with both outputs like the reference (GMW/model/model.py:195-207): reg_weights [b,E] and the Sinkhorn
    transport plan edge_P [b,E,E], both differentiable w.r.t. the parameter blobs.  Backward = implicit gradient of the
    Sinkhorn fixed point (optimal_transport.py:75-128 as one conjugate-gradient solve) -> feature gradients -> MLP backward."""

    @staticmethod
    def forward(ctx, kpts_2d, kpts_3d, params4, params6, depth, need_grad, lam, tol, iters):
        N, n = kpts_2d.shape[0], kpts_2d.shape[1]
        E = _num_edges(n)
        dev = kpts_2d.device
        L = _lib.lib()
        reg_w = torch.empty((N, E), dtype=torch.float32, device=dev)
        P = torch.empty((N, E, E), dtype=torch.float32, device=dev)
        u = torch.empty((N, E), dtype=torch.float32, device=dev)
        v = torch.empty((N, E), dtype=torch.float32, device=dev)
        f4 = torch.empty((N, 128, E), dtype=torch.float32, device=dev)
        f6 = torch.empty((N, 128, E), dtype=torch.float32, device=dev)
        save = 1 if need_grad else 0
        ws = None
        if N:
            ws = _alloc_bytes(L.dcd_gmw_workspace_bytes(N, n, depth, save), dev)
            check(L.dcd_gmw_weights_fwd(ptr(kpts_2d), ptr(kpts_3d), ptr(params4), ptr(params6), N, n, depth, save,
                                        ptr(reg_w), ptr(f4), ptr(f6), ptr(ws), ws.numel() * 4, stream_ptr()), "dcd_gmw_weights_fwd")
            tws = _alloc_bytes(L.dcd_gmw_transport_workspace_bytes(N, n), dev)
            check(L.dcd_gmw_transport_fwd(ptr(f4), ptr(f6), N, n, lam, tol, iters, ptr(P), ptr(u), ptr(v), 0,
                                          ptr(tws), tws.numel() * 4, stream_ptr()), "dcd_gmw_transport_fwd")
        if need_grad:
            ctx.save_for_backward(kpts_2d, kpts_3d, params4, params6, f4, f6, P, u, v)
            ctx.ws = ws
            ctx.cfg = (depth, lam)
        return reg_w, P

    @staticmethod
    def backward(ctx, g_reg_w, g_P):
        kpts_2d, kpts_3d, params4, params6, f4, f6, P, u, v = ctx.saved_tensors
        depth, lam = ctx.cfg
        N, n = kpts_2d.shape[0], kpts_2d.shape[1]
        L = _lib.lib()
        g4 = torch.zeros_like(params4)
        g6 = torch.zeros_like(params6)
        if N and (g_reg_w is not None or g_P is not None):
            gn4 = gn6 = None
            if g_P is not None:
                gP = f32c(g_P)
                gn4, gn6 = torch.empty_like(f4), torch.empty_like(f6)
                tws = _alloc_bytes(L.dcd_gmw_transport_bwd_workspace_bytes(N, n), gP.device)
                check(L.dcd_gmw_transport_bwd(ptr(f4), ptr(f6), ptr(P), ptr(u), ptr(v), ptr(gP), N, n, lam,
                                              SINKHORN_CG_ITERATIONS, SINKHORN_CG_TOLERANCE, ptr(gn4), ptr(gn6), 0,
                                              ptr(tws), tws.numel() * 4, stream_ptr()), "dcd_gmw_transport_bwd")
            g = f32c(g_reg_w) if g_reg_w is not None else None
            scratch = _alloc_bytes(L.dcd_gmw_bwd_scratch_bytes(N, n, depth), kpts_2d.device)
            check(L.dcd_gmw_weights_bwd(ptr(kpts_2d), ptr(kpts_3d), ptr(params4), ptr(params6), N, n, depth,
                                        ptr(g), ptr(gn4), ptr(gn6), ptr(g4), ptr(g6), ptr(ctx.ws), ctx.ws.numel() * 4,
                                        ptr(scratch), scratch.numel() * 4, stream_ptr()), "dcd_gmw_weights_bwd")
        ctx.ws = None
        return None, None, g4, g6, None, None, None, None, None


class _GmwAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reg_w, depths, idx):
        N, E = reg_w.shape
        k = idx.shape[1]
        out = torch.empty((N,), dtype=torch.float32, device=reg_w.device)
        if N:
            check(_lib.lib().dcd_gmw_aggregate_fwd(ptr(reg_w), ptr(depths), ptr(idx), N, E, k, 0, ptr(out), 0,
                                                   stream_ptr()), "dcd_gmw_aggregate_fwd")
        ctx.save_for_backward(reg_w, depths, idx)
        return out

    @staticmethod
    def backward(ctx, g_out):
        reg_w, depths, idx = ctx.saved_tensors
        N, E = reg_w.shape
        k = idx.shape[1]
        g_w = torch.empty_like(reg_w)
        g_z = torch.empty_like(depths) if ctx.needs_input_grad[1] else None
        if N:
            check(_lib.lib().dcd_gmw_aggregate_bwd(ptr(reg_w), ptr(depths), ptr(idx), N, E, k, 0, ptr(f32c(g_out)),
                                                   ptr(g_w), ptr(g_z), stream_ptr()), "dcd_gmw_aggregate_bwd")
        return g_w, g_z, None


class GMW(nn.Module):
    """Drop-in for the reference `GMW` module's regression branch (GMW/model/model.py:103-207).

    forward(kpts_2d, kpts_3d, pred_rot, args) -> (reg_weights [b,E], edge_P); `pred_rot` and `args`
    are ignored as in the reference.  edge_P (the Sinkhorn correspondence matrix
    of the classification branch) is None unless the module attribute `with_edge_P`
    is set (dcd_b200.patch.install sets it: both loops of GMW/main.py consume edge_P, :456 and :526); then it
    is the [b,E,E] transport plan, differentiable like the reference's (implicit Sinkhorn gradient,
    optimal_transport.py:75-128).  `edge_transport` gives the plan's (sum, trace) without materialising it.
    Parameters live in two flat blobs; `load_state_dict`/`state_dict` of the *reference* format
    are available through `load_reference_state_dict` / `reference_state_dict`.
    """

    def __init__(self, args=None, depth: int = NET_DEPTH):
        super().__init__()
        self.depth = depth
        self.params4 = nn.Parameter(torch.zeros(blob_size(4, depth)))
        self.params6 = nn.Parameter(torch.zeros(blob_size(6, depth)))
        self.with_edge_P = False

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]) -> "GMW":
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}   # main.py:286-289
        with torch.no_grad():
            for (name, cin), p in zip(NET_NAMES, (self.params4, self.params6)):
                p.copy_(pack_state_dict(sd, name, cin, self.depth).to(p.device))
        return self

    def reference_state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for (name, cin), p in zip(NET_NAMES, (self.params4, self.params6)):
            out.update({k: v.detach().clone() for k, v in unpack_blob(p.data, name, cin, self.depth).items()})
        return out

    def reference_grads(self) -> Dict[str, torch.Tensor]:
        out = {}
        for (name, cin), p in zip(NET_NAMES, (self.params4, self.params6)):
            out.update({k: v.clone() for k, v in unpack_blob(p.grad, name, cin, self.depth).items()})
        return out

    SINKHORN_LAMBDA, SINKHORN_TOLERANCE, SINKHORN_ITERATIONS = 10.0, 1e-9, 100     # GMW/model/model.py:117-119

    @torch.no_grad()
    def edge_transport(self, kpts_2d, kpts_3d, materialise: bool = True, params=None):
        """Correspondence branch, forward only (GMW/model/model.py:170-192): returns
        (reg_weights [b,E], edge_P [b,E,E] or None, cls_terms [b,2] = (sum P, trace P)).
        correspondenceLoss(edge_P, eye) of GMW/main.py:456-457 is `(cls_terms[:, 0] - 2 * cls_terms[:, 1]).mean()`;
        with materialise=False the 4 E^2 bytes per object of edge_P are never written.  No gradient."""
        p4, p6 = params if params is not None else (self.params4, self.params6)
        require_cuda(kpts_2d, kpts_3d, p4, p6)
        k2, k3 = f32c(kpts_2d), f32c(kpts_3d)
        N, n = k2.shape[0], k2.shape[1]
        E = _num_edges(n)
        dev = k2.device
        L = _lib.lib()
        reg_w = torch.empty((N, E), dtype=torch.float32, device=dev)
        P = torch.empty((N, E, E), dtype=torch.float32, device=dev) if materialise else None
        sums = torch.empty((N, 2), dtype=torch.float32, device=dev)
        if N:
            f4 = torch.empty((N, 128, E), dtype=torch.float32, device=dev)
            f6 = torch.empty((N, 128, E), dtype=torch.float32, device=dev)
            ws = _alloc_bytes(L.dcd_gmw_workspace_bytes(N, n, self.depth, 0), dev)
            check(L.dcd_gmw_weights_fwd(ptr(k2), ptr(k3), ptr(p4), ptr(p6), N, n, self.depth, 0,
                                        ptr(reg_w), ptr(f4), ptr(f6), ptr(ws), ws.numel() * 4, stream_ptr()), "dcd_gmw_weights_fwd")
            del ws
            tws = _alloc_bytes(L.dcd_gmw_transport_workspace_bytes(N, n), dev)
            check(L.dcd_gmw_transport_fwd(ptr(f4), ptr(f6), N, n, self.SINKHORN_LAMBDA, self.SINKHORN_TOLERANCE,
                                          self.SINKHORN_ITERATIONS, ptr(P) if P is not None else 0, 0, 0, ptr(sums),
                                          ptr(tws), tws.numel() * 4, stream_ptr()), "dcd_gmw_transport_fwd")
        return reg_w, P, sums

    def forward(self, kpts_2d, kpts_3d, pred_rot=None, args=None):
        return self.forward_blobs(kpts_2d, kpts_3d, self.params4, self.params6)

    def forward_blobs(self, kpts_2d, kpts_3d, params4, params6):
        """forward() on explicit parameter blobs (dcd_b200.patch passes blobs packed from the LIVE reference parameters,
        so that their optimizer, checkpoints and DDP hooks keep working unchanged)."""
        require_cuda(kpts_2d, kpts_3d, params4, params6)
        k2, k3 = f32c(kpts_2d), f32c(kpts_3d)
        if k2.dim() != 3 or k2.shape[-1] != 2 or tuple(k3.shape) != (k2.shape[0], k2.shape[1], 3):
            raise ValueError("kpts_2d must be [b,n,2] and kpts_3d [b,n,3]")
        need_grad = torch.is_grad_enabled() and (params4.requires_grad or params6.requires_grad)
        if self.with_edge_P:
            return _GmwWeightsTransport.apply(k2, k3, params4, params6, self.depth, need_grad, self.SINKHORN_LAMBDA,
                                              self.SINKHORN_TOLERANCE, self.SINKHORN_ITERATIONS)
        reg_w = _GmwWeights.apply(k2, k3, params4, params6, self.depth, need_grad)
        if _CHECK_FINITE and not bool(torch.isfinite(reg_w).all()):
            # the tensor-core path carries activations as FP16 hi+lo pairs: |activation| must stay below 65504
            raise FloatingPointError("dcd_b200.GMW: non-finite edge weights (activation outside the FP16 hi/lo range?)")
        return reg_w, None


def compute_reg_loss(pre_depths, edge_weight, gt_depth, good_idx=None, validate: bool = False):
    """Drop-in for GMW/main.py:364-371 -> (reg_loss, Z_select_weighted).  Shapes are checked; `validate=True` also checks
    the index range (a host synchronisation; the kernels clamp an out-of-range index instead of faulting)."""
    if good_idx is None:
        raise UnboundLocalError("compute_reg_loss needs good_idx (the reference fails without it too)")
    require_cuda(pre_depths, edge_weight, good_idx)
    if edge_weight.dim() != 2 or tuple(pre_depths.shape) != tuple(edge_weight.shape):
        raise RuntimeError("compute_reg_loss: pre_depths %s and edge_weight %s must both be [b,E]"
                           % (tuple(pre_depths.shape), tuple(edge_weight.shape)))
    if good_idx.dim() != 2 or good_idx.shape[0] != edge_weight.shape[0] or good_idx.shape[1] > edge_weight.shape[1]:
        raise RuntimeError("compute_reg_loss: good_idx must be [b,k] with k <= E, got %s" % (tuple(good_idx.shape),))
    idx = good_idx.to(torch.int64).contiguous()
    if validate and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= edge_weight.shape[1]):
        raise RuntimeError("index out of range in compute_reg_loss (torch.gather raises here too)")
    Z = _GmwAggregate.apply(f32c(edge_weight), f32c(pre_depths), idx)
    reg_loss = (Z - gt_depth).abs().mean()
    return reg_loss, Z


def gmw_weighted_depth(kpts_2d, kpts_3d, pred_rot, model: GMW, chunk: int = 1024, num_k: int = K_SEL,
                       return_idx: bool = False):
    """Fused GMW inference (GMW/main.py:524-533): per-object softmax-weighted depth [b]."""
    require_cuda(kpts_2d, kpts_3d, pred_rot, model.params4)
    k2, k3 = f32c(kpts_2d), f32c(kpts_3d)
    rot = f32c(pred_rot).reshape(-1)
    N, n = _check_shapes(k2, k3, rot)
    if _num_edges(n) < num_k:
        raise RuntimeError("selected index k out of range")
    L = _lib.lib()
    out = torch.empty((N,), dtype=torch.float32, device=k2.device)
    idx = torch.empty((N, num_k), dtype=torch.int64, device=k2.device) if return_idx else None
    if N:
        chunk = max(1, min(chunk, N))
        ws = _alloc_bytes(L.dcd_gmw_depth_workspace_bytes(N, n, model.depth, chunk), k2.device)
        lo, hi = GMW_CLAMP
        with torch.no_grad():
            check(L.dcd_gmw_depth_fwd(ptr(k2), ptr(k3), ptr(rot), ptr(model.params4), ptr(model.params6), N, n,
                                      model.depth, num_k, lo, hi, chunk, ptr(out), ptr(idx), 0, ptr(ws),
                                      ws.numel() * 4, stream_ptr()), "dcd_gmw_depth_fwd")
    return (out, idx) if return_idx else out

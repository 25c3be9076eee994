"""ORACLE — CPU restatement of DCD's densely-constrained-depth hot path (TEST INFRASTRUCTURE).

This module is the checker, never the product: only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product (`dcd_b200/`) never imports anything under `oracle/` and has no CPU fallback.

The reference is pure Python/PyTorch, so the restatement is written with the same torch
elementwise operations in the same order (that is what makes it bit-identical on CPU);
it is device-agnostic, so GPU tests can also evaluate it with torch-CUDA ops as the
"reference run on CUDA" of SURVEY.md section 8c.

Parity pin: the reference holds NO golden vectors or tests for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself, produced in
the build container by `oracle/make_golden.py` (which imports the unmodified reference
through `oracle/ref_loader.py`) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` checks every function below against those fixtures.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

K_SEL = 1500          # DGDE/model/anno_encoder.py:378, GMW/main.py:413
CN_EPS = 1e-3         # GMW/model/yi2018cvpr/ops.py:14
NET_DEPTH = 12        # GMW/model/yi2018cvpr/config.py:69
NET_CH = 128          # GMW/model/yi2018cvpr/config.py:72


# ----------------------------------------------------------------------------------
# edge enumeration  (anno_encoder.py:313-324, GMW/main.py:351-362, GMW/model/model.py:121-135)
# ----------------------------------------------------------------------------------
def edge_pairs(n: int, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-major strict upper triangle: for i in range(n): for j in range(i+1, n)."""
    ii, jj = [], []
    for i in range(n):
        for j in range(i + 1, n):
            ii.append(i)
            jj.append(j)
    return (torch.tensor(ii, dtype=torch.int64, device=device),
            torch.tensor(jj, dtype=torch.int64, device=device))


def get_up(matrix: torch.Tensor, faithful: bool = False) -> torch.Tensor:
    """[b,n,n] -> [b,E] strict upper triangle, row-major (anno_encoder.py:313-324).

    faithful=True walks the pairs one column at a time like the reference does (this is
    what its run time consists of, so the CPU baseline uses it); the default gathers.
    The reference returns a float32 tensor whatever the input dtype (torch.zeros default);
    the restatement keeps the input dtype for floating inputs so FP64 evaluation works.
    """
    b, n = matrix.shape[0], matrix.shape[1]
    out_dtype = matrix.dtype if matrix.dtype.is_floating_point else torch.float32
    if faithful:
        upper = torch.zeros((b, n * (n - 1) // 2), dtype=out_dtype, device=matrix.device)
        e = 0
        for i in range(n):
            for j in range(i + 1, n):
                upper[:, e] = matrix[:, i, j]
                e += 1
        return upper
    ii, jj = edge_pairs(n, matrix.device)
    return matrix[:, ii, jj].to(out_dtype)


# ----------------------------------------------------------------------------------
# edge-depth solve  (anno_encoder.py:326-390 ; GMW/main.py:373-416)
# ----------------------------------------------------------------------------------
def _edge_terms(v: torch.Tensor, kps_3d: torch.Tensor, rot: torch.Tensor, faithful: bool):
    """Shared body: per-edge H, V for normalised vertical coordinate v [N,n].

    anno_encoder.py:339-369 / main.py:379-404: only the odd (v) rows of B and C are used.
    """
    X = kps_3d[:, :, 0:1]
    Y = kps_3d[:, :, 1:2]
    Z = kps_3d[:, :, 2:3]
    cosori = torch.cos(rot).unsqueeze(-1).expand_as(X)
    sinori = torch.sin(rot).unsqueeze(-1).expand_as(X)
    C = X * sinori - Z * cosori                       # :346-347  (two products, then subtract)
    H_1 = Y                                           # :345,349
    H_2 = v.unsqueeze(-1) * C                         # :350-353
    n = v.shape[1]
    if faithful:
        H1_1 = H_1.expand(-1, n, n)
        H1_2 = H_2.expand(-1, n, n)
        V1 = v.unsqueeze(-1).expand(-1, n, n)
        H_mat = (H1_1 - H1_1.permute(0, 2, 1)) + (H1_2 - H1_2.permute(0, 2, 1))   # :367
        V_mat = V1 - V1.permute(0, 2, 1)                                          # :369
        Zm = H_mat.abs() / V_mat.abs().clamp_min(1e-10)                           # :371
        return get_up(Zm, True), get_up(V_mat, True)
    ii, jj = edge_pairs(n, v.device)
    y = H_1[:, :, 0]
    h2 = H_2[:, :, 0]
    H = (y[:, ii] - y[:, jj]) + (h2[:, ii] - h2[:, jj])
    V = v[:, ii] - v[:, jj]
    return H.abs() / V.abs().clamp_min(1e-10), V


def normalise_v(kps: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """anno_encoder.py:331-334 (v column only; the u column is dead)."""
    return (kps[:, :, 1] - K[:, None, 1, 2]) / K[:, None, 1, 1]


def canonical_topk(absV: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k largest |V| sorted by (|V| descending, edge id ascending).

    torch.topk (anno_encoder.py:379) breaks ties arbitrarily; this is the canonical rule
     the CUDA kernel implements bit-exactly.
    """
    order = torch.sort(absV, dim=-1, descending=True, stable=True).indices
    return order[:, :k]


def topk_is_canonical_equivalent(absV: torch.Tensor, idx: torch.Tensor, k: int) -> bool:
    """True when `idx` (e.g. from torch.topk) selects the same keys in the same sorted order as
    the canonical rule, i.e. differs at most by permutations inside groups of equal keys and by
    the choice among equal keys at the rank-k boundary."""
    can = canonical_topk(absV, k)
    return bool(torch.equal(absV.gather(-1, idx), absV.gather(-1, can)))


def decode_pairs_kpts_depth(kps, kps_3d, rot_y, K, training=False, kpts_2d_mask=None,
                            faithful=False, canonical=True, num_k=K_SEL,
                            return_idx=False):
    """DGDE edge solve, anno_encoder.py:326-390.

    kps [N,n,2] pixels, kps_3d [N,n,3], rot_y [N,1], K [N,3,4]; returns (depth_all, depth_mask)
    exactly like the reference: [N,E] when not training, [N,num_k] (sorted by |V|) when training;
    depth_mask float32 0/1 or None.  `canonical` selects the tie rule for the top-k.
    """
    b3 = K[:, 2, 3]
    v = normalise_v(kps, K).to(kps.dtype)
    Zraw, V = _edge_terms(v, kps_3d, rot_y, faithful)
    Zraw = Zraw.clamp_min(2.).clamp_max(80)                                        # :375
    depth_mask = None
    if kpts_2d_mask is not None:
        m = kpts_2d_mask
        ii, jj = edge_pairs(m.shape[1], m.device)
        depth_mask = (m[:, ii] * m[:, jj]).to(torch.float32)                       # :362-365
    good_idx = None
    if training:
        absV = V.abs()
        good_idx = canonical_topk(absV, num_k) if canonical else torch.topk(absV, num_k, dim=-1)[1]
        depth_all = Zraw.gather(-1, good_idx)                                      # :380
        if depth_mask is not None:
            depth_mask = depth_mask.gather(-1, good_idx)                           # :382
    else:
        depth_all = Zraw
    depth_all = depth_all - b3.unsqueeze(-1).to(depth_all.dtype)                   # :385
    if return_idx:
        return depth_all, depth_mask, good_idx
    return depth_all, depth_mask


def compute_z(kpts_2d, kpts_3d, pred_rot, faithful=False, canonical=True, num_k=K_SEL):
    """GMW edge solve, GMW/main.py:373-416: pre-normalised 2D points, clamp [0.1, 80], no b3."""
    v = kpts_2d[:, :, 1]
    Zraw, V = _edge_terms(v, kpts_3d, pred_rot, faithful)
    Zraw = Zraw.clamp_min(0.1).clamp_max(80.)                                      # :410
    absV = V.abs()
    good_idx = canonical_topk(absV, num_k) if canonical else torch.topk(absV, num_k, dim=-1)[1]
    return Zraw, good_idx


def edge_abs_v(v: torch.Tensor) -> torch.Tensor:
    """|V| keys of every edge for normalised v [N,n] (the top-k ranking key)."""
    ii, jj = edge_pairs(v.shape[1], v.device)
    return (v[:, ii] - v[:, jj]).abs()


# ----------------------------------------------------------------------------------
# GMW edge features, edge MLP, weights, aggregation
# ----------------------------------------------------------------------------------
def edge_expand(f: torch.Tensor) -> torch.Tensor:
    """GMW/model/model.py:153-163: [b,n,c] -> [b,E,2c] = concat(kp_i, kp_j) in edge order."""
    ii, jj = edge_pairs(f.shape[1], f.device)
    return torch.cat((f[:, ii], f[:, jj]), dim=-1)


def context_norm(x: torch.Tensor) -> torch.Tensor:
    """GMW/model/yi2018cvpr/ops.py:12-19, x [b,C,E]; unbiased variance over the edge axis."""
    m = torch.mean(x, 2, keepdim=True)
    v = torch.var(x, 2, keepdim=True)
    inv = 1. / torch.sqrt(v + CN_EPS)
    return (x - m) * inv


def _conv(x, sd, key):
    return F.conv1d(x, sd[key + ".0.weight"], sd[key + ".0.bias"])


def edge_net(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, depth: int = NET_DEPTH,
             keep: Optional[list] = None) -> torch.Tensor:
    """yi2018cvpr/model.py:63-67 + ops.py:125-131, x [b,Cin,E] -> [b,128,E].

    Block = preconv (plain conv) -> conv1 (conv + CN) -> conv2 (conv + CN) -> ReLU -> + input
    (no BatchNorm anywhere: SURVEY fact 8).  `sd` uses the reference state_dict key names.
    """
    x = _conv(x, sd, prefix + ".conv_in")
    for k in range(depth):
        xorg = x
        p = prefix + ".conv_%d" % k
        x = _conv(x, sd, p + ".preconv")
        x = context_norm(_conv(x, sd, p + ".conv1"))
        x = context_norm(_conv(x, sd, p + ".conv2"))
        x = F.relu(x) + xorg
        if keep is not None:
            keep.append(x)
    return x


def edge_distance_diag(f4: torch.Tensor, f6: torch.Tensor) -> torch.Tensor:
    """Diagonal of pairwiseL2Dist (GMW/model/model.py:28-35) for unit-normalised f4,f6 [b,E,128].

    Same expansion and association as the reference: ((|c|^2) + (-2)(a.c)) + |a|^2, clamp 1e-30, sqrt.
    """
    a2 = f4.pow(2).sum(dim=-1)
    c2 = f6.pow(2).sum(dim=-1)
    ac = (f4 * f6).sum(dim=-1)
    return ((c2 + (-2.0) * ac) + a2).clamp_min(1e-30).sqrt()


def gmw_reg_weights(kpts_2d, kpts_3d, sd, depth: int = NET_DEPTH, full_matrix: bool = False,
                    return_feats: bool = False):
    """GMW.forward's first output (GMW/model/model.py:195-207 -> :170-181): reg_weights [b,E].

    full_matrix=True evaluates the E x E baddbmm like the reference (the only way the reference
    can reach the diagonal; used for the CPU baseline); otherwise diagonal only.
    """
    f4 = edge_expand(kpts_2d)
    f6 = edge_expand(kpts_3d)
    f4 = edge_net(f4.transpose(-2, -1), sd, "FeatureExtractor4d", depth).transpose(-2, -1)
    f6 = edge_net(f6.transpose(-2, -1), sd, "FeatureExtractor6d", depth).transpose(-2, -1)
    f4 = F.normalize(f4, p=2, dim=-1)
    f6 = F.normalize(f6, p=2, dim=-1)
    if full_matrix:
        x1n = f4.pow(2).sum(dim=-1, keepdim=True)
        x2n = f6.pow(2).sum(dim=-1, keepdim=True)
        M = torch.baddbmm(x2n.transpose(-2, -1), f4, f6.transpose(-2, -1), alpha=-2
                          ).add_(x1n).clamp_min_(1e-30).sqrt_()
        w = 1. / M.diagonal(offset=0, dim1=-2, dim2=-1)
    else:
        w = 1. / edge_distance_diag(f4, f6)
    if return_feats:
        return w, f4, f6
    return w


def pairwise_l2_dist(x1, x2):
    """GMW/model/model.py:17-36."""
    x1_norm2 = x1.pow(2).sum(dim=-1, keepdim=True)
    x2_norm2 = x2.pow(2).sum(dim=-1, keepdim=True)
    return torch.baddbmm(x2_norm2.transpose(-2, -1), x1, x2.transpose(-2, -1), alpha=-2).add_(x1_norm2).clamp_min_(1e-30).sqrt_()


def sinkhorn(M, r, c, lmbda=10.0, tolerance=1e-9, max_iterations=100, max_distance=5.0):
    """GMW/lib/optimal_transport.py:52-72 (RegularisedTransportFn.sinkhorn) with tensor r, c."""
    K = (-lmbda * M.clamp_max(max_distance)).exp()
    r = r.unsqueeze(-1)
    c = c.unsqueeze(-1)
    u = r.clone()
    u_prev = torch.ones_like(u)
    for _ in range(max_iterations):
        if torch.all(torch.isclose(u, u_prev, atol=tolerance, rtol=0.0)):
            break
        u_prev = u
        u = r / K.matmul(c / K.transpose(-2, -1).matmul(u))
    v = c / K.transpose(-2, -1).matmul(u)
    return (u * K) * v.transpose(-2, -1)


def gmw_edge_transport(kpts_2d, kpts_3d, sd, depth: int = NET_DEPTH):
    """GMW/model/model.py:170-192: (edge_P [b,E,E], reg_weights [b,E])."""
    f4 = edge_net(edge_expand(kpts_2d).transpose(-2, -1), sd, "FeatureExtractor4d", depth).transpose(-2, -1)
    f6 = edge_net(edge_expand(kpts_3d).transpose(-2, -1), sd, "FeatureExtractor6d", depth).transpose(-2, -1)
    f4 = F.normalize(f4, p=2, dim=-1)
    f6 = F.normalize(f6, p=2, dim=-1)
    M = pairwise_l2_dist(f4, f6)
    diag = 1. / M.diagonal(offset=0, dim1=-2, dim2=-1)
    b, m, n = M.size()
    r = M.new_ones((b, m)) / m
    c = M.new_ones((b, n)) / n
    return sinkhorn(M, r, c), diag


def correspondence_loss(P, C):
    """GMW/lib/losses.py:22-26,115-119."""
    return ((1.0 - 2.0 * C) * P).sum(dim=(-2, -1)).mean()


def compute_reg_loss(pre_depths, edge_weight, gt_depth, good_idx):
    """GMW/main.py:364-371."""
    z = pre_depths.gather(-1, good_idx)
    w = edge_weight.gather(-1, good_idx).softmax(dim=-1)
    Z_select_weighted = (z * w).sum(-1)
    reg_loss = (Z_select_weighted - gt_depth).abs().mean()
    return reg_loss, Z_select_weighted


def gmw_pipeline(kpts_2d, kpts_3d, pred_rot, sd, depth: int = NET_DEPTH, faithful: bool = False):
    """Whole GMW forward of GMW/main.py:524-533: compute_z -> GMW.forward -> weighted depth [b]."""
    with torch.no_grad():
        Z, idx = compute_z(kpts_2d, kpts_3d, pred_rot, faithful=faithful)
    w = gmw_reg_weights(kpts_2d, kpts_3d, sd, depth, full_matrix=faithful)
    z = Z.gather(-1, idx)
    p = w.gather(-1, idx).softmax(dim=-1)
    return (z * p).sum(-1)


def dgde_pipeline(kps, kps_3d, rot_y, K, faithful: bool = False):
    """DGDE inference: decode_pairs_kpts_depth(training=False) + mean (detector_infer.py:222-225)."""
    d, _ = decode_pairs_kpts_depth(kps, kps_3d, rot_y, K, training=False, faithful=faithful)
    return d.mean(1)


# ---------------------------------------------------------------------------------------------
# Frame epilogue around the DGDE edge solve: image-space keypoints from the
# regression channels, then the object's 3D location from the solved depth.
# ---------------------------------------------------------------------------------------------
DOWN_RATIO = 4      # DGDE/config/defaults.py (MODEL.BACKBONE.DOWN_RATIO); hard-coded `*4` at detector_infer.py:217


def decode_kpts_2d_img(kpts_off, points, offsets, pad_size, down_ratio=DOWN_RATIO):
    """DGDE/model/head/detector_infer.py:216-217 (`compute_pairs_kpts_depth`, `generate_infer_data`):
    real_pred_2d = (pred_extra_kpts_2d + (pred_bbox_points + pred_offset_3D)[:, None]) * 4 - pad_size.
    kpts_off [N,n,2], points/offsets [N,2], pad_size [2] or [N,2] -> [N,n,2]."""
    centre = (points + offsets).unsqueeze(1).expand_as(kpts_off)
    pad = pad_size if pad_size.dim() == 1 else pad_size.unsqueeze(1)
    return (kpts_off + centre) * down_ratio - pad


def calib_scalars(P):
    """DGDE/data/datasets/kitti_utils.py:239-244: intrinsics of a 3x4 projection matrix, computed in the
    matrix' own precision (the reference holds P as numpy float64)."""
    P = np.asarray(P)
    c_u, c_v, f_u, f_v = P[0, 2], P[1, 2], P[0, 0], P[1, 1]
    return c_u, c_v, f_u, f_v, P[0, 3] / (-f_u), P[1, 3] / (-f_v)


def project_image_to_rect(uv_depth, P):
    """kitti_utils.py:399-417 on a torch tensor [n,3] (u, v, depth) -> rect camera coordinates [n,3]."""
    c_u, c_v, f_u, f_v, b_x, b_y = (float(x) for x in calib_scalars(P))
    x = ((uv_depth[:, 0] - c_u) * uv_depth[:, 2]) / f_u + b_x
    y = ((uv_depth[:, 1] - c_v) * uv_depth[:, 2]) / f_v + b_y
    out = torch.zeros_like(uv_depth)
    out[:, 0] = x
    out[:, 1] = y
    out[:, 2] = uv_depth[:, 2]
    return out


def decode_location_flatten(points, offsets, depths, Ps, pad_size, batch_idxs, down_ratio=DOWN_RATIO):
    """DGDE/model/anno_encoder.py:147-161.  Ps: one 3x4 matrix per image of the batch, pad_size [B,2],
    batch_idxs [N] int64 -> locations [N,3]."""
    gts = torch.unique(batch_idxs, sorted=True).tolist()
    locations = points.new_zeros(points.shape[0], 3).float()
    pts = (points + offsets) * down_ratio - pad_size[batch_idxs]
    for gt in gts:
        sel = torch.nonzero(batch_idxs == gt).squeeze(-1)
        locations[sel] = project_image_to_rect(torch.cat((pts[sel], depths[sel, None]), dim=1), Ps[gt])
    return locations


def compute_pairs_kpts_depth(kpts_off, points, offsets, pad_size, kps_3d, rot_y, K, faithful=False):
    """detector_infer.py:215-227: image-space keypoints -> edge solve (inference form) -> mean over the edges."""
    real_2d = decode_kpts_2d_img(kpts_off, points, offsets, pad_size)
    d, _ = decode_pairs_kpts_depth(real_2d, kps_3d, rot_y, K, training=False, faithful=faithful)
    return d.mean(1)


def frame_locations(kpts_off, points, offsets, pad_size, kps_3d, rot_y, P, dims):
    """detector_infer.py:186-192 for one image: edge depths -> decode_location_flatten -> y += h / 2.
    P [3,4] (numpy float64 like the reference's calib.P), pad_size [1,2], dims [N,3] (l, h, w)."""
    N = kpts_off.shape[0]
    K = torch.from_numpy(np.asarray(P)).unsqueeze(0).expand(N, -1, -1)       # float64, stride 0 (detector_infer.py:221)
    depth = compute_pairs_kpts_depth(kpts_off, points, offsets, pad_size.reshape(-1)[:2], kps_3d, rot_y, K)
    batch_idxs = torch.zeros(N, dtype=torch.int64)
    loc = decode_location_flatten(points, offsets, depth, [P], pad_size.reshape(1, 2), batch_idxs)
    loc[:, 1] += dims[:, 1] / 2
    return depth, loc


def select_point_of_interest(batch, index, feature_maps):
    """DGDE/model/layers/utils.py:120-145: regression channels at the points of interest, [B,C,H,W] -> [B,K,C]."""
    w = feature_maps.shape[3]
    if index.dim() == 3:
        index = index[:, :, 1] * w + index[:, :, 0]
    index = index.view(batch, -1)
    fm = feature_maps.permute(0, 2, 3, 1).contiguous()
    channel = fm.shape[-1]
    fm = fm.view(batch, -1, channel)
    return fm.gather(1, index.unsqueeze(-1).repeat(1, 1, channel).long())


def decode_depth_from_keypoints_batch(pred_keypoints, pred_dimensions, f_us, batch_idxs=None, down_ratio=DOWN_RATIO,
                                      eps=1e-3, depth_range=(0.1, 100.0)):
    """DGDE/model/anno_encoder.py:193-224: depth from the projected heights of the 3D box (centre line and the two
    diagonal corner pairs).  pred_keypoints [N,10,2] (8 corners, bottom centre, top centre; feature-map units),
    pred_dimensions [N,3] (l, h, w), f_us: calib.f_u of every image -> [N,3] (centre, corner_02, corner_13)."""
    h3d = pred_dimensions[:, 1].clone()
    if len(f_us) == 1:
        batch_idxs = pred_dimensions.new_zeros(pred_dimensions.shape[0])
    center_height = pred_keypoints[:, -2, 1] - pred_keypoints[:, -1, 1]
    corner_02_height = pred_keypoints[:, [0, 2], 1] - pred_keypoints[:, [4, 6], 1]
    corner_13_height = pred_keypoints[:, [1, 3], 1] - pred_keypoints[:, [5, 7], 1]
    out = {"center": [], "corner_02": [], "corner_13": []}
    for idx, gt_idx in enumerate(torch.unique(batch_idxs, sorted=True).tolist()):
        f_u = float(f_us[idx])
        sel = torch.nonzero(batch_idxs == gt_idx).squeeze(-1)
        out["center"].append(f_u * h3d[sel] / (F.relu(center_height[sel]) * down_ratio + eps))
        c02 = f_u * h3d[sel].unsqueeze(-1) / (F.relu(corner_02_height[sel]) * down_ratio + eps)
        c13 = f_u * h3d[sel].unsqueeze(-1) / (F.relu(corner_13_height[sel]) * down_ratio + eps)
        out["corner_02"].append(c02.mean(dim=1))
        out["corner_13"].append(c13.mean(dim=1))
    return torch.stack([torch.clamp(torch.cat(v), min=depth_range[0], max=depth_range[1]) for v in out.values()], dim=1)


def depth_ensemble(direct_depths, keypoint_depths, direct_log_unc, keypoint_log_unc):
    """DGDE/model/head/detector_infer.py:141,152,156-171: uncertainties = exp(channels); inverse-uncertainty weighted
    soft ensemble of the direct depth and the three keypoint depths (or of the three alone when direct_depths is None).
    -> (pred_depths [N], estimated_depth_error [N], argmax of the weights [N])."""
    kp_unc = keypoint_log_unc.exp()
    if direct_depths is not None:
        depths = torch.cat((direct_depths.unsqueeze(1), keypoint_depths), dim=1)
        unc = torch.cat((direct_log_unc.reshape(-1, 1).exp(), kp_unc), dim=1)
    else:
        depths, unc = keypoint_depths.clone(), kp_unc.clone()
    w = 1 / unc
    amax = w.argmax(dim=1)
    w = w / w.sum(dim=1, keepdim=True)
    return torch.sum(depths * w, dim=1), torch.sum(w * unc, dim=1), amax


def uncertainty_scores(scores, estimated_depth_error):
    """detector_infer.py:197-203: scores * (1 - clamp(error, 0.01, 1)), NaN -> 0.   scores [N,1] -> [N,1]."""
    conf = 1 - torch.clamp(estimated_depth_error, min=0.01, max=1)
    out = scores * conf.view(-1, 1)
    out[torch.isnan(out)] = 0.0
    return out


def ray_rescale(raw_location, pred_depth, dim):
    """GMW/main.py:542-547: move the detector's location along its viewing ray (through the box centre) to the GMW depth."""
    raw_location = raw_location.clone()
    scale = pred_depth / raw_location[:, 2]
    h = dim[:, 0]
    raw_location[:, 1] -= h / 2
    pred_location = scale.unsqueeze(-1) * raw_location
    pred_location[:, 1] += h / 2
    return pred_location


def random_state_dict(seed: int, depth: int = NET_DEPTH, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded weights with torch's Conv1d default init bounds and the reference's key names
    (used on the GPU box, where the reference module cannot be instantiated)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cin):
        bound = 1.0 / math.sqrt(cin)
        sd[key + ".0.weight"] = ((torch.rand((NET_CH, cin, 1), generator=g) * 2 - 1) * bound).to(dtype)
        sd[key + ".0.bias"] = ((torch.rand((NET_CH,), generator=g) * 2 - 1) * bound).to(dtype)

    for name, cin in (("FeatureExtractor4d", 4), ("FeatureExtractor6d", 6)):
        conv(name + ".conv_in", cin)
        for k in range(depth):
            for sub in ("preconv", "conv1", "conv2"):
                conv("%s.conv_%d.%s" % (name, k, sub), NET_CH)
    return sd

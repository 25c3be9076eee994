"""GPU parity of the detector-head frame epilogue through the C ABI:
image-space keypoints -> edge solve -> mean depth -> 3D location, against the reference-generated fixtures
(`oracle/make_golden.py::locate_fixture`: unmodified decode_pairs_kpts_depth, decode_location_flatten and
Calibration.project_image_to_rect) and against the oracle on other shapes.

Bars: depth rel <= 1e-5 (the north-star's depth tolerance; CPU and GPU differ by the rounding of sin/cos and of the
mean's summation order); locations abs <= 1e-4 m + rel 1e-5 (x, y scale with depth; b_x, b_y are formed in FP32 here and
in float64 by the reference).  The fused depth equals `edge_depth_mean` on the materialised keypoints to 1e-6.
"""
import numpy as np
import pytest
import torch

import dcd_b200
from dcd_b200 import synth
from oracle import dcd_oracle as O
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(*ts):
    return [t.to(DEV) if torch.is_tensor(t) else t for t in ts]


@pytest.mark.parametrize("name", ["locate_n73_N50", "locate_n20_N7"])
def test_frame_epilogue_vs_reference_fixture(golden, name):
    G = golden(name)
    off, pts, ofs, pad, dims, k3, rot = cu(G["kpts_off"], G["points"], G["offsets"], G["pad"], G["dims"], G["kps_3d"], G["rot_y"])
    P = G["P"]
    depth, loc = dcd_b200.compute_pairs_kpts_depth(off, pts, ofs, pad, k3, rot, P, dims=dims, return_locations=True)
    assert depth.shape == (off.shape[0],) and loc.shape == (off.shape[0], 3)
    assert rel_err(depth.cpu(), G["depth"]) < 1e-5
    ref = G["locations"]
    assert bool(((loc.cpu() - ref).abs() <= 1e-4 + 1e-5 * ref.abs()).all())
    assert torch.equal(loc[:, 2], depth)
    # depth only, and the un-fused route through the materialised image-space keypoints
    d_only = dcd_b200.compute_pairs_kpts_depth(off, pts, ofs, pad, k3, rot, P)
    assert torch.equal(d_only, depth)
    real_2d = G["real_2d"].to(DEV)
    K = torch.as_tensor(P).unsqueeze(0).expand(off.shape[0], -1, -1).to(DEV)          # float64, stride 0 (detector_infer.py:221)
    unfused = dcd_b200.edge_depth_mean(real_2d, k3, rot, K)
    assert rel_err(depth, unfused) < 1e-6
    # decode_location_flatten on its own (the reference's first call, with the ensemble depth): same kernel, no solve
    bi = torch.zeros(off.shape[0], dtype=torch.int64, device=DEV)
    loc2 = dcd_b200.decode_location_flatten(pts, ofs, depth, P, pad, bi)
    loc2[:, 1] += dims[:, 1] / 2
    assert torch.allclose(loc2, loc, rtol=0, atol=1e-6)


def test_frame_epilogue_vs_oracle_batched_frames():
    """Several images in one call: per-object calibration and pad (pad_size[batch_idxs]), ragged object counts."""
    counts = torch.tensor([3, 50, 1, 17], dtype=torch.int64)
    ob = synth.make_objects(n=73, seed=77, counts=counts, jitter_fx=0.02)
    N = ob.N
    g = torch.Generator().manual_seed(5)
    pad_b = torch.tensor([[19.0, 5.0], [0.0, 0.0], [7.0, 3.0], [12.0, 8.0]])
    pad = pad_b[ob.frame_id]
    centre = (ob.kps.mean(1) + pad) / 4
    pts = centre.floor()
    ofs = centre - pts + 0.05 * (torch.rand(centre.shape, generator=g) - 0.5)
    off = (ob.kps + pad.unsqueeze(1)) / 4 - (pts + ofs).unsqueeze(1)
    dims = torch.stack((torch.full((N,), 3.9), -ob.kps_3d[:, -1, 1], torch.full((N,), 1.6)), dim=1)
    depth, loc = dcd_b200.compute_pairs_kpts_depth(*cu(off, pts, ofs), pad_b.to(DEV), *cu(ob.kps_3d, ob.rot_y), ob.K.to(DEV),
                                                   batch_idxs=ob.frame_id.to(DEV), dims=dims.to(DEV), return_locations=True)
    # oracle: per image, with that image's matrix in float64 like the reference's calib.P
    for f in range(len(counts)):
        sel = (ob.frame_id == f).nonzero().squeeze(-1)
        P = ob.K[sel[0]].double().numpy()
        d_o, loc_o = O.frame_locations(off[sel], pts[sel], ofs[sel], pad_b[f:f + 1], ob.kps_3d[sel], ob.rot_y[sel], P, dims[sel])
        assert rel_err(depth[sel.to(DEV)].cpu(), d_o) < 1e-5, f
        assert bool(((loc[sel.to(DEV)].cpu() - loc_o).abs() <= 1e-4 + 1e-5 * loc_o.abs()).all()), f
    # the solved depth tracks the generating depth, the location the generating translation
    assert float(((depth.cpu() - ob.gt_depth).abs() / ob.gt_depth).median()) < 0.05


def test_frame_epilogue_edge_cases():
    e = dcd_b200.compute_pairs_kpts_depth(torch.empty((0, 73, 2), device=DEV), torch.empty((0, 2), device=DEV),
                                          torch.empty((0, 2), device=DEV), torch.zeros(2, device=DEV),
                                          torch.empty((0, 73, 3), device=DEV), torch.empty((0, 1), device=DEV),
                                          np.array(synth.P2), return_locations=True)
    assert e[0].shape == (0,) and e[1].shape == (0, 3)
    with pytest.raises(RuntimeError):
        dcd_b200.compute_pairs_kpts_depth(torch.zeros((2, 73, 2)), torch.zeros((2, 2)), torch.zeros((2, 2)), torch.zeros(2),
                                          torch.zeros((2, 73, 3)), torch.zeros((2, 1)), np.array(synth.P2))     # CPU tensors


def test_depth_ensemble_and_ray_rescale_vs_reference_fixture(golden):
    """Rest of row N4 against outputs of the unmodified decode_depth_from_keypoints_batch (with the reference's
    Calibration) and of the ensemble / confidence / ray-rescale lines executed verbatim (oracle/make_golden.py).
    Tolerance 2e-6 relative: expf on the GPU and torch's CPU exp differ by an ulp."""
    G = golden("ensemble_N50")
    kp10, dims, direct, lud, luk, scores = cu(G["kp10"], G["dims"], G["direct"], G["log_unc_direct"], G["log_unc_kp"], G["scores"])
    P = G["P"]
    kd = dcd_b200.decode_depth_from_keypoints_batch(kp10, dims, P)
    assert torch.equal(kd.cpu(), G["keypoint_depths"])            # no transcendental here: bit-exact
    assert int(((kd <= 0.1) | (kd >= 100)).sum()) > 0              # the clamp and the relu + EPS path are exercised
    out = dcd_b200.depth_ensemble(kp10, dims, P, luk, direct_depths=direct, direct_log_uncertainty=lud, scores=scores)
    assert torch.equal(out["keypoint_depths"], kd)
    assert rel_err(out["depth"].cpu(), G["depth"]) < 2e-6
    assert rel_err(out["depth_error"].cpu(), G["depth_error"]) < 2e-6
    assert torch.equal(out["min_uncertainty"].cpu(), G["argmax"])
    assert rel_err(out["scores"].cpu(), G["scores_out"]) < 2e-6
    # keypoint depths only (no direct depth head): oracle
    o3 = dcd_b200.depth_ensemble(kp10, dims, P, luk)
    d3, e3, a3 = O.depth_ensemble(None, G["keypoint_depths"], None, G["log_unc_kp"])
    assert rel_err(o3["depth"].cpu(), d3) < 2e-6 and rel_err(o3["depth_error"].cpu(), e3) < 2e-6
    assert torch.equal(o3["min_uncertainty"].cpu(), a3) and o3["scores"] is None
    # NaN confidence -> score 0 (detector_infer.py:200-202)
    bad = luk.clone()
    bad[0, :] = float("nan")
    assert float(dcd_b200.depth_ensemble(kp10, dims, P, bad, scores=scores)["scores"][0]) == 0.0
    # GMW validation ray rescale
    loc = dcd_b200.ray_rescale(*cu(G["raw_location"], G["direct"], G["dim_hwl"]))
    assert torch.equal(loc.cpu(), G["pred_location"])
    assert torch.equal(loc[:, 2].cpu(), (G["direct"] / G["raw_location"][:, 2]) * G["raw_location"][:, 2])


def test_poi_gather_vs_reference_fixture(golden):
    """Row N2 gather: bit-exact against the unmodified select_point_of_interest, index and point forms, plus a
    detector-sized map (the reference would copy all of it to NHWC)."""
    G = golden("poi_gather")
    fm, idx = cu(G["feature_maps"], G["index"])
    out = dcd_b200.select_point_of_interest(fm.shape[0], idx, fm)
    assert torch.equal(out.cpu(), G["pois"])
    W = fm.shape[3]
    pts = torch.stack((idx % W, idx // W), dim=-1)
    assert torch.equal(dcd_b200.select_point_of_interest(fm.shape[0], pts, fm).cpu(), G["pois"])
    big = torch.randn((2, 440, 96, 320), device=DEV)
    bi = torch.randint(0, 96 * 320, (2, 50), device=DEV)
    assert torch.equal(dcd_b200.select_point_of_interest(2, bi, big), O.select_point_of_interest(2, bi, big))
    with pytest.raises(RuntimeError):
        dcd_b200.select_point_of_interest(2, bi + 96 * 320, big, validate=True)
    assert bool(torch.isnan(dcd_b200.select_point_of_interest(2, bi + 96 * 320, big)).all())     # no fault without it
    assert dcd_b200.select_point_of_interest(2, bi[:, :0], big).shape == (2, 0, 440)


def test_frame_from_regression_map_matches_gather_then_epilogue():
    """Row N2 fused into the load stage: reading the [B,C,H,W] map directly == select_point_of_interest + channel slices +
    the frame epilogue (bit-identical: same arithmetic on the same values), and the oracle's frame_locations."""
    B, C, H, W, n = 2, 415, 96, 320, 73                      # DGDE.yaml: 415 regression channels, 384 x 1280 input / 4
    g = torch.Generator().manual_seed(9)
    ob = synth.make_objects(n=n, seed=81, counts=torch.tensor([37, 50]))
    N = ob.N
    fmap = torch.randn(B, C, H, W, generator=g) * 0.1
    pos = torch.stack([torch.randperm(H * W, generator=g)[:N]])[0]
    bi = ob.frame_id.clone()
    pad_b = torch.tensor([[19.0, 5.0], [3.0, 7.0]])
    pad = pad_b[bi]
    px, py = (pos % W).float(), (pos // W).float()
    ofs = torch.rand(N, 2, generator=g) - 0.5
    # plant geometry-consistent values in the map: offsets such that (off + point + sub-pixel) * 4 - pad reproduces ob.kps
    centre = torch.stack((px, py), dim=1) + ofs
    off = (ob.kps + pad.unsqueeze(1)) / 4 - centre.unsqueeze(1)
    ch = dcd_b200.ops.DGDE_CHANNELS
    for d in range(N):
        b, p = int(bi[d]), int(pos[d])
        fmap[b, ch["3d_offset"]:ch["3d_offset"] + 2].view(2, -1)[:, p] = ofs[d]
        fmap[b, ch["extra_kpts_2d"]:ch["extra_kpts_2d"] + 2 * n].view(2 * n, -1)[:, p] = off[d].reshape(-1)
        fmap[b, ch["extra_kpts_3d"]:ch["extra_kpts_3d"] + 3 * n].view(3 * n, -1)[:, p] = ob.kps_3d[d].reshape(-1)
    dims = torch.stack((torch.full((N,), 3.9), -ob.kps_3d[:, -1, 1], torch.full((N,), 1.6)), dim=1)
    P = np.stack([np.array(synth.P2, dtype=np.float64)] * B)
    fm, posd, rot, dimsd = cu(fmap, pos, ob.rot_y, dims)
    with torch.no_grad():
        depth, loc, kimg, k3o = dcd_b200.frame_depths_from_map(fm, posd, rot, P, pad_b.to(DEV), dims=dimsd, batch_idxs=bi.to(DEV),
                                                               return_keypoints=True)
        # un-fused route through the gather (one image at a time, like the reference's batch-1 inference loop)
        for b in range(B):
            sel = (bi == b).nonzero().reshape(-1).to(DEV)
            pois = dcd_b200.select_point_of_interest(1, posd[sel].reshape(1, -1), fm[b:b + 1]).reshape(-1, C)
            o2 = pois[:, ch["extra_kpts_2d"]:ch["extra_kpts_2d"] + 2 * n].reshape(-1, n, 2)
            o3 = pois[:, ch["extra_kpts_3d"]:ch["extra_kpts_3d"] + 3 * n].reshape(-1, n, 3)
            of = pois[:, ch["3d_offset"]:ch["3d_offset"] + 2]
            pts = torch.stack((px, py), dim=1).to(DEV)[sel]
            d2, l2 = dcd_b200.compute_pairs_kpts_depth(o2, pts, of, pad_b[b].to(DEV), o3, rot[sel], P[b], dims=dimsd[sel],
                                                       return_locations=True)
            assert torch.equal(depth[sel], d2) and torch.equal(loc[sel], l2)
            assert torch.equal(k3o[sel], o3)
    assert float((kimg.cpu() - ob.kps).abs().max()) < 2e-3
    for b in range(B):
        sel = (bi == b).nonzero().reshape(-1)
        d_o, loc_o = O.frame_locations(off[sel], torch.stack((px, py), dim=1)[sel], ofs[sel], pad_b[b:b + 1], ob.kps_3d[sel],
                                       ob.rot_y[sel], P[b], dims[sel])
        assert rel_err(depth.cpu()[sel], d_o) < 1e-5
        assert bool(((loc.cpu()[sel] - loc_o).abs() <= 1e-4 + 1e-5 * loc_o.abs()).all())
    # a position outside the map: NaN, no fault
    bad = posd.clone()
    bad[0] = H * W + 5
    with torch.no_grad():
        d_bad, _ = dcd_b200.frame_depths_from_map(fm, bad, rot, P, pad_b.to(DEV), batch_idxs=bi.to(DEV))
    assert bool(torch.isnan(d_bad[0])) and torch.equal(d_bad[1:], depth[1:])

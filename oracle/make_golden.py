"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # from the repo root

For every fixture the script (1) runs the reference functions imported through
oracle/ref_loader.py on seeded synthetic inputs, (2) asserts that oracle/dcd_oracle.py
reproduces them (bit-exactly on CPU wherever the operation sequence is deterministic), and
(3) stores inputs + reference outputs.  The fixtures pin the oracle; the GPU parity tests
compare the CUDA path with the oracle and with these stored reference outputs.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dcd_oracle as O      # noqa: E402
from oracle import ref_loader as rl     # noqa: E402
from dcd_b200 import synth              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy()


def inputs_dict(ob):
    return dict(kps=npy(ob.kps), kps_norm=npy(ob.kps_norm), kps_3d=npy(ob.kps_3d), rot_y=npy(ob.rot_y),
                K=npy(ob.K), mask=npy(ob.mask), gt_depth=npy(ob.gt_depth))


def dgde_fixture(name, N, n, seed, with_grad=True):
    enc = rl.load_dgde_anno_encoder()
    ob = synth.make_objects(N=N, n=n, seed=seed)
    out = inputs_dict(ob)
    report = {}
    t0 = time.time()
    d_inf, none = enc.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, ob.K)
    report["t_ref_infer_s"] = time.time() - t0
    assert none is None
    d_or, _ = O.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, ob.K)
    assert torch.equal(d_inf, d_or), "oracle != reference (infer)"
    # inference call with the float64 stride-0 K of detector_infer.py:221
    K64 = ob.K[0].double().unsqueeze(0).expand(N, -1, -1) if n == 73 and N <= 50 else ob.K.double()
    d64, _ = enc.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, K64)
    out["infer_depth_mean"] = npy(d_inf.mean(1))
    out["infer_f64K_equal"] = np.array(bool(torch.equal(d64, d_inf)) if K64.shape == ob.K.shape and torch.equal(K64.float(), ob.K) else False)
    if n <= 73:
        out["infer_depth"] = npy(d_inf)
    # FP64 evaluation of the same formula (conditioning reference)
    d_f64, _ = O.decode_pairs_kpts_depth(ob.kps.double(), ob.kps_3d.double(), ob.rot_y.double(), ob.K.double())
    out["infer_depth_mean_f64"] = npy(d_f64.mean(1))
    E = n * (n - 1) // 2
    if E >= O.K_SEL:
        t0 = time.time()
        d_tr, m_tr = enc.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, ob.K, training=True, kpts_2d_mask=ob.mask)
        report["t_ref_train_s"] = time.time() - t0
        d_o, m_o, idx_t = O.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, ob.K, training=True,
                                                    kpts_2d_mask=ob.mask, canonical=False, return_idx=True)
        assert torch.equal(d_tr, d_o) and torch.equal(m_tr, m_o), "oracle != reference (train)"
        assert m_tr.dtype == torch.float32
        d_c, m_c, idx_c = O.decode_pairs_kpts_depth(ob.kps, ob.kps_3d, ob.rot_y, ob.K, training=True,
                                                    kpts_2d_mask=ob.mask, canonical=True, return_idx=True)
        absV = O.edge_abs_v(O.normalise_v(ob.kps, ob.K))
        assert O.topk_is_canonical_equivalent(absV, idx_t, O.K_SEL)
        out.update(train_depth_ref=npy(d_tr), train_mask_ref=npy(m_tr), train_idx_ref=npy(idx_t),
                   train_depth=npy(d_c), train_mask=npy(m_c), train_idx=npy(idx_c),
                   train_ties=np.array(int((idx_t != idx_c).sum())))
        same_set = all(set(a.tolist()) == set(b.tolist()) for a, b in zip(idx_t, idx_c))
        out["train_same_set"] = np.array(same_set)
        if with_grad:
            g = torch.Generator().manual_seed(seed + 99)
            G_edge = torch.randn(N, E, generator=g)
            kps = ob.kps.clone().requires_grad_(True)
            k3 = ob.kps_3d.clone().requires_grad_(True)
            t0 = time.time()
            d_g, m_g = enc.decode_pairs_kpts_depth(kps, k3, ob.rot_y, ob.K, training=True, kpts_2d_mask=ob.mask)
            (d_g * G_edge.gather(-1, idx_t)).sum().backward()
            report["t_ref_train_fwd_bwd_s"] = time.time() - t0
            kps_o = ob.kps.clone().requires_grad_(True)
            k3_o = ob.kps_3d.clone().requires_grad_(True)
            d_go, _, idx_go = O.decode_pairs_kpts_depth(kps_o, k3_o, ob.rot_y, ob.K, training=True,
                                                        kpts_2d_mask=ob.mask, canonical=True, return_idx=True)
            (d_go * G_edge.gather(-1, idx_go)).sum().backward()
            if same_set:
                for a, b in ((kps.grad, kps_o.grad), (k3.grad, k3_o.grad)):
                    assert (a - b).abs().max() <= 1e-4 * a.abs().max(), "oracle grad != reference grad"
            assert float(kps.grad[:, :, 0].abs().max()) == 0.0      # u-gradient is exactly 0
            out.update(G_edge=npy(G_edge), grad_kps=npy(kps.grad), grad_kps_3d=npy(k3.grad))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: round(v, 3) for k, v in report.items()}, "ties", out.get("train_ties"))


def gmw_fixture(name, N, seed, wseed):
    main, _ = rl.load_gmw()
    ob = synth.make_objects(N=N, n=73, seed=seed)
    sd = O.random_state_dict(wseed)
    model = rl.new_gmw_model(0)
    model.load_state_dict(sd, strict=True)
    model.train()
    out = inputs_dict(ob)
    out["weight_seed"] = np.array(wseed)
    with torch.no_grad():
        Z_ref, idx_ref = main.compute_z(ob.kps_norm, ob.kps_3d, ob.rot_y)
    Z_o, idx_t = O.compute_z(ob.kps_norm, ob.kps_3d, ob.rot_y, canonical=False)
    assert torch.equal(Z_ref, Z_o) and torch.equal(idx_ref, idx_t)
    _, idx_c = O.compute_z(ob.kps_norm, ob.kps_3d, ob.rot_y, canonical=True)
    absV = O.edge_abs_v(ob.kps_norm[:, :, 1])
    assert O.topk_is_canonical_equivalent(absV, idx_ref, O.K_SEL)
    # edge_expand parity
    model_ee4 = model.edge_expand(ob.kps_norm)
    assert torch.equal(model_ee4, O.edge_expand(ob.kps_norm))
    assert torch.equal(model.edge_expand(ob.kps_3d), O.edge_expand(ob.kps_3d))
    t0 = time.time()
    w_ref, _P = model(ob.kps_norm, ob.kps_3d, ob.rot_y, None)       # full forward incl. E x E + Sinkhorn
    t_fwd = time.time() - t0
    loss_ref, zsel_ref = main.compute_reg_loss(Z_ref, w_ref, ob.gt_depth, idx_c)
    model.zero_grad()
    t0 = time.time()
    loss_ref.backward()
    t_bwd = time.time() - t0
    names = [k for k, _ in model.named_parameters()]
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    # oracle parity (full-matrix form is bit-identical, diagonal form within FP32 noise)
    with torch.no_grad():
        w_full = O.gmw_reg_weights(ob.kps_norm, ob.kps_3d, sd, full_matrix=True)
        w_diag, f4, f6 = O.gmw_reg_weights(ob.kps_norm, ob.kps_3d, sd, return_feats=True)
    assert torch.equal(w_full, w_ref.detach()), "oracle(full) != reference GMW.forward"
    rel = ((w_diag - w_ref).abs() / w_ref.abs()).max().item()
    assert rel < 2e-5, rel
    _, zsel_o = O.compute_reg_loss(Z_ref, w_diag, ob.gt_depth, idx_c)
    assert ((zsel_o - zsel_ref).abs() / zsel_ref.abs()).max() < 1e-6
    # oracle gradient parity through the diagonal form
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    w_g = O.gmw_reg_weights(ob.kps_norm, ob.kps_3d, sd_g)
    l_g, _ = O.compute_reg_loss(Z_ref, w_g, ob.gt_depth, idx_c)
    l_g.backward()
    worst = 0.0
    for k in names:
        a, b = grads[k], sd_g[k].grad
        worst = max(worst, float((a - b).abs().max() / max(float(a.abs().max()), 1e-12)) if a.abs().max() > 1e-6 else 0.0)
    flat = torch.cat([grads[k].reshape(-1) for k in names])
    # FP64 evaluation (conditioning reference)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        w64 = O.gmw_reg_weights(ob.kps_norm.double(), ob.kps_3d.double(), sd64)
        _, zsel64 = O.compute_reg_loss(Z_ref.double(), w64, ob.gt_depth.double(), idx_c)
    out.update(Z=npy(Z_ref), idx_ref=npy(idx_ref), idx=npy(idx_c), reg_weights=npy(w_ref), reg_weights_f64=npy(w64),
               z_select_weighted=npy(zsel_ref), z_select_weighted_f64=npy(zsel64), reg_loss=npy(loss_ref),
               feat4_sample=npy(f4[:, ::97, :]), feat6_sample=npy(f6[:, ::97, :]),
               grad_names=np.array(names), grad_sample=npy(flat[::61]),
               grad_norms=np.array([float(grads[k].norm()) for k in names], dtype=np.float64),
               grad_absmax=np.array([float(grads[k].abs().max()) for k in names], dtype=np.float64),
               grad_conv_in4_w=npy(grads["FeatureExtractor4d.conv_in.0.weight"]),
               grad_conv_in6_w=npy(grads["FeatureExtractor6d.conv_in.0.weight"]),
               grad_last4_w=npy(grads["FeatureExtractor4d.conv_11.conv2.0.weight"]),
               grad_first6_w=npy(grads["FeatureExtractor6d.conv_0.preconv.0.weight"]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ref fwd %.2fs bwd %.2fs" % (t_fwd, t_bwd), "diag-vs-full rel", rel, "oracle grad worst rel", worst)


def frame_inputs(ob, seed, pad=(19.0, 5.0)):
    """Detector-head form of synthetic objects (one image): heat-map peak `points` (integer feature-map pixel),
    sub-pixel `offsets`, per-keypoint offsets such that (kpts_off + points + offsets) * 4 - pad reproduces ob.kps,
    and the (l, h, w) dimensions the template was drawn with."""
    g = torch.Generator().manual_seed(seed)
    pad_t = torch.tensor([pad], dtype=torch.float32)
    centre = (ob.kps.mean(1) + pad_t) / 4                      # somewhere inside the object, feature-map units
    points = centre.floor()
    offsets = centre - points + 0.1 * (torch.rand(centre.shape, generator=g) - 0.5)
    kpts_off = (ob.kps + pad_t) / 4 - (points + offsets).unsqueeze(1)
    h = -ob.kps_3d[:, -1, 1]                                   # top centre of the template sits at y = -h
    dims = torch.stack((ob.kps_3d[:, -10, 0].abs() * 2, h, ob.kps_3d[:, -10, 2].abs() * 2), dim=1)
    return kpts_off.contiguous(), points.contiguous(), offsets.contiguous(), pad_t, dims.contiguous()


def load_reference_calibration(P):
    """The reference's own Calibration (DGDE/data/datasets/kitti_utils.py:206-244) on a temporary KITTI calib file."""
    import importlib.util
    import tempfile
    rl._stub_matplotlib()
    spec = importlib.util.spec_from_file_location(
        "ref_kitti_utils", os.path.join(rl.REFERENCE_ROOT, "DGDE", "data", "datasets", "kitti_utils.py"))
    ku = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ku)
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        for key in ("P2", "P3"):
            f.write(key + ": " + " ".join(repr(float(x)) for x in np.asarray(P).reshape(-1)) + "\n")
        f.write("R0_rect: 1 0 0 0 1 0 0 0 1\n")
        f.write("Tr_velo_to_cam: 1 0 0 0 0 1 0 0 0 0 1 0\n")
        path = f.name
    calib = ku.Calibration(path)
    os.unlink(path)
    return calib


def locate_fixture(name, N, n, seed):
    """Frame epilogue: detector_infer.py:215-227 + :186-192 with the reference's own
    decode_pairs_kpts_depth, decode_location_flatten and Calibration.project_image_to_rect."""
    enc = rl.load_dgde_anno_encoder()
    enc.down_ratio = O.DOWN_RATIO
    ob = synth.make_objects(N=N, n=n, seed=seed)
    P = np.array(synth.P2, dtype=np.float64)
    calib = load_reference_calibration(P)
    kpts_off, points, offsets, pad, dims = frame_inputs(ob, seed)
    # --- reference, line by line
    real_2d = (kpts_off + (points + offsets).unsqueeze(1).expand_as(kpts_off)) * 4 - pad          # :216-217
    Calib_P = torch.from_numpy(calib.P).unsqueeze(0).expand(N, -1, -1)                               # :221
    pairs, _ = enc.decode_pairs_kpts_depth(real_2d, ob.kps_3d, ob.rot_y, Calib_P)                    # :222
    depth = pairs.mean(1)                                                                            # :225
    bi = depth.new_zeros(N).long()                                                                   # :186
    loc = enc.decode_location_flatten(points, offsets, depth, [calib], pad, bi)                      # :187
    loc[:, 1] += dims[:, 1] / 2                                                                      # :188
    # --- oracle restatement
    assert torch.equal(O.decode_kpts_2d_img(kpts_off, points, offsets, pad.reshape(-1)), real_2d)
    d_o, loc_o = O.frame_locations(kpts_off, points, offsets, pad, ob.kps_3d, ob.rot_y, P, dims)
    assert torch.equal(d_o, depth) and torch.equal(loc_o, loc), "oracle != reference (frame epilogue)"
    assert float((real_2d - ob.kps).abs().max()) < 1e-3
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kpts_off=npy(kpts_off), points=npy(points), offsets=npy(offsets),
                        pad=npy(pad), dims=npy(dims), kps_3d=npy(ob.kps_3d), rot_y=npy(ob.rot_y), P=P,
                        real_2d=npy(real_2d), depth=npy(depth), locations=npy(loc), gt_depth=npy(ob.gt_depth))
    print(name, "frame epilogue: depth err vs gt (median rel) %.4f" % float(((depth - ob.gt_depth).abs() / ob.gt_depth).median()))


def ensemble_fixture(name, N, seed):
    """Rest of row N4: decode_depth_from_keypoints_batch (anno_encoder.py:193-224, the unmodified method with the
    reference's Calibration), the 4-depth uncertainty ensemble and the confidence (detector_infer.py:141-171,197-203,
    executed line by line), and the GMW-val ray rescale (GMW/main.py:542-547)."""
    enc = rl.load_dgde_anno_encoder()
    enc.down_ratio, enc.EPS, enc.depth_range = O.DOWN_RATIO, 1e-3, [0.1, 100]
    ob = synth.make_objects(N=N, n=73, seed=seed)
    P = np.array(synth.P2, dtype=np.float64)
    calib = load_reference_calibration(P)
    kpts_off, points, offsets, pad, dims = frame_inputs(ob, seed)
    g = torch.Generator().manual_seed(seed + 1)
    kp10 = kpts_off[:, -10:, :].contiguous()                  # the box corners + bottom/top centres, feature-map units
    kp10[::7, 1, 1] = kp10[::7, 5, 1] - 0.3                   # some inverted pairs: exercises relu + EPS and the clamp
    direct = (ob.gt_depth * (1 + 0.05 * torch.randn(N, generator=g))).contiguous()
    lu_d = (-1.5 + 0.5 * torch.randn((N, 1), generator=g)).contiguous()
    lu_k = (-1.0 + 0.7 * torch.randn((N, 3), generator=g)).contiguous()
    scores = torch.rand((N, 1), generator=g)
    # --- reference
    kd = enc.decode_depth_from_keypoints_batch(kp10, dims, [calib])                                  # :150
    pred_direct_uncertainty = lu_d.exp()                                                             # :141
    pred_keypoint_uncertainty = lu_k.exp()                                                           # :154
    pred_combined_depths = torch.cat((direct.unsqueeze(1), kd), dim=1)                               # :159
    pred_combined_uncertainty = torch.cat((pred_direct_uncertainty, pred_keypoint_uncertainty), dim=1)
    depth_weights = 1 / pred_combined_uncertainty                                                    # :165
    amax = depth_weights.argmax(dim=1)
    depth_weights = depth_weights / depth_weights.sum(dim=1, keepdim=True)                           # :168
    pred_depths = torch.sum(pred_combined_depths * depth_weights, dim=1)
    estimated_depth_error = torch.sum(depth_weights * pred_combined_uncertainty, dim=1)              # :171
    uncertainty_conf = 1 - torch.clamp(estimated_depth_error, min=0.01, max=1)                       # :198
    sc = scores * uncertainty_conf.view(-1, 1)
    sc[torch.isnan(sc)] = 0.0
    # GMW validation: rescale a detector location to the GMW depth (main.py:542-547; dim = (h, w, l) there)
    raw_location = torch.stack((0.3 * ob.gt_depth, torch.full((N,), 1.6), ob.gt_depth), dim=1)
    dim_hwl = dims.roll(shifts=-1, dims=1)
    pred_depth = direct
    rl_in = raw_location.clone()
    raw_depth = rl_in[:, 2]
    scale = pred_depth / raw_depth
    h = dim_hwl[:, 0]
    rl_in[:, 1] -= h / 2
    pred_location = scale.unsqueeze(-1) * rl_in
    pred_location[:, 1] += h / 2
    # --- oracle
    kd_o = O.decode_depth_from_keypoints_batch(kp10, dims, [calib.f_u])
    assert torch.equal(kd_o, kd), "oracle != reference (decode_depth_from_keypoints_batch)"
    d_o, e_o, a_o = O.depth_ensemble(direct, kd, lu_d, lu_k)
    assert torch.equal(d_o, pred_depths) and torch.equal(e_o, estimated_depth_error) and torch.equal(a_o, amax)
    assert torch.equal(O.uncertainty_scores(scores, e_o), sc)
    assert torch.equal(O.ray_rescale(raw_location, pred_depth, dim_hwl), pred_location)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kp10=npy(kp10), dims=npy(dims), P=P, direct=npy(direct),
                        log_unc_direct=npy(lu_d), log_unc_kp=npy(lu_k), scores=npy(scores), keypoint_depths=npy(kd),
                        depth=npy(pred_depths), depth_error=npy(estimated_depth_error), argmax=npy(amax), scores_out=npy(sc),
                        raw_location=npy(raw_location), dim_hwl=npy(dim_hwl), pred_location=npy(pred_location))
    print(name, "ensemble: clamped keypoint depths", int(((kd <= 0.1) | (kd >= 100)).sum()), "of", kd.numel())


def poi_fixture(name, seed):
    """Row N2 gather: the unmodified select_point_of_interest (DGDE/model/layers/utils.py:120-145)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_layers_utils", os.path.join(rl.REFERENCE_ROOT, "DGDE", "model", "layers", "utils.py"))
    lu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lu)
    g = torch.Generator().manual_seed(seed)
    B, C, H, W, K = 2, 37, 24, 80, 50
    fm = torch.randn((B, C, H, W), generator=g)
    idx = torch.randint(0, H * W, (B, K), generator=g)
    ref = lu.select_point_of_interest(B, idx, fm)
    pts = torch.stack((idx % W, idx // W), dim=-1)
    assert torch.equal(lu.select_point_of_interest(B, pts, fm), ref)
    assert torch.equal(O.select_point_of_interest(B, idx, fm), ref) and torch.equal(O.select_point_of_interest(B, pts, fm), ref)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), feature_maps=npy(fm), index=npy(idx), pois=npy(ref))
    print(name, "poi gather", tuple(ref.shape))


def wire_fixture(seed):
    """Row N3: small gen_data_train.json / gen_data_infer.json in the reference writers' layout (detector_loss.py:148-173 +
    trainer.py:208-215; inference.py:59-84, json.dump(..., indent=4)) and what the unmodified reader
    (GMW/utilities/dataset_utilities.py:11-56) makes of them."""
    import importlib.util
    import json
    import types
    spec = importlib.util.spec_from_file_location(
        "ref_dataset_utilities", os.path.join(rl.REFERENCE_ROOT, "GMW", "utilities", "dataset_utilities.py"))
    du = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(du)
    # training file: per iteration a list over the batch's objects (two iterations of 3 and 2 objects)
    train = {k: [] for k in ("kpts_2d", "kpts_3d", "img_idx", "pred_rot", "gt_location", "pred_location")}
    for it, nobj in enumerate((3, 2)):
        ob = synth.make_objects(N=nobj, n=73, seed=seed + it)
        train["kpts_2d"].append(ob.kps_norm.numpy().tolist())
        train["kpts_3d"].append(ob.kps_3d.numpy().tolist())
        train["img_idx"].append(["%06d" % (10 * it + j) for j in range(nobj)])
        train["pred_rot"].append(ob.rot_y.reshape(-1).numpy().tolist())
        loc = torch.stack((0.3 * ob.gt_depth, torch.full((nobj,), 1.6), ob.gt_depth), dim=1)
        train["gt_location"].append(loc.numpy().tolist())
        train["pred_location"].append((loc * 1.01).numpy().tolist())
    # inference file: image id -> list of detections (83 keypoints regressed, GMW keeps the first 73)
    infer = {}
    for img, nobj in (("000007", 2), ("000123", 3)):
        ob = synth.make_objects(N=nobj, n=83, seed=seed + int(img))
        infer[img] = []
        for j in range(nobj):
            infer[img].append({"kpts_2d": ob.kps_norm[j].numpy().tolist(), "kpts_3d": ob.kps_3d[j].numpy().tolist(),
                               "pred_rot": ob.rot_y[j].numpy().tolist(), "box": [10.0, 20.0, 110.0, 90.0],
                               "dim": [1.5, 1.6, 3.9], "pred_location": [0.3 * float(ob.gt_depth[j]), 1.6, float(ob.gt_depth[j])],
                               "score": [0.9], "cat": "Car"})
    tp, ip = os.path.join(OUT, "gen_data_train_small.json"), os.path.join(OUT, "gen_data_infer_small.json")
    json.dump(train, open(tp, "w"), indent=4)
    json.dump(infer, open(ip, "w"), indent=4)
    args = types.SimpleNamespace(train_data_path=tp, val_data_path=ip)
    ref_t = du.load_data(args, "train")
    ref_v = du.load_data(args, "valid")
    from dcd_b200 import wire
    for ref, split, path in ((ref_t, "train", tp), (ref_v, "valid", ip)):
        mine = wire.load_reference_json(path, split)
        for k, v in ref.items():
            assert mine[k].dtype == v.dtype and mine[k].shape == v.shape and np.array_equal(mine[k], v), (split, k)
    np.savez_compressed(os.path.join(OUT, "wire_reference_load_data.npz"),
                        **{"train_" + k: v for k, v in ref_t.items()}, **{"valid_" + k: v for k, v in ref_v.items()})
    print("wire fixture:", {k: v.shape for k, v in ref_v.items()}, os.path.getsize(tp), os.path.getsize(ip))


def transport_fixture(name, N, seed, wseed):
    """Row N1 forward: the unmodified GMW.forward (E x E pairwiseL2Dist + RegularisedTransport Sinkhorn, GMW/model/model.py:170-207,
    GMW/lib/optimal_transport.py:52-72) on N objects.  edge_P is 27.6 MB per object, so the fixture keeps what the loops use
    (sum P, trace P, correspondenceLoss against eye: main.py:456-457) plus its diagonal, marginals and three full rows."""
    main, _ = rl.load_gmw()
    ob = synth.make_objects(N=N, n=73, seed=seed)
    sd = O.random_state_dict(wseed)
    model = rl.new_gmw_model(0)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        t0 = time.time()
        w_ref, P_ref = model(ob.kps_norm, ob.kps_3d, ob.rot_y, None)
        t_ref = time.time() - t0
        P_o, diag_o = O.gmw_edge_transport(ob.kps_norm, ob.kps_3d, sd)
    assert torch.equal(P_o, P_ref) and torch.equal(diag_o, w_ref), "oracle != reference (edge transport)"
    eye = torch.eye(P_ref.shape[1]).expand_as(P_ref)
    from lib.losses import correspondenceLoss
    cls_ref = correspondenceLoss(P_ref, eye)
    assert torch.equal(O.correspondence_loss(P_ref, eye), cls_ref)
    rows = [0, 1313, 2627]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kps_norm=npy(ob.kps_norm), kps_3d=npy(ob.kps_3d),
                        weight_seed=np.array(wseed), reg_weights=npy(w_ref), P_sum=npy(P_ref.sum((-2, -1))),
                        P_trace=npy(P_ref.diagonal(dim1=-2, dim2=-1).sum(-1)), P_diag=npy(P_ref.diagonal(dim1=-2, dim2=-1)),
                        P_rowsum=npy(P_ref.sum(-1)), P_colsum=npy(P_ref.sum(-2)), P_rows=npy(P_ref[:, rows, :]),
                        rows=np.array(rows), cls_loss=npy(cls_ref))
    print(name, "reference forward incl. Sinkhorn %.1f s, cls_loss %.6f, sum P" % (t_ref, float(cls_ref)), P_ref.sum((-2, -1)).tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    dgde_fixture("dgde_n73_N10", N=10, n=73, seed=synth.BASE_SEED)          # BASELINE configs[0]
    dgde_fixture("dgde_n73_N64", N=64, n=73, seed=synth.BASE_SEED + 10, with_grad=False)
    dgde_fixture("dgde_n60_N5", N=5, n=60, seed=synth.BASE_SEED + 11)
    dgde_fixture("dgde_n8_N7", N=7, n=8, seed=synth.BASE_SEED + 12)         # E=28 < 1500: inference only
    dgde_fixture("dgde_n256_N4", N=4, n=256, seed=synth.BASE_SEED + 4)      # BASELINE configs[4]
    gmw_fixture("gmw_n73_N4", N=4, seed=synth.BASE_SEED + 3, wseed=7)
    locate_fixture("locate_n73_N50", N=50, n=73, seed=synth.BASE_SEED + 20)   # one full frame
    locate_fixture("locate_n20_N7", N=7, n=20, seed=synth.BASE_SEED + 21)
    ensemble_fixture("ensemble_N50", N=50, seed=synth.BASE_SEED + 22)
    poi_fixture("poi_gather", seed=synth.BASE_SEED + 23)
    wire_fixture(seed=synth.BASE_SEED + 24)
    transport_fixture("transport_n73_N2", N=2, seed=synth.BASE_SEED + 25, wseed=7)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "round2"):
    main()


def transport_bwd_fixture(name, E, N, seed, sharp=0.0, store_full=True):
    """Row N1 backward at the feature level: the unmodified pairwiseL2Dist (GMW/model/model.py:17-36) + RegularisedTransport
    (GMW/lib/optimal_transport.py, forward :52-72 and the implicit backward :75-128,184-222) on given edge features
    [N,E,128]; gradients of  sum(V * P)  w.r.t. the L2-normalised features, V = the correspondence-loss weights
    (1 - 2 eye) / N of GMW/main.py:456-457 for object 0.. and a random V for the last object.  FP32 (the reference as it
    runs) and FP64 (the same code in double: the conditioning anchor).  sharp > 0 makes f6 ~ f4 + sharp * noise: a
    trained-like, nearly diagonal plan."""
    rl.load_gmw()
    from model.model import pairwiseL2Dist
    from lib.optimal_transport import RegularisedTransport
    g = torch.Generator().manual_seed(seed)
    f4 = torch.randn(N, E, 128, generator=g) * (0.5 + torch.rand(N, E, 1, generator=g))
    f6 = (f4 + sharp * torch.randn(N, E, 128, generator=g)) if sharp > 0 else torch.randn(N, E, 128, generator=g) * 1.3
    V = ((1.0 - 2.0 * torch.eye(E)) / N).expand(N, E, E).clone()
    V[-1] = torch.randn(E, E, generator=g) / N
    res = {}
    for dt in (torch.float32, torch.float64):
        a = torch.nn.functional.normalize(f4.to(dt), p=2, dim=-1).requires_grad_(True)
        c = torch.nn.functional.normalize(f6.to(dt), p=2, dim=-1).requires_grad_(True)
        M = pairwiseL2Dist(a, c)
        r = M.new_ones((N, E)) / E
        cc = M.new_ones((N, E)) / E
        P = RegularisedTransport(10.0, 1e-9)(M, r, cc)
        (P * V.to(dt)).sum().backward()
        res[dt] = (P.detach(), a.grad.detach(), c.grad.detach())
    P32, ga32, gc32 = res[torch.float32]
    P64, ga64, gc64 = res[torch.float64]
    out = dict(E=np.array(E), N=np.array(N), seed=np.array(seed), sharp=np.array(sharp),
               P_sum=npy(P32.sum((-2, -1))), P_trace=npy(P32.diagonal(dim1=-2, dim2=-1).sum(-1)),
               P_trace_f64=npy(P64.diagonal(dim1=-2, dim2=-1).sum(-1)))
    if store_full:
        out.update(feat4=npy(f4), feat6=npy(f6), V=npy(V), grad_a=npy(ga32), grad_c=npy(gc32), grad_a_f64=npy(ga64), grad_c_f64=npy(gc64))
    else:       # inputs are regenerated from the seed by the test (same torch.Generator stream); gradients sampled
        out.update(grad_a=npy(ga32[:, ::16, :]), grad_c=npy(gc32[:, ::16, :]), grad_a_f64=npy(ga64[:, ::16, :]),
                   grad_c_f64=npy(gc64[:, ::16, :]), V_last_sample=npy(V[-1, ::97, ::97]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    e32 = max(float((ga32.double() - ga64).abs().max() / ga64.abs().max()), float((gc32.double() - gc64).abs().max() / gc64.abs().max()))
    print(name, "trace P", P64.diagonal(dim1=-2, dim2=-1).sum(-1).tolist(), "reference FP32 vs FP64 gradient (rel max)", e32)


def transport_inputs(E, N, seed, sharp):
    """The inputs of transport_bwd_fixture(store_full=False), regenerated from the seed."""
    g = torch.Generator().manual_seed(seed)
    f4 = torch.randn(N, E, 128, generator=g) * (0.5 + torch.rand(N, E, 1, generator=g))
    f6 = (f4 + sharp * torch.randn(N, E, 128, generator=g)) if sharp > 0 else torch.randn(N, E, 128, generator=g) * 1.3
    V = ((1.0 - 2.0 * torch.eye(E)) / N).expand(N, E, E).clone()
    V[-1] = torch.randn(E, E, generator=g) / N
    return f4, f6, V


def gmw_train_fixture(name, N, seed, wseed, cls_weight, reg_weight):
    """One training step's loss and gradients through the UNMODIFIED reference (GMW/main.py:453-465): compute_z, GMW.forward
    (both outputs), correspondenceLoss(edge_P, eye), compute_reg_loss, loss = cls_weight * cls + reg_weight * reg, backward
    (incl. RegularisedTransportFn.backward) — in FP32 as the reference runs and in FP64 (same modules in double) as the
    conditioning anchor.  Stores per-tensor norms, max-norms and a 1/61 sample of every gradient in both precisions."""
    main, _ = rl.load_gmw()
    from lib.losses import correspondenceLoss
    ob = synth.make_objects(N=N, n=73, seed=seed)
    sd = O.random_state_dict(wseed)
    res = {}
    for dt in (torch.float32, torch.float64):
        model = rl.new_gmw_model(0)
        model.load_state_dict(sd, strict=True)
        model = model.to(dt).train()
        k2, k3, rot, gt = ob.kps_norm.to(dt), ob.kps_3d.to(dt), ob.rot_y.to(dt), ob.gt_depth.to(dt)
        with torch.no_grad():
            Z, _ = main.compute_z(k2, k3, rot)
        _, idx_c = O.compute_z(ob.kps_norm, ob.kps_3d, ob.rot_y, canonical=True)       # FP32 keys: same selection in both runs
        w, P = model(k2, k3, rot, None)
        eye = torch.eye(P.shape[1], dtype=dt).expand_as(P)
        cls = correspondenceLoss(P, eye)
        reg, zsel = main.compute_reg_loss(Z, w, gt, idx_c)
        loss = cls_weight * cls + reg_weight * reg
        model.zero_grad()
        loss.backward()
        res[dt] = (float(cls), float(reg), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    names = list(res[torch.float32][2].keys())
    g32, g64 = res[torch.float32][2], res[torch.float64][2]
    out = inputs_dict(ob)
    out.update(weight_seed=np.array(wseed), cls_weight=np.array(cls_weight), reg_weight=np.array(reg_weight),
               cls_loss=np.array(res[torch.float32][0]), reg_loss=np.array(res[torch.float32][1]),
               cls_loss_f64=np.array(res[torch.float64][0]), reg_loss_f64=np.array(res[torch.float64][1]),
               grad_names=np.array(names),
               grad_sample=npy(torch.cat([g32[k].reshape(-1) for k in names])[::61]),
               grad_sample_f64=npy(torch.cat([g64[k].reshape(-1) for k in names])[::61]),
               grad_absmax_f64=np.array([float(g64[k].abs().max()) for k in names]),
               grad_ref_err=np.array([float((g32[k].double() - g64[k]).abs().max()) for k in names]),
               grad_conv_in4_w_f64=npy(g64["FeatureExtractor4d.conv_in.0.weight"]),
               grad_last6_w_f64=npy(g64["FeatureExtractor6d.conv_11.conv2.0.weight"]),
               grad_conv_in4_w=npy(g32["FeatureExtractor4d.conv_in.0.weight"]),
               grad_last6_w=npy(g32["FeatureExtractor6d.conv_11.conv2.0.weight"]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    worst = max(float((g32[k].double() - g64[k]).abs().max() / max(float(g64[k].abs().max()), 1e-30)) for k in names
                if float(g64[k].abs().max()) > 1e-6)
    print(name, "cls %.6f reg %.6f; reference FP32 vs FP64 weight gradients, worst tensor (rel max-norm) %.3g" % (
        res[torch.float32][0], res[torch.float32][1], worst))


def main_round2():
    """Fixtures added in round 2 (the Sinkhorn backward and the full training loss)."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    transport_bwd_fixture("transport_bwd_E190_N3", E=190, N=3, seed=synth.BASE_SEED + 30)
    transport_bwd_fixture("transport_bwd_E190_sharp", E=190, N=2, seed=synth.BASE_SEED + 31, sharp=0.05)
    transport_bwd_fixture("transport_bwd_E2628_N2", E=2628, N=2, seed=synth.BASE_SEED + 32, store_full=False)
    transport_bwd_fixture("transport_bwd_E2628_sharp", E=2628, N=1, seed=synth.BASE_SEED + 33, sharp=0.03, store_full=False)
    gmw_train_fixture("gmw_train_n73_N2", N=2, seed=synth.BASE_SEED + 34, wseed=7, cls_weight=1.0, reg_weight=0.0)
    gmw_train_fixture("gmw_train_n73_N2_reg", N=2, seed=synth.BASE_SEED + 35, wseed=7, cls_weight=0.1, reg_weight=1.0)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "round2":
    main_round2()

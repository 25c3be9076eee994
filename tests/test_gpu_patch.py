"""The GMW drop-in end to end on the GPU: dcd_b200.patch.install on a module shaped like the reference GMW (same parameter
tree: FeatureExtractor{4,6}d.conv_in.0 / conv_<k>.{preconv,conv1,conv2}.0 — the reference tree itself cannot travel to the
GPU box) and the reference's training-loop lines GMW/main.py:453-466 run VERBATIM around it, optimizer included.

What this pins: the patched forward reads the LIVE parameters (no install-time snapshot), gradients arrive in the
reference parameters' .grad (zero_grad / step of the reference optimizer just work, nothing accumulates), `edge_P` is a real
differentiable tensor (main.py:456-457 does not crash), a load_state_dict after install is honoured.
Oracle: the same two optimizer steps by autograd through the FP64 oracle.
"""
import types

import pytest
import torch
import torch.nn as nn

import dcd_b200
from dcd_b200 import patch, synth
from oracle import dcd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def reference_shaped_gmw(sd, depth):
    """nn.Module with the reference GMW's parameter names and shapes (GMW/model/model.py:103-112, yi2018cvpr/model.py:21-47)."""
    def conv(cin):
        return nn.Sequential(nn.Conv1d(cin, 128, 1))

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.preconv, self.conv1, self.conv2 = conv(128), conv(128), conv(128)

    class Extractor(nn.Module):
        def __init__(self, cin):
            super().__init__()
            self.conv_in = conv(cin)
            for k in range(depth):
                setattr(self, "conv_%d" % k, Block())

    class RefGMW(nn.Module):
        def __init__(self):
            super().__init__()
            self.FeatureExtractor4d, self.FeatureExtractor6d = Extractor(4), Extractor(6)

        def forward(self, kpts_2d, kpts_3d, pred_rot, args):
            raise AssertionError("the reference forward must not run once patched")

    m = RefGMW()
    m.load_state_dict(sd, strict=True)
    return m


def oracle_step_f64(sd64, k2, k3, rot, gt, cls_weight, reg_weight, with_cls):
    """loss and gradients of GMW/main.py:453-461 by autograd through the FP64 oracle (Sinkhorn unrolled)."""
    depth = sum(1 for k in sd64 if k.startswith("FeatureExtractor4d.conv_") and k.endswith("preconv.0.weight"))
    sd = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
    with torch.no_grad():
        Z, _ = O.compute_z(k2.double(), k3.double(), rot.double())
        _, idx = O.compute_z(k2, k3, rot)                                   # FP32 keys decide the selection
    if with_cls:
        P, w = O.gmw_edge_transport(k2.double(), k3.double(), sd, depth)
        eye = torch.eye(P.shape[1], dtype=P.dtype, device=P.device).expand_as(P)
        cls = O.correspondence_loss(P, eye)
    else:
        w = O.gmw_reg_weights(k2.double(), k3.double(), sd, depth)
        cls = torch.zeros((), dtype=torch.float64, device=k2.device)
    reg, _ = O.compute_reg_loss(Z, w, gt.double(), idx)
    loss = cls_weight * cls + reg_weight * reg
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("with_cls,cls_weight,reg_weight", [(True, 1.0, 0.0), (True, 0.1, 1.0), (False, 0.0, 1.0)])
def test_reference_training_loop_runs_unmodified_through_the_patch(with_cls, cls_weight, reg_weight):
    depth, b, lr = 2, 2, 0.05
    sd = O.random_state_dict(21, depth=depth)
    model = reference_shaped_gmw(sd, depth).to(DEV)
    main = types.ModuleType("main")                                  # stands for the imported GMW/main.py
    main.compute_z = lambda *a: (_ for _ in ()).throw(AssertionError("reference compute_z must be replaced"))
    main.compute_reg_loss = main.compute_z
    args = types.SimpleNamespace(cls_weight=cls_weight, reg_weight=reg_weight)
    ob = synth.make_objects(N=b, n=73, seed=77)
    kpts_2d, kpts_3d, pred_rot = ob.kps_norm.to(DEV), ob.kps_3d.to(DEV), ob.rot_y.to(DEV)
    gt_location = torch.stack((torch.zeros(b), torch.zeros(b), ob.gt_depth), dim=1).to(DEV)
    correspondenceLoss = O.correspondence_loss
    optimizer = torch.optim.SGD(model.parameters(), lr=lr)
    patch.install(gmw_main=main, gmw_model=model, with_edge_P=with_cls)
    try:
        compute_z, compute_reg_loss = main.compute_z, main.compute_reg_loss
        sd64 = {k: v.detach().double().to(DEV) for k, v in model.state_dict().items()}
        for step in range(2):
            want_loss, want_grads = oracle_step_f64(sd64, kpts_2d, kpts_3d, pred_rot, gt_location[:, -1], cls_weight, reg_weight, with_cls)
            # ---- GMW/main.py:453-466, verbatim (edge_P handling only when the cls branch is on)
            pre_depths, good_idx = compute_z(kpts_2d, kpts_3d, pred_rot)
            reg_weights, edge_P = model(kpts_2d, kpts_3d, pred_rot, args)
            if with_cls:
                edge_P_gt = torch.eye(edge_P.shape[1]).expand_as(edge_P).to(edge_P.device)
                cls_loss = correspondenceLoss(edge_P, edge_P_gt)
            else:
                assert edge_P is None
                cls_loss = 0.0
            reg_loss, pred_depth = compute_reg_loss(pre_depths, reg_weights, gt_location[:, -1], good_idx)
            loss = args.cls_weight * cls_loss + args.reg_weight * reg_loss
            optimizer.zero_grad()
            if not torch.isnan(loss).any():
                loss.backward()
            # ---- gradients are in the reference parameters' .grad, equal to the FP64 oracle's
            assert abs(float(loss) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
            worst = 0.0
            for k, p in model.named_parameters():
                assert p.grad is not None, k
                g64 = want_grads[k]
                amax = float(g64.abs().max())
                if amax < 1e-7 * max(1.0, abs(want_loss)):           # dead block biases (cancelled by the context norm)
                    assert float(p.grad.abs().max()) < 1e-6
                    continue
                e = float((p.grad.double() - g64).abs().max()) / amax
                worst = max(worst, e)
                assert e < 5e-2, (step, k, e)         # depth-2 nets in FP32: the worst tensor measures 6e-3 / 1.6e-2 (step 0 / 1) on B200
            optimizer.step()
            # the oracle takes the same SGD step in FP64
            sd64 = {k: v - lr * want_grads[k] for k, v in sd64.items()}
            print("step", step, "loss", float(loss), "worst gradient tensor rel err vs FP64 oracle %.3g" % worst)
        # after two steps the live parameters ARE the stepped ones (no frozen snapshot): compare the parameter UPDATE
        sd0 = {k: v.to(DEV) for k, v in sd.items()}
        for k, p in model.state_dict().items():
            upd, want = p.double() - sd0[k].double(), sd64[k] - sd0[k].double()
            if float(want.abs().max()) > 1e-9:     # (+ the FP32 resolution of the parameter itself: an update can be below its ulp)
                assert float((upd - want).abs().max()) <= 5e-2 * float(want.abs().max()) + 2.4e-7 * float(p.abs().max()), k
        # a checkpoint loaded after install (GMW/main.py:275-297 --resume) is what the next forward uses
        with torch.no_grad():
            w_before, _ = model(kpts_2d, kpts_3d, pred_rot, args)
            model.load_state_dict({k: v.to(DEV) for k, v in O.random_state_dict(22, depth=depth).items()})
            w_after, _ = model(kpts_2d, kpts_3d, pred_rot, args)
            fresh = dcd_b200.GMW(depth=depth).to(DEV).load_reference_state_dict(O.random_state_dict(22, depth=depth))
            w_fresh, _ = fresh(kpts_2d, kpts_3d, pred_rot, None)
        assert not torch.equal(w_before, w_after) and torch.equal(w_after, w_fresh)
    finally:
        patch.uninstall()
    with pytest.raises(AssertionError):
        model(kpts_2d, kpts_3d, pred_rot, args)                       # the original forward is back


def test_validation_loop_lines_through_the_patch():
    """GMW/main.py:524-547 (validation): compute_z, forward (edge_P consumed), reg loss, ray rescale."""
    depth, b = 2, 3
    sd = O.random_state_dict(23, depth=depth)
    model = reference_shaped_gmw(sd, depth).to(DEV).eval()
    main = types.ModuleType("main")
    main.compute_z = main.compute_reg_loss = None
    ob = synth.make_objects(N=b, n=73, seed=78)
    kpts_2d, kpts_3d, pred_rot = ob.kps_norm.to(DEV), ob.kps_3d.to(DEV), ob.rot_y.to(DEV)
    raw_location = torch.stack((ob.gt_depth * 0.1, torch.full((b,), 1.6), ob.gt_depth * 1.05), dim=1).to(DEV)
    patch.install(gmw_main=main, gmw_model=model)
    try:
        with torch.no_grad():
            pre_depths, good_idx = main.compute_z(kpts_2d, kpts_3d, pred_rot)
            reg_weights, edge_P = model(kpts_2d, kpts_3d, pred_rot, None)
            edge_P_gt = torch.eye(edge_P.shape[1]).expand_as(edge_P).to(edge_P.device)
            cls_loss = O.correspondence_loss(edge_P, edge_P_gt)
            reg_loss, pred_depth = main.compute_reg_loss(pre_depths, reg_weights, raw_location[:, -1], good_idx)
            P_o, w_o = O.gmw_edge_transport(kpts_2d, kpts_3d, {k: v.to(DEV) for k, v in sd.items()}, depth)
        assert float((cls_loss - O.correspondence_loss(P_o, edge_P_gt)).abs()) < 1e-5
        ref = O.gmw_pipeline(kpts_2d, kpts_3d, pred_rot, {k: v.to(DEV) for k, v in sd.items()}, depth)
        assert float(((pred_depth - ref).abs() / ref.abs()).max()) < 1e-5
    finally:
        patch.uninstall()

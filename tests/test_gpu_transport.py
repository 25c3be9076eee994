"""GPU parity of the correspondence branch, forward AND backward, through the C ABI.

Oracle = the unmodified reference: pairwiseL2Dist (GMW/model/model.py:17-36), RegularisedTransport (GMW/lib/
optimal_transport.py: Sinkhorn :52-72, implicit backward :75-128,184-222), correspondenceLoss (lib/losses.py:115-119) and the
training-step lines of GMW/main.py:453-465, run in the build container by oracle/make_golden.py (`round2`) in FP32 (as the
reference runs) and in FP64 (the same code in double).

Conditioning finding pinned here: the reference's FP32 backward factorises S = diag(colsum B) - B'^T diag(1/rowsum) B' whose
condition number is >= E; at E = 2628 its FP32 gradients differ from the FP64 evaluation of the same code by 16 % (random
features) to > 100 % (trained-like, nearly diagonal plans).  A bar "rel <= 1e-4 vs the reference's FP32 values" is therefore
meaningless for this row; the kernels are held to the FP64 evaluation instead:  |ours - f64| <= max(2 |ref32 - f64|, tol * max|f64|)
with a tolerance the FP32 forward (P itself carries ~1e-6 of rounding) supports.
"""
import pytest
import torch

import dcd_b200
from dcd_b200 import _lib, synth
from dcd_b200._lib import check, ptr, stream_ptr
from oracle import dcd_oracle as O
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
LAMBDA, TOL, ITERS = 10.0, 1e-9, 100


def n_of_edges(E):
    n = int(round((1 + (1 + 8 * E) ** 0.5) / 2))
    assert n * (n - 1) // 2 == E
    return n


def transport_fwd_bwd(f4, f6, V, cg_iters=32, cg_tol=1e-7):
    """f4, f6 [N,E,128] un-normalised edge features (CPU), V [N,E,E] = dL/dP -> P, dL/d(normalised f4), dL/d(normalised f6), cg info."""
    L = _lib.lib()
    N, E, _ = f4.shape
    n = n_of_edges(E)
    a = f4.transpose(1, 2).contiguous().to(DEV)          # [N,128,E] channel-major, as dcd_gmw_weights_fwd emits them
    c = f6.transpose(1, 2).contiguous().to(DEV)
    Vd = V.contiguous().to(DEV)
    P = torch.empty((N, E, E), device=DEV)
    u, v = torch.empty((N, E), device=DEV), torch.empty((N, E), device=DEV)
    sums = torch.empty((N, 2), device=DEV)
    ws = torch.full((L.dcd_gmw_transport_workspace_bytes(N, n) // 4 + 64,), float("nan"), device=DEV)
    check(L.dcd_gmw_transport_fwd(ptr(a), ptr(c), N, n, LAMBDA, TOL, ITERS, ptr(P), ptr(u), ptr(v), ptr(sums), ptr(ws),
                                  ws.numel() * 4, stream_ptr()), "transport_fwd")
    ga, gc = torch.empty_like(a), torch.empty_like(c)
    info = torch.empty((N, 8), device=DEV)
    ws2 = torch.full((L.dcd_gmw_transport_bwd_workspace_bytes(N, n) // 4 + 64,), float("nan"), device=DEV)
    check(L.dcd_gmw_transport_bwd(ptr(a), ptr(c), ptr(P), ptr(u), ptr(v), ptr(Vd), N, n, LAMBDA, cg_iters, cg_tol, ptr(ga), ptr(gc),
                                  ptr(info), ptr(ws2), ws2.numel() * 4, stream_ptr()), "transport_bwd")
    torch.cuda.synchronize()
    return P.cpu(), ga.transpose(1, 2).cpu(), gc.transpose(1, 2).cpu(), info.cpu(), sums.cpu()


def anchored_ok(ours, f64, ref32, tol):
    """|ours - f64| <= max(2 |ref32 - f64|, tol * max|f64|) in max-norm; returns (ok, our error, reference error), relative."""
    scale = float(f64.abs().max())
    e_ours = float((ours.double() - f64).abs().max()) / scale
    e_ref = float((ref32.double() - f64).abs().max()) / scale
    return e_ours <= max(2 * e_ref, tol), e_ours, e_ref


@pytest.mark.parametrize("name,tol", [("transport_bwd_E190_N3", 2e-4), ("transport_bwd_E190_sharp", 2e-2)])
def test_transport_backward_small_vs_reference(golden, name, tol):
    G = golden(name)
    P, ga, gc, info, sums = transport_fwd_bwd(G["feat4"], G["feat6"], G["V"])
    assert rel_err(sums[:, 0], G["P_sum"]) < 1e-5
    assert rel_err(sums[:, 1], G["P_trace_f64"].float()) < 2e-3
    assert bool((info[:, 2] == 1).all()), "conjugate gradients did not converge: %s" % info
    assert float(info[:, 3].max()) <= 32
    for ours, f64, r32, what in ((ga, G["grad_a_f64"], G["grad_a"], "d/da"), (gc, G["grad_c_f64"], G["grad_c"], "d/dc")):
        ok, e_o, e_r = anchored_ok(ours, f64, r32, tol)
        print(name, what, "ours vs f64 %.3g, reference FP32 vs f64 %.3g, cg iterations %s" % (e_o, e_r, info[:, 3].tolist()))
        assert ok, (what, e_o, e_r)
        assert e_o <= tol                                   # and an absolute bar against the FP64 evaluation


@pytest.mark.parametrize("name,E,N,sharp,tol", [("transport_bwd_E2628_N2", 2628, 2, 0.0, 2e-3),
                                                ("transport_bwd_E2628_sharp", 2628, 1, 0.03, 5e-2)])
def test_transport_backward_full_size_vs_reference(golden, name, E, N, sharp, tol):
    """n = 73 (E = 2628): inputs regenerated from the fixture's seed, gradients compared on the stored 1/16 sample."""
    from oracle.make_golden import transport_inputs
    G = golden(name)
    f4, f6, V = transport_inputs(E, N, int(G["seed"]), sharp)
    assert torch.equal(V[-1, ::97, ::97], G["V_last_sample"])           # same generator stream as the fixture
    P, ga, gc, info, sums = transport_fwd_bwd(f4, f6, V)
    assert rel_err(sums[:, 0], G["P_sum"]) < 1e-5
    assert rel_err(sums[:, 1], G["P_trace_f64"].float()) < 5e-3
    assert bool((info[:, 2] == 1).all()), info
    for ours, f64, r32, what in ((ga[:, ::16], G["grad_a_f64"], G["grad_a"], "d/da"), (gc[:, ::16], G["grad_c_f64"], G["grad_c"], "d/dc")):
        ok, e_o, e_r = anchored_ok(ours, f64, r32, tol)
        print(name, what, "ours vs f64 %.3g, reference FP32 vs f64 %.3g, cg iterations %s" % (e_o, e_r, info[:, 3].tolist()))
        assert ok and e_o <= tol, (what, e_o, e_r)


def test_transport_backward_matches_autograd_through_fp64_sinkhorn():
    """Independent of the fixtures: autograd through the oracle's UNROLLED Sinkhorn iterations in FP64 (not the implicit
    formula) on a converged problem gives the same gradient."""
    E, N = 190, 2
    g = torch.Generator().manual_seed(5)
    f4, f6 = torch.randn(N, E, 128, generator=g), torch.randn(N, E, 128, generator=g)
    V = torch.randn(N, E, E, generator=g)
    a = torch.nn.functional.normalize(f4.double(), dim=-1).requires_grad_(True)
    c = torch.nn.functional.normalize(f6.double(), dim=-1).requires_grad_(True)
    M = O.pairwise_l2_dist(a, c)
    r = M.new_ones((N, E)) / E
    P64 = O.sinkhorn(M, r, r, max_iterations=100)
    (P64 * V.double()).sum().backward()
    P, ga, gc, info, _ = transport_fwd_bwd(f4, f6, V)
    assert float((P.double() - P64.detach()).abs().max() / P64.abs().max()) < 1e-4
    for ours, ref in ((ga, a.grad), (gc, c.grad)):
        assert float((ours.double() - ref).abs().max() / ref.abs().max()) < 2e-4


@pytest.mark.parametrize("name", ["gmw_train_n73_N2", "gmw_train_n73_N2_reg"])
def test_training_step_gradients_vs_fp64_reference(golden, name):
    """GMW/main.py:453-465 through the drop-in module: loss = cls_weight * correspondenceLoss(edge_P, eye) + reg_weight * reg.
    All 148 gradient tensors (1/61 sample each) against the FP64 run of the unmodified reference, FP64-anchored:
    |ours - f64| <= max(2 |ref32 - f64|, 5e-2 max|f64|) per tensor (the MLP's own FP32 conditioning)."""
    G = golden(name)
    sd = O.random_state_dict(int(G["weight_seed"]))
    model = dcd_b200.GMW().to(DEV).load_reference_state_dict(sd)
    model.with_edge_P = True
    k2, k3, rot, gt = (G[k].to(DEV) for k in ("kps_norm", "kps_3d", "rot_y", "gt_depth"))
    Z, idx = dcd_b200.compute_z(k2, k3, rot)
    w, P = model(k2, k3, rot, None)
    assert P.shape == (2, 2628, 2628) and P.requires_grad and w.requires_grad
    eye = torch.eye(P.shape[1], device=DEV).expand_as(P)                      # main.py:456
    cls = ((1.0 - 2.0 * eye) * P).sum(dim=(-2, -1)).mean()                    # correspondenceLoss, lib/losses.py:22-26,115-119
    reg, zsel = dcd_b200.compute_reg_loss(Z, w, gt, idx)
    loss = float(G["cls_weight"]) * cls + float(G["reg_weight"]) * reg
    loss.backward()
    assert abs(float(cls) - float(G["cls_loss_f64"])) < 1e-5
    assert abs(float(reg) - float(G["reg_loss_f64"])) <= 1e-5 * max(1.0, abs(float(G["reg_loss_f64"])))
    grads = model.reference_grads()
    names = [str(x) for x in G["grad_names"]]
    flat = torch.cat([grads[k].reshape(-1) for k in names]).cpu()[::61]
    f64, r32 = G["grad_sample_f64"], G["grad_sample"]
    sizes = [grads[k].numel() for k in names]
    pos, worst, worst_ref = 0, 0.0, 0.0
    bounds = []
    for k, sz in zip(names, sizes):
        bounds.append((pos, pos + sz))
        pos += sz
    for (lo, hi), k, amax, ref_err in zip(bounds, names, G["grad_absmax_f64"].tolist(), G["grad_ref_err"].tolist()):
        s_lo, s_hi = (lo + 60) // 61, (hi + 60) // 61               # sample indices t with lo <= 61 t < hi
        if s_hi <= s_lo or amax < 1e-6:                              # dead block biases: ~0 on both sides
            continue
        e = float((flat[s_lo:s_hi].double() - f64[s_lo:s_hi]).abs().max())
        worst = max(worst, e / amax)
        worst_ref = max(worst_ref, ref_err / amax)
        assert e <= max(2 * ref_err, 5e-2 * amax), (k, e / amax, ref_err / amax)
    print(name, "worst tensor: ours vs f64 %.3g (rel max-norm), reference FP32 vs f64 %.3g" % (worst, worst_ref))
    # two fully stored tensors
    for key, full in (("FeatureExtractor4d.conv_in.0.weight", "grad_conv_in4_w"), ("FeatureExtractor6d.conv_11.conv2.0.weight", "grad_last6_w")):
        f = G[full + "_f64"]
        e_o = float((grads[key].cpu().double() - f).abs().max() / f.abs().max())
        e_r = float((G[full].double() - f).abs().max() / f.abs().max())
        assert e_o <= max(2 * e_r, 5e-2), (key, e_o, e_r)


def test_edge_p_matches_edge_transport_and_no_grad_path():
    ob = synth.make_objects(N=2, n=73, seed=17)
    model = dcd_b200.GMW().to(DEV).load_reference_state_dict(O.random_state_dict(3))
    k2, k3, rot = ob.kps_norm.to(DEV), ob.kps_3d.to(DEV), ob.rot_y.to(DEV)
    model.with_edge_P = True
    with torch.no_grad():
        w0, P0 = model(k2, k3, rot, None)
        w1, P1, sums = model.edge_transport(k2, k3)
    assert not P0.requires_grad and torch.equal(P0, P1) and torch.equal(w0, w1)
    assert rel_err(P0.sum((-2, -1)), sums[:, 0]) < 1e-5
    w2, P2 = model(k2, k3, rot, None)                            # training forward (layer-wise kernels): same plan to FP32 noise
    assert P2.requires_grad
    assert float((P2 - P0).abs().max() / P0.abs().max()) < 1e-3
    model.with_edge_P = False
    w3, none = model(k2, k3, rot, None)
    assert none is None and rel_err(w3, w2) < 1e-6

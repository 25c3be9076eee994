#!/usr/bin/env python
"""Profiling driver: one full GMW training step through the drop-in module (GMW/main.py:453-465) — compute_z, forward with
both outputs (reg_weights, edge_P), correspondence loss + reg loss, backward incl. the Sinkhorn backward.

    python profiles/run_train.py [N] [n] [cls_weight]

Prints CUDA-event times of forward, backward; used under ncu to capture mlp_tc / mlp_bwd_tc / tb_* kernels in isolation.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200  # noqa: E402
from dcd_b200 import synth  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 73
    cls_w = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
    ob = synth.make_objects(N=N, n=n, seed=5)
    model = dcd_b200.GMW(depth=12).cuda().load_reference_state_dict(synth.random_state_dict(7))
    model.with_edge_P = cls_w != 0.0
    k2, k3, rot, gt = ob.kps_norm.cuda(), ob.kps_3d.cuda(), ob.rot_y.cuda(), ob.gt_depth.cuda()
    for rep in range(3):
        model.zero_grad(set_to_none=True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        Z, idx = dcd_b200.compute_z(k2, k3, rot)
        w, P = model(k2, k3, rot, None)
        reg, _ = dcd_b200.compute_reg_loss(Z, w, gt, idx)
        loss = reg
        if P is not None:
            eye = torch.eye(P.shape[1], device=P.device).expand_as(P)
            loss = cls_w * ((1.0 - 2.0 * eye) * P).sum(dim=(-2, -1)).mean() + reg
        e[1].record()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        print("N=%d n=%d cls_weight=%g: forward %.3f ms, backward %.3f ms, loss %.6f" % (N, n, cls_w, e[0].elapsed_time(e[1]),
                                                                                       e[1].elapsed_time(e[2]), float(loss)))


if __name__ == "__main__":
    main()

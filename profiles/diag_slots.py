#!/usr/bin/env python
"""Diagnostic: are the fused forward's edge weights bit-identical whichever place of a triple / schedule an object takes?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200
from dcd_b200 import synth
depth = int(os.environ.get("DEPTH", "12"))
ob = synth.make_objects(N=36, n=73, seed=5)
model = dcd_b200.GMW(depth=depth).cuda().load_reference_state_dict(synth.random_state_dict(7, depth=depth) if depth != 12 else synth.random_state_dict(7))
k2, k3 = ob.kps_norm.cuda(), ob.kps_3d.cuda()
with torch.no_grad():
    wp, _ = model(k2, k3)
    for step in (7, 6, 5):
        wu = torch.cat([model(k2[i:i + step].contiguous(), k3[i:i + step].contiguous())[0] for i in range(0, 36, step)])
        print("chunk", step, "equal", torch.equal(wp, wu))
        for i in range(36):
            d = (wp[i] != wu[i])
            if d.any():
                idx = d.nonzero().flatten()
                rel = ((wp[i] - wu[i]).abs() / wp[i].abs()).max().item()
                print(" obj %2d place paired %d unpaired %d: %4d edges differ, first %s last %d, max rel %.2e" % (i, i % 3, (i % step) % 3, idx.numel(), idx[:4].tolist(), idx[-1].item(), rel))

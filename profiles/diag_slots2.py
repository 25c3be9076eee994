#!/usr/bin/env python
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200
from dcd_b200 import synth
ob = synth.kitti_val_batch(ragged=True, frames=6)
model = dcd_b200.GMW(depth=12).cuda().load_reference_state_dict(synth.random_state_dict(9))
k2, k3, rot = ob.kps_norm.cuda(), ob.kps_3d.cuda(), ob.rot_y.cuda()
N = k2.shape[0]
print("N", N)
with torch.no_grad():
    a = dcd_b200.gmw_weighted_depth(k2, k3, rot, model, chunk=1024)
    b = dcd_b200.gmw_weighted_depth(k2, k3, rot, model, chunk=17)
    print("depth equal", torch.equal(a, b), (a != b).nonzero().flatten().tolist())
    wp, _ = model(k2, k3)
    wp2, _ = model(k2, k3)
    print("deterministic", torch.equal(wp, wp2))
    wu = torch.cat([model(k2[i:i + 17].contiguous(), k3[i:i + 17].contiguous())[0] for i in range(0, N, 17)])
    print("weights equal", torch.equal(wp, wu))
    for i in range(N):
        d = (wp[i] != wu[i])
        if d.any():
            idx = d.nonzero().flatten()
            rel = ((wp[i] - wu[i]).abs() / wp[i].abs()).max().item()
            print(" obj %3d place paired %d unpaired %d: %4d edges differ, first %s last %d, max rel %.2e" % (i, i % 3, (i % 17) % 3, idx.numel(), idx[:4].tolist(), idx[-1].item(), rel))

#!/usr/bin/env python
"""Launch the DGDE-side kernels a few times on the BASELINE configs[1] batch (for ncu captures; not a benchmark).

    ncu --set full --clock-control none --import-source on -k regex:'edge_mean_group|edge_select_radix|edge_solve_bwd' -c 6 \
        -o gpurun_out/prof_solve python profiles/run_solve.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcd_b200 import _lib, synth  # noqa: E402
from dcd_b200._lib import check, ptr, stream_ptr  # noqa: E402

n = int(os.environ.get("KPTS", "73"))
frames = int(os.environ.get("FRAMES", "3769"))
ob = synth.kitti_val_batch(ragged=True, frames=frames, n=n)
dev = torch.device("cuda")
kps, k3, rot, K = ob.kps.to(dev), ob.kps_3d.to(dev), ob.rot_y.reshape(-1).contiguous().to(dev), ob.K.to(dev)
N = ob.N
L = _lib.lib()
mean = torch.empty(N, device=dev)
sel_n = min(N, 16384)
idx = torch.empty((sel_n, 1500), dtype=torch.int64, device=dev)
z = torch.empty((sel_n, 1500), device=dev)
g = torch.randn((sel_n, 1500), device=dev)
gk, g3 = torch.empty((sel_n, n, 2), device=dev), torch.empty((sel_n, n, 3), device=dev)
for rep in range(2):
    for flags in (3, 7):
        check(L.dcd_edge_solve_fwd(ptr(kps), ptr(k3), ptr(rot), ptr(K), N, n, 2.0, 80.0, flags, 0, ptr(mean), stream_ptr()), "solve")
    check(L.dcd_edge_select_fwd(ptr(kps), ptr(k3), ptr(rot), ptr(K), 0, sel_n, n, 1500, 2.0, 80.0, 3, ptr(idx), ptr(z), 0, 0, stream_ptr()), "select")
    check(L.dcd_edge_solve_bwd(ptr(kps), ptr(k3), ptr(rot), ptr(K), sel_n, n, 2.0, 80.0, 3, ptr(idx), 1500, ptr(g), 0, ptr(gk), ptr(g3),
                               stream_ptr()), "bwd")
torch.cuda.synchronize()
print("ok", N, float(mean.mean()))

#!/usr/bin/env python
"""A/B timing of the blocked edge-mean kernel for the group sizes selectable with DCD_B200_BLOCK_G (tuning aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcd_b200 import _lib, synth
from dcd_b200._lib import check, ptr, stream_ptr
ob = synth.kitti_val_batch(ragged=True)
dev = torch.device("cuda")
kps, k3, rot, K = ob.kps.to(dev), ob.kps_3d.to(dev), ob.rot_y.reshape(-1).contiguous().to(dev), ob.K.to(dev)
N = ob.N
L = _lib.lib()
mean = torch.empty(N, device=dev)
def t(flags, reps=20):
    for _ in range(5):
        check(L.dcd_edge_solve_fwd(ptr(kps), ptr(k3), ptr(rot), ptr(K), N, 73, 2.0, 80.0, flags, 0, ptr(mean), stream_ptr()), "s")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        check(L.dcd_edge_solve_fwd(ptr(kps), ptr(k3), ptr(rot), ptr(K), N, 73, 2.0, 80.0, flags, 0, ptr(mean), stream_ptr()), "s")
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for G in ("3", "5", "7"):
    os.environ["DCD_B200_BLOCK_G"] = G
    print("G", G, "exact %.4f ms" % t(3), "fast %.4f ms" % t(7), flush=True)

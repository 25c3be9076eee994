#!/usr/bin/env python
"""Profiling driver: correspondence-branch forward (row N1) and the detector-head frame epilogue (rows N2/N4).

    python profiles/run_transport.py [N]

Prints CUDA-event times; used under ncu to capture the kernels of these rows in isolation.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200  # noqa: E402
from dcd_b200 import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ob = synth.make_objects(N=N, n=73, seed=5)
    model = dcd_b200.GMW(depth=12).cuda().load_reference_state_dict(synth.random_state_dict(7))
    k2, k3 = ob.kps_norm.cuda(), ob.kps_3d.cuda()
    ms = timed(lambda: model.edge_transport(k2, k3, materialise=False))
    print("edge_transport N=%d: %.3f ms (%.0f objects/s)" % (N, ms, N / ms * 1e3))
    big = synth.make_objects(N=50000, n=73, seed=6)
    pad = torch.tensor([[19.0, 5.0]])
    ctr = (big.kps.mean(1) + pad) / 4
    pts, ofs = ctr.floor(), ctr - ctr.floor()
    off = (big.kps + pad) / 4 - ctr.unsqueeze(1)
    dims = torch.stack((torch.full((big.N,), 3.9), -big.kps_3d[:, -1, 1], torch.full((big.N,), 1.6)), dim=1)
    args = [t.cuda() for t in (off, pts, ofs, pad, big.kps_3d, big.rot_y)]
    P = np.array(synth.P2, dtype=np.float64)
    dd = dims.cuda()
    ms = timed(lambda: dcd_b200.compute_pairs_kpts_depth(*args, P, dims=dd, return_locations=True))
    print("frame epilogue N=%d: %.3f ms (%.1f M objects/s)" % (big.N, ms, big.N / ms / 1e3))


if __name__ == "__main__":
    main()

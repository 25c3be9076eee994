#!/usr/bin/env python
"""Profiling driver: the inference forward of the GMW edge-weight network on N synthetic objects.

    python profiles/run_gmw_infer.py [N] [reps]

Prints the CUDA-event time of the whole dcd_gmw_weights_fwd call (ms, objects/s).  Used under ncu to capture
the kernels of this path in isolation (see B200_PROFILING.md for the ncu command lines).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200  # noqa: E402
from dcd_b200 import _lib, synth  # noqa: E402

if os.environ.get("DCD_B200_LIB"):            # experiment builds of the library
    _lib._lib = _lib.load_library(os.environ["DCD_B200_LIB"])


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ob = synth.make_objects(N=N, n=73, seed=5)
    model = dcd_b200.GMW(depth=12).cuda().load_reference_state_dict(synth.random_state_dict(7))
    k2, k3 = ob.kps_norm.cuda(), ob.kps_3d.cuda()
    with torch.no_grad():
        model(k2, k3)
        torch.cuda.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            w, _ = model(k2, k3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("N=%d  %.3f ms  %.0f objects/s  finite=%s" % (N, ms, N / ms * 1e3, bool(torch.isfinite(w).all())))


if __name__ == "__main__":
    main()

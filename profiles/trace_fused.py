#!/usr/bin/env python
"""In-kernel timeline of the fused GMW forward (debug aid, not a benchmark).

Build the library with the trace hooks, run this on a GPU, then rebuild without them:

    python -m dcd_b200.build --trace            # -> dcd_b200/libdcd_b200_trace.so, the product library is untouched
    gpurun -- 'DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_trace.so python profiles/trace_fused.py > gpurun_out/trace.txt'

Prints "slot tag delta_cycles absolute" (slot 0 = converter warp 0, slot 1 = MMA warp, slot 2 = first statistics warp of
CTA 0).  Converter tags: 1000*kind + {100 step start, 200 operand buffer free, 300 accumulators loaded, 400 operand stored,
500 arrived} + 10*half + sub-tile; 603 partial posted, 606/608 wait for / got a layer's statistics, 600/601 final features.
MMA warp: 100.. wait for operand, 200.. operand ready, 10/11 weights ready, 20/21/22 next weights: prefetched / layer's MMAs
complete / published.  Statistics warp: 700+half wait for the partials, 710 got them, 720 published, 730 posted.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcd_b200  # noqa: E402
from dcd_b200 import _lib, synth  # noqa: E402

ob = synth.make_objects(N=int(os.environ.get('TRACE_OBJECTS', '8')), n=73, seed=5)
model = dcd_b200.GMW(depth=12).cuda().load_reference_state_dict(synth.random_state_dict(7))
with torch.no_grad():
    model(ob.kps_norm.cuda(), ob.kps_3d.cuda())
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 12288)()
n = (ctypes.c_int * 3)()
L.dcd_debug_fused_trace(buf, n)
t0 = min(buf[slot * 4096 + 1] for slot in range(3) if n[slot] > 0)
for slot in range(3):
    prev = None
    for i in range(min(n[slot], 2048)):
        tag, t = buf[slot * 4096 + 2 * i], buf[slot * 4096 + 2 * i + 1]
        print(slot, tag, t - (prev if prev is not None else t), t - t0)       # slot, tag, cycles since the slot's previous event, absolute
        prev = t

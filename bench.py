#!/usr/bin/env python
"""Benchmark of the densely-constrained-depth hot path (BASELINE.json metric: objects/sec, edge solve +
GMW aggregate).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference algorithm on the host CPU

A *step* is one forward pass of the GMW pipeline (compute_z -> edge-weight MLP -> softmax-weighted
depth, GMW/main.py:524-533) over one KITTI-val-shaped synthetic batch (BASELINE configs[1]:
3769 frames, <= 50 objects per frame, 73 keypoints).  With N GPUs every rank owns a contiguous
shard of N x 3769 frames (weak scaling) and the step ends with the all-gather of the per-object
depths.  One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "objects/sec (edge solve + GMW aggregate)"
UNIT = "objects/s"
N_KPTS = 73
EDGES = N_KPTS * (N_KPTS - 1) // 2
K_SEL = 1500
DEPTH = 12
FRAMES = 3769
WEIGHT_SEED = 7
# algorithmic work per object (SURVEY.md section 8d / BASELINE.md section 4)
F_SOLVE = 6 * N_KPTS + 11 * EDGES                 # 29 346 FLOP
B_SOLVE = 20 * N_KPTS + 56                        # 1516 B in + mean out
F_MLP_REF = EDGES * (2 * 128 * (4 + 6) + 36 * 2 * 2 * 128 * 128)     # 6.207 GFLOP: the reference's 36 GEMM layers per net
# the fused forward folds preconv.conv1 (no non-linearity between them): 24 GEMM layers per net are executed, and only
# those are counted
F_MLP = EDGES * (2 * 128 * (4 + 6) + 24 * 2 * 2 * 128 * 128)         # 4.140 GFLOP (both nets, GEMM FLOPs only)
FUSED_TRAFFIC_2048 = 5.61e9      # ncu: dram read 15 MB + write 5.59 GB for one fused launch on 2048 objects


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dcd_b200", choices=["dcd_b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU (default: the KITTI val split)")
    ap.add_argument("--chunk", type=int, default=2048, help="objects per MLP workspace chunk")
    ap.add_argument("--cpu-sample", type=int, default=512, help="objects of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        sm.append(float(c[1]))
                        smax.append(float(c[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


class CpuReference:
    """The reference's algorithm for this path on the host CPU (oracle port, `faithful` mode: the per-pair
    Python loops of get_up and the E x E distance matrix, i.e. what the reference executes; the Sinkhorn
    branch is outside the path), batch 8 like the reference's `-b 8`, all host threads."""

    def __init__(self):
        from oracle import dcd_oracle as O
        self.O = O
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = O.random_state_dict(WEIGHT_SEED)

    def run(self, objects: int, seed: int) -> float:
        """Process `objects` synthetic objects; returns seconds."""
        from dcd_b200 import synth
        ob = synth.make_objects(N=objects, n=N_KPTS, seed=seed)
        t0 = time.perf_counter()
        with torch.no_grad():
            for lo in range(0, objects, 8):
                self.O.gmw_pipeline(ob.kps_norm[lo:lo + 8], ob.kps_3d[lo:lo + 8], ob.rot_y[lo:lo + 8], self.sd, faithful=True)
        return time.perf_counter() - t0


def cpu_reference_rate(sample_objects: int, seed: int):
    ref = CpuReference()
    ref.run(8, seed + 1)                      # warm-up (thread pools, oneDNN primitives)
    secs = ref.run(sample_objects, seed)
    return sample_objects / secs, secs, ref.cores


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU reference arm; other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ref = CpuReference()
    sample = max(8, args.cpu_sample // 2)
    for w in range(args.warmup):
        ref.run(8, 1 + w)
    secs = sum(ref.run(sample, 100 + s) for s in range(args.steps))
    rate = sample * args.steps / secs
    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: KITTI-val-shaped synthetic batch (3769 frames x U{1..50} objects/frame, 73 keypoints): "
                               "compute_z + GMW.forward reg branch + weighted depth, forward only; each step is a bounded "
                               "sample of %d objects in batches of 8" % sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": ref.cores, "kind": "port",
                         "sample": "%d objects per step in batches of 8, torch %s CPU, oracle port in faithful mode "
                                   "(Python get_up loops + E x E distance matrix; Sinkhorn branch excluded)" % (sample, torch.__version__)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    import dcd_b200
    from dcd_b200 import _lib, synth
    from dcd_b200 import dist as ddist
    from dcd_b200._lib import check, ptr, stream_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (dcd_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L = _lib.lib()
    peaks = measured_peaks()

    # ---- workload: world x FRAMES frames, sharded contiguously by frame (weak scaling)
    seed = synth.BASE_SEED + 1
    counts = synth.frame_counts(args.frames * world, 50, True, seed)
    bounds = ddist.shard_bounds(counts.tolist(), world)
    cum = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    lo_obj, hi_obj = bounds[rank]
    f_lo = int((cum == lo_obj).nonzero()[0]) if lo_obj < int(cum[-1]) else len(counts)
    f_hi = int((cum == hi_obj).nonzero()[-1])
    ob = synth.make_objects(n=N_KPTS, seed=seed + 1000 * rank, counts=counts[f_lo:f_hi])
    N = ob.N
    assert N == hi_obj - lo_obj
    N_total = int(cum[-1])
    model = dcd_b200.GMW().to(dev).load_reference_state_dict(synth.random_state_dict(WEIGHT_SEED))
    p4, p6 = model.params4.detach(), model.params6.detach()

    # host (pinned) and device copies of the inputs
    h_k2, h_k3, h_rot = ob.kps_norm.pin_memory(), ob.kps_3d.pin_memory(), ob.rot_y.reshape(-1).contiguous().pin_memory()
    d_k2, d_k3, d_rot = h_k2.to(dev), h_k3.to(dev), h_rot.to(dev)
    h2d_bytes = (h_k2.numel() + h_k3.numel() + h_rot.numel()) * 4
    chunk = min(args.chunk, N)
    ws = torch.empty((L.dcd_gmw_workspace_bytes(chunk, N_KPTS, DEPTH, 0) // 4 + 64,), dtype=torch.float32, device=dev)
    idx = torch.empty((chunk, K_SEL), dtype=torch.int64, device=dev)
    zsel = torch.empty((chunk, K_SEL), dtype=torch.float32, device=dev)
    regw = torch.empty((chunk, EDGES), dtype=torch.float32, device=dev)
    depth_out = torch.empty((N,), dtype=torch.float32, device=dev)
    h_out = torch.empty((N,), dtype=torch.float32).pin_memory()
    launches = [0]
    mlp_events = []

    def step(k2, k3, rot, record_mlp: bool):
        """select -> edge-weight MLP -> aggregate per chunk, through the C ABI on the current stream."""
        st = stream_ptr()
        for c0 in range(0, N, chunk):
            nc = min(chunk, N - c0)
            a2, a3, ar = k2[c0:c0 + nc], k3[c0:c0 + nc], rot[c0:c0 + nc]
            check(L.dcd_edge_select_fwd(ptr(a2), ptr(a3), ptr(ar), 0, 0, nc, N_KPTS, K_SEL, 0.1, 80.0, 0,
                                        ptr(idx), ptr(zsel), 0, 0, st), "select")
            if record_mlp:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            check(L.dcd_gmw_weights_fwd(ptr(a2), ptr(a3), ptr(p4), ptr(p6), nc, N_KPTS, DEPTH, 0, ptr(regw), 0, 0,
                                        ptr(ws), ws.numel() * 4, st), "mlp")
            if record_mlp:
                e1.record()
                mlp_events.append((e0, e1, nc))
            check(L.dcd_gmw_aggregate_fwd(ptr(regw), ptr(zsel), ptr(idx), nc, EDGES, K_SEL, 1, ptr(depth_out[c0:]), 0, st),
                  "aggregate")
            # select | layer folding, weight image, fused MLP (all layers of both nets), edge weights | aggregate
            launches[0] += 1 + 4 + 1

    def gather():
        if world > 1:
            return ddist.all_gather_depths(depth_out, bounds)
        return depth_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        step(d_k2, d_k3, d_rot, False)
        gather()
    launches[0] = 0
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for _ in range(args.steps):
        step(d_k2, d_k3, d_rot, True)
        full = gather()
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = t_beg.elapsed_time(t_end)
    timed_launches = launches[0]
    mlp_ms = sum(a.elapsed_time(b) for a, b, _ in mlp_events)
    mlp_objs = sum(nc for _, _, nc in mlp_events)
    assert full.numel() == N_total

    # ---- end-to-end timing through the public API with HOST buffers (e2e)
    def e2e_step():
        k2 = h_k2.to(dev, non_blocking=True)
        k3 = h_k3.to(dev, non_blocking=True)
        rot = h_rot.to(dev, non_blocking=True)
        out = dcd_b200.gmw_weighted_depth(k2, k3, rot, model, chunk=chunk)
        if world > 1:
            out = ddist.all_gather_depths(out, bounds)[lo_obj:hi_obj]
        h_out.copy_(out, non_blocking=True)
    e2e_step()
    barrier()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record()
    for _ in range(args.steps):
        e2e_step()
    e_end.record()
    barrier()
    e2e_ms = e_beg.elapsed_time(e_end)
    assert torch.isfinite(h_out).all()

    # ---- stage breakdown of the DGDE pipeline (edge solve + mean) on the same objects, kernel-only
    d_kps, d_K = ob.kps.to(dev), ob.K.to(dev)
    mean = torch.empty((N,), dtype=torch.float32, device=dev)
    edges = torch.empty((N, EDGES), dtype=torch.float32, device=dev)

    def time_kernel(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    st = stream_ptr()
    ms_mean = time_kernel(lambda: check(L.dcd_edge_solve_fwd(ptr(d_kps), ptr(d_k3), ptr(d_rot), ptr(d_K), N, N_KPTS, 2.0, 80.0, 3,
                                                             0, ptr(mean), st), "solve"))
    ms_edges = time_kernel(lambda: check(L.dcd_edge_solve_fwd(ptr(d_kps), ptr(d_k3), ptr(d_rot), ptr(d_K), N, N_KPTS, 2.0, 80.0, 3,
                                                              ptr(edges), 0, st), "solve"))
    sel_n = min(N, 8192)
    idx_b = torch.empty((sel_n, K_SEL), dtype=torch.int64, device=dev)
    z_b = torch.empty((sel_n, K_SEL), dtype=torch.float32, device=dev)
    ms_sel = time_kernel(lambda: check(L.dcd_edge_select_fwd(ptr(d_kps), ptr(d_k3), ptr(d_rot), ptr(d_K), 0, sel_n, N_KPTS, K_SEL,
                                                             2.0, 80.0, 3, ptr(idx_b), ptr(z_b), 0, 0, st), "select"), reps=5)
    del edges
    # frame epilogue: image-space keypoints -> edge solve + mean -> 3D location, one launch
    pad_o = torch.tensor([19.0, 5.0], device=dev).expand(N, 2).contiguous()
    ctr = (d_kps.mean(1) + pad_o) / 4
    pts_o = ctr.floor()
    ofs_o = (ctr - pts_o).contiguous()
    off_o = ((d_kps + pad_o.unsqueeze(1)) / 4 - ctr.unsqueeze(1)).contiguous()
    dims_o = torch.stack((torch.full((N,), 3.9, device=dev), -d_k3[:, -1, 1], torch.full((N,), 1.6, device=dev)), dim=1).contiguous()
    loc_o = torch.empty((N, 3), dtype=torch.float32, device=dev)
    ms_loc = time_kernel(lambda: check(L.dcd_dgde_locate_fwd(ptr(off_o), ptr(d_k3), ptr(d_rot), ptr(d_K), ptr(pts_o), ptr(ofs_o),
                                                             ptr(pad_o), ptr(dims_o), 0, N, N_KPTS, 2.0, 80.0, 3, 4.0, ptr(mean),
                                                             ptr(loc_o), st), "locate"))

    # one frame of the detector head (50 detections) through the public API: POI gather on a [1,440,96,320] regression map,
    # fused keypoints -> edge solve -> location, depth ensemble: latency, launch overheads included
    fr = 50
    fmap = torch.randn((1, 440, 96, 320), device=dev)
    fidx = torch.randint(0, 96 * 320, (1, fr), device=dev)
    P_np = ob.K[0].double().numpy()
    lu_k = torch.zeros((fr, 3), device=dev)

    def frame_epilogue():
        dcd_b200.select_point_of_interest(1, fidx, fmap)
        d, _ = dcd_b200.compute_pairs_kpts_depth(off_o[:fr], pts_o[:fr], ofs_o[:fr], pad_o[:1], d_k3[:fr], d_rot[:fr], P_np,
                                                 dims=dims_o[:fr], return_locations=True)
        dcd_b200.depth_ensemble(off_o[:fr, -10:], dims_o[:fr], P_np, lu_k, direct_depths=d, direct_log_uncertainty=lu_k[:, 0])
    ms_frame = time_kernel(frame_epilogue, reps=20)

    # ---- GMW training step, batch 8 per GPU (BASELINE configs[2]): compute_z + forward (saved) + loss + backward
    tb = 8
    t_k2, t_k3, t_rot, t_gt = d_k2[:tb].contiguous(), d_k3[:tb].contiguous(), d_rot[:tb].reshape(-1, 1).contiguous(), ob.gt_depth[:tb].to(dev)
    train_model = dcd_b200.GMW().to(dev).load_reference_state_dict(synth.random_state_dict(WEIGHT_SEED))

    def train_step():
        train_model.zero_grad(set_to_none=True)
        Zt, idxt = dcd_b200.compute_z(t_k2, t_k3, t_rot)
        wt, _ = train_model(t_k2, t_k3, t_rot, None)
        loss, _ = dcd_b200.compute_reg_loss(Zt, wt, t_gt, idxt)
        loss.backward()
        if world > 1:
            ddist.allreduce_gradients(train_model)
    ms_train = time_kernel(train_step, reps=5)
    # correspondence branch, forward: E x E distances + Sinkhorn for the same 8 objects, P not materialised
    ms_cls = time_kernel(lambda: train_model.edge_transport(t_k2, t_k3, materialise=False), reps=3)

    # ---- reduce over ranks (max time) and report
    times = torch.tensor([ms, e2e_ms, mlp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, mlp_ms_max = [float(x) for x in times]
    if rank == 0:
        props = torch.cuda.get_device_properties(dev)
        sm_max_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        fp32_peak = props.multi_processor_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12            # TFLOP/s
        value = N_total * args.steps / (ms * 1e-3)
        e2e_value = N_total * args.steps / (e2e_ms * 1e-3)
        mlp_tflops = F_MLP * mlp_objs / (mlp_ms * 1e-3) / 1e12
        n_mlp_launches = len(mlp_events)
        out_bytes = 2 * 128 * 128 * ((EDGES + 127) // 128) * 4            # final features of both nets, per object
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: KITTI-val-shaped synthetic batch, %d frames x U{1..50} objects/frame "
                                   "(%d objects per GPU), 73 keypoints, 2628 edges: compute_z + edge-weight MLP + "
                                   "softmax-weighted depth, forward only%s" % (args.frames, N, ", + all-gather of depths" if world > 1 else ""),
                       "objects_per_gpu": N, "objects_total": N_total, "chunk_objects": chunk, "net_depth": DEPTH,
                       "l2_policy": "every chunk streams %.1f GB of final features through L2 (126 MB): nothing survives between "
                                    "chunks or steps" % (min(chunk, N) * 2 * 128 * 128 * ((EDGES + 127) // 128) * 4 / 1e9),
                       "weights": "random init, seed %d, reference state_dict layout" % WEIGHT_SEED},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": N * 4,
                    "api": "dcd_b200.gmw_weighted_depth on pinned host tensors"},
            "gpu_launches": timed_launches,
            "roofline": {"kernel": "mlp_fused_kernel (edge-feature MLP, the whole net of an object in one launch by a group of 8 co-resident CTAs: activations "
                                   "stay in shared/tensor memory; preconv.conv1 folded: 24 GEMM layers x 2 nets on tcgen05, FP16x3 split, FP32 accumulate "
                                   "in TMEM)",
                         "bound": "tensor",
                         "achieved": mlp_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": mlp_tflops / peaks["bf16_tflops_sustained"],
                         # ncu dram__bytes_read+write of ONE mlp_fused_kernel launch on a 2048-object chunk
                         # (profiles/r01_mlp_fused_kernel.md); algorithmic: 2048 x 2.75 MB of final features out
                         "traffic": FUSED_TRAFFIC_2048 if chunk == 2048 else None,
                         "peak_source": "%s dense bf16 (sustained, of measured); `achieved` counts the algorithmic FP32 GEMM FLOPs, the "
                                        "tensor pipe executes 3 FP16 MMAs per FP32 product (x3 = %.1f TFLOP/s issued); vs the FP32 "
                                        "CUDA-core roofline (%.1f TFLOP/s) the same number is %.2fx. The events bracket dcd_gmw_weights_fwd "
                                        "(layer folding + weight image, fused MLP, edge weights; the fused kernel is ~96%% of it)" % (
                                            peaks["source"], 3 * mlp_tflops, fp32_peak, mlp_tflops / fp32_peak),
                         "hbm_gbs": out_bytes * mlp_objs / (mlp_ms * 1e-3) / 1e9,
                         "frac_hbm": out_bytes * mlp_objs / (mlp_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "hbm_bytes_per_object": out_bytes, "hbm_bytes_per_object_layerwise": 198.0e6,
                         "frac_fp32_roofline": mlp_tflops / fp32_peak,
                         "flops_per_object": F_MLP, "flops_per_object_unfolded_reference": F_MLP_REF,
                         "avg_launch_ms": mlp_ms / max(n_mlp_launches, 1),
                         "share_of_step": mlp_ms_max / ms},
            "stages": {
                "dgde_solve_mean": {"objects_per_s": N / (ms_mean * 1e-3), "ms": ms_mean,
                                    "fp32_tflops": F_SOLVE * N / (ms_mean * 1e-3) / 1e12,
                                    "frac_fp32_roofline": F_SOLVE * N / (ms_mean * 1e-3) / 1e12 / fp32_peak,
                                    "hbm_gbs": B_SOLVE * N / (ms_mean * 1e-3) / 1e9},
                "dgde_solve_edges": {"objects_per_s": N / (ms_edges * 1e-3), "ms": ms_edges,
                                     "hbm_gbs": (B_SOLVE + 4 * EDGES) * N / (ms_edges * 1e-3) / 1e9,
                                     "frac_hbm": (B_SOLVE + 4 * EDGES) * N / (ms_edges * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "dgde_frame_epilogue": {"objects_per_s": N / (ms_loc * 1e-3), "ms": ms_loc,
                                        "fp32_tflops": F_SOLVE * N / (ms_loc * 1e-3) / 1e12,
                                        "hbm_gbs": (B_SOLVE + 52) * N / (ms_loc * 1e-3) / 1e9,
                                        "what": "keypoint offsets -> image keypoints -> edge solve + mean -> 3D location "
                                                "(detector_infer.py:215-227,186-188), one launch"},
                "dgde_frame_latency": {"us": ms_frame * 1e3, "objects": fr,
                                       "what": "one frame, 50 detections, python API: select_point_of_interest + compute_pairs_kpts_depth "
                                               "(with locations) + depth_ensemble; three launches, host overheads included"},
                "edge_select_top1500": {"objects_per_s": sel_n / (ms_sel * 1e-3), "ms": ms_sel, "objects": sel_n},
                "gmw_train_step_b8": {"ms": ms_train, "objects_per_s": tb * world / (ms_train * 1e-3),
                                      "what": "configs[2]: compute_z + edge MLP fwd + softmax aggregate + L1 loss + full backward (all GEMMs on "
                                              "tcgen05) for 8 objects per GPU%s; optimizer step excluded" % (" + gradient all-reduce" if world > 1 else "")},
                "gmw_cls_forward_b8": {"ms": ms_cls, "objects_per_s": tb / (ms_cls * 1e-3),
                                       "what": "edge MLP + E x E feature distances + Sinkhorn (lambda 10, <= 100 iterations, "
                                               "GMW/model/model.py:170-192) -> sum P, trace P for 8 objects; forward only"},
                "fp32_peak_tflops": fp32_peak,
            },
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            rate, secs, cores = cpu_reference_rate(args.cpu_sample, 99)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d objects in batches of 8 (%.1f s), oracle port in faithful mode on the host CPU"
                                              % (args.cpu_sample, secs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-GPU path on real devices (skipped on a single-GPU box): two NCCL ranks run the sharded GMW pipeline of bench.py and
rank 0 verifies SURVEY 8e's correctness check — the depths gathered from the other rank are bit-identical to a single-rank
recomputation of that rank's shard (same kernels, same per-object arithmetic, whatever the chunking)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two CUDA devices")
@pytest.mark.parametrize("config,extra", [("kitti_val", ["--frames", "60"]), ("sweep1m", ["--objects", "20000"])])
def test_two_rank_nccl_run_matches_single_rank_bits(config, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1", "--quick",
           "--no-cpu-baseline", "--config", config, "--chunk", "512"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0
    assert line["shard_check"]["bit_identical"] is True and line["shard_check"]["objects"] > 0
    if config == "sweep1m":
        assert line["scaling"] == "strong" and line["stages"]["dgde_pipeline"]["bit_identical_to_single_rank"] is True

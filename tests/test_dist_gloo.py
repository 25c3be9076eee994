"""CPU, world_size 2 over gloo: the multi-GPU host logic (frame sharding, all-gather of per-object depths,
DDP-style gradient averaging).  The per-shard compute is stood in by the oracle (tests may use it); the
R-rank result must equal the 1-rank result bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcd_b200 import dist as ddist
from dcd_b200 import synth
from oracle import dcd_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        ob = synth.make_objects(n=73, seed=42, counts=torch.tensor(counts))
        bounds = ddist.shard_bounds(counts, world)
        lo, hi = bounds[rank]
        local = O.dgde_pipeline(ob.kps[lo:hi], ob.kps_3d[lo:hi], ob.rot_y[lo:hi], ob.K[lo:hi])
        full = ddist.all_gather_depths(local, bounds)
        single = O.dgde_pipeline(ob.kps, ob.kps_3d, ob.rot_y, ob.K)
        assert full.shape == single.shape
        assert torch.equal(full[lo:hi], single[lo:hi])
        assert float((full - single).abs().max()) <= 1e-5 * float(single.abs().max())

        # gradient averaging: every rank contributes rank-dependent grads; result = mean over ranks
        class M:
            pass
        m = M()
        m.params4 = torch.nn.Parameter(torch.zeros(1000))
        m.params6 = torch.nn.Parameter(torch.zeros(1200))
        m.params4.grad = torch.full((1000,), float(rank + 1))
        m.params6.grad = torch.arange(1200, dtype=torch.float32) * (rank + 1)
        works = ddist.allreduce_gradients(m, async_op=True)
        for w in works:
            w.wait()
        mean_scale = sum(range(1, world + 1)) / world
        assert torch.allclose(m.params4.grad, torch.full((1000,), mean_scale))
        assert torch.allclose(m.params6.grad, torch.arange(1200, dtype=torch.float32) * mean_scale)
        if rank == 0:
            torch.save(full, os.path.join(out_dir, "full.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[3, 5, 2, 4], [7, 1, 1], [4]])
def test_two_rank_shard_gather_equals_single_rank(tmp_path, counts):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, counts, str(tmp_path)), nprocs=2, join=True)
    full = torch.load(os.path.join(str(tmp_path), "full.pt"))
    assert full.numel() == sum(counts)

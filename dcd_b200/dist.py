"""Multi-GPU plumbing: one process per GPU, objects sharded by image frame.

Objects are independent, so the data path needs no collective at all; the only exchanges are
  * forward:  ONE all-gather of the per-object depths ([N/R] FP32 per rank -> [N] on every rank);
  * training: ONE all-reduce per edge net of the GMW weight gradient (the DDP all-reduce of
              GMW/main.py:252), summed then divided by the world size like DistributedDataParallel.
Both go through torch.distributed (NCCL over NVLink/NVSwitch on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(counts: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous frame shards balanced by object count.

    counts[f] = objects of frame f.  Returns per-rank (lo, hi) OBJECT ranges; every frame goes to
    exactly one rank and rank r gets the frames whose cumulative object count first reaches
    r+1 shares of the total (so shards differ by at most one frame's worth of objects).
    """
    counts = [int(c) for c in counts]
    total = sum(counts)
    bounds = []
    frame = 0
    done = 0
    for r in range(world_size):
        lo = done
        target = (total * (r + 1) + world_size - 1) // world_size if r + 1 < world_size else total
        last = r + 1 == world_size                 # the last rank also takes trailing frames without objects
        while frame < len(counts) and (done < target or last):
            done += counts[frame]
            frame += 1
        bounds.append((lo, done))
    assert done == total and frame == len(counts)
    return bounds


def all_gather_depths(local: torch.Tensor, bounds: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
    """All-gather per-object depths: `local` [hi-lo] of this rank -> [N] on every rank.

    Shards are padded to the largest one so that a single fixed-size all_gather_into_tensor
    (ncclAllGather) moves everything; pads are stripped afterwards.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [hi - lo for lo, hi in bounds]
    assert len(sizes) == world and local.numel() == sizes[rank]
    m = max(sizes) if sizes else 0
    send = local.new_zeros((m,))
    send[: sizes[rank]] = local.reshape(-1)
    recv = local.new_empty((world * m,))
    if m:
        dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, m)
    if all(s == m for s in sizes):
        return recv.reshape(-1)
    return torch.cat([recv[r, : sizes[r]] for r in range(world)])


def allreduce_gradients(model, group=None, async_op: bool = False):
    """DDP-equivalent gradient averaging of the two parameter blobs of dcd_b200.ops.GMW.

    One all-reduce per edge net (2.38 MB each), called after backward() has produced both blobs
    (at 4.76 MB the exchange costs < 0.1 ms on NVSwitch, so nothing is overlapped with the backward).
    Returns the work handles when async_op=True.
    """
    world = dist.get_world_size(group)
    works = []
    for p in (model.params4, model.params6):
        if p.grad is None:
            continue
        p.grad.div_(world)
        w = dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def local_slice(t: Optional[torch.Tensor], bounds: Sequence[Tuple[int, int]], rank: Optional[int] = None):
    if t is None:
        return None
    rank = dist.get_rank() if rank is None else rank
    lo, hi = bounds[rank]
    return t[lo:hi]

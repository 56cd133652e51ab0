"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).  The reference is single-device; this is the
B200 scaling layer of SURVEY.md section 8(e):

* sampling shards the particle set -- no data-path collective; in-kernel Philox noise is keyed by the GLOBAL particle
  index (``particle_offset``), so any sharding gives bit-identical particles;
* training shards the batch and all-reduces ONE flat gradient buffer per iteration (34 k floats for the MLP).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [start, start+count) of ``total`` rows for ``rank``; the remainder goes to the first ranks."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def sample_sharded(sampler, sde, x_0_global, num_steps, seed, **kw):
    """Run ``sampler`` (one of sde_scheme's) on this rank's rows of ``x_0_global``; returns (start, local_result)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    start, count = shard_range(x_0_global.shape[0], rank, world)
    out = sampler(sde, x_0_global[start:start + count], num_steps, seed=seed, particle_offset=start, **kw)
    return start, out


def allreduce_grads_(params, average: bool = True, group=None):
    """All-reduce the ``.grad`` of ``params`` through one flat buffer (one collective per iteration)."""
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    o = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n

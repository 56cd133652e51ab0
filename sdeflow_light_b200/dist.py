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


class P2PAllreduceAdam:
    """Handle of the peer-memory all-reduce + Adam (csrc/p2p.cu) for one flat gradient of ``nfloats`` floats: allocates this
    rank's receive buffer, exchanges the CUDA IPC handles over ``torch.distributed`` and opens the peers' buffers.  All ranks
    of the group must live on one node (NVLink / PCIe peer access); construction raises otherwise and the caller keeps the
    NCCL path."""

    def __init__(self, device, nfloats: int, group=None):
        import ctypes as C
        from . import _lib
        self.world, self.rank, self._group = dist.get_world_size(group), dist.get_rank(group), group
        self.dev = torch.device(device)
        self._h = C.c_void_p()
        buf = C.create_string_buffer(64)
        _lib.check(_lib.lib().msgm_p2p_create(_lib.ctx(self.dev), int(nfloats), self.world, self.rank, C.byref(self._h), buf))
        mine = torch.tensor(list(buf.raw), dtype=torch.uint8, device=self.dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        blob = bytes(torch.cat(allh).cpu().tolist())
        _lib.check(_lib.lib().msgm_p2p_connect(_lib.ctx(self.dev), self._h, blob))
        dist.barrier(group)  # nobody pushes before every rank has opened every buffer

    def step(self, seg_table, n_tensors, total, grad_flat, exp_avg, exp_avg_sq, lr, adam_step):
        from . import _lib
        _lib.check(_lib.lib().msgm_p2p_allreduce_adam(
            _lib.ctx(self.dev), self._h, _lib.ptr(seg_table), n_tensors, total, _lib.ptr(grad_flat), _lib.ptr(exp_avg),
            _lib.ptr(exp_avg_sq), _lib.ptr(lr), _lib.ptr(adam_step), 0.9, 0.999, 1e-8, _lib.stream_ptr(self.dev)))

    def close(self):
        """Collective teardown: every rank unmaps its peers' buffers, the ranks meet, then each frees its own."""
        from . import _lib
        if self._h:
            torch.cuda.synchronize(self.dev)
            _lib.lib().msgm_p2p_disconnect(self._h)
            if dist.is_initialized():
                dist.barrier(self._group)
            _lib.lib().msgm_p2p_destroy(self._h)
            self._h = None

"""CPU, world_size 2, gloo: the host-side multi-GPU logic (sharding arithmetic, flat-buffer gradient all-reduce)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdeflow_light_b200 import dist as D


def test_shard_range_partitions_everything():
    for total in (0, 1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                s, c = D.shard_range(total, r, world)
                assert s == pos and c >= 0
                pos += c
            assert pos == total
            counts = [D.shard_range(total, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    D.allreduce_grads_(net.parameters(), average=True)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
    # sharded "sampler": every rank gets disjoint rows and the right global offset
    seen = {}

    def fake_sampler(sde, x, n, seed=None, particle_offset=0, **kw):
        seen["off"], seen["rows"] = particle_offset, x.shape[0]
        return x + particle_offset

    start, out = D.sample_sharded(fake_sampler, None, torch.zeros(11, 2), 4, seed=1)
    ok = ok and seen["off"] == start and out.shape[0] == D.shard_range(11, rank, world)[1]
    q.put((rank, ok, start, out.shape[0]))
    dist.destroy_process_group()


def test_flat_allreduce_and_sharded_sampling_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert res[0][2] == 0 and res[1][2] == res[0][3] and res[0][3] + res[1][3] == 11

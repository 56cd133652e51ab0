"""torchrun --nproc-per-node N tools/p2p_check.py : the peer-memory all-reduce + Adam (csrc/p2p.cu) against the NCCL path.

Every rank builds the same generative SDE (same seed), trains `iters` graphed iterations on its own shard of one global batch
with (a) p2p=True and (b) p2p=False (NCCL all-reduce inside the graph), same Philox seed, and checks that
  * the parameters of all ranks are bit-identical after the p2p run (rank-ordered sum),
  * the p2p and NCCL runs agree to rounding,
and prints the iteration time of both."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from sdeflow_light_b200.train import GraphedSsmStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d, batch, iters = 8, int(os.environ.get("P2P_BATCH", 256)), 20
prob = bench.build_problem(d)
data = prob["data"].to(dev)
results = {}


def say(*a):
    if rank == 0:
        print("[p2p_check]", *a, file=sys.stderr, flush=True)


for mode in ("p2p", "nccl"):
    say("mode", mode)
    P, gen = bench.package_objects(prob, dev)
    gen.ssm_precision = os.environ.get("P2P_PREC", "fp32")
    step = GraphedSsmStep(gen, (batch, d), lr=1e-3, seed=11, p2p=(mode == "p2p"))
    assert (step.p2p is not None) == (mode == "p2p")
    say("trainer built")
    x = data[(torch.arange(batch, device=dev) + rank * batch) % data.shape[0]]
    for _ in range(iters):
        loss = step(x)
    torch.cuda.synchronize()
    say("20 iterations done")
    flat = torch.cat([p.detach().reshape(-1) for p in gen.a.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    say("gathered")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    step.close()
    results[mode] = dict(params=flat.clone(), same=same, ms=float(ms), loss=float(loss), launches=step.launches_per_iter,
                         flags=P._lib.debug_flags(dev))
diff = float((results["p2p"]["params"] - results["nccl"]["params"]).abs().max())
if rank == 0:
    print(json.dumps({"world": world, "batch_per_gpu": batch, "precision": os.environ.get("P2P_PREC", "fp32"),
                      "ranks_bit_identical_p2p": results["p2p"]["same"], "ranks_bit_identical_nccl": results["nccl"]["same"],
                      "max_param_diff_p2p_vs_nccl_after_20_iters": diff, "ms_per_iter_p2p": results["p2p"]["ms"],
                      "ms_per_iter_nccl": results["nccl"]["ms"], "launches_p2p": results["p2p"]["launches"],
                      "launches_nccl": results["nccl"]["launches"], "flags": [results["p2p"]["flags"], results["nccl"]["flags"]]}))
ok = results["p2p"]["same"] and diff < 1e-4 and results["p2p"]["flags"] == 0
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0 if ok else 1)  # skip the process-group teardown: live CUDA graphs hold captured NCCL work

"""One score-matching training iteration as a CUDA graph.

The reference's training loop (MSGM_higherDim.py:803-809) is

    optim.zero_grad(); loss = gen_sde.ssm(x).mean(); loss.backward(); optim.step()

At the reference's batch size (256) that iteration is a handful of ~10 us kernels and the host (Python, the allocator,
~25 launches) is the bottleneck.  ``GraphedSsmStep`` records exactly that sequence once -- device-side draws of t, the
forward-noising launch, the probe v, the fused SSM forward/backward kernels, the flat gradient all-reduce when several
ranks train together, and the fused Adam update -- into CUDA graphs and replays them, one ``cudaGraphLaunch`` per
iteration.  Every random draw inside the graph comes from torch's CUDA Philox generator, whose offset advances with
each replay, so iterations see fresh t / noise / v.

    step = GraphedSsmStep(gen_sde, lr=1e-3, batch_shape=(256, d))
    for it in range(n):
        loss = step(x_batch)        # x_batch: (256, d) tensor on the device (or host: copied into the static input)
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class _SsmModule(torch.nn.Module):
    """``forward = gen.ssm`` so that torch.func.functional_call can run the loss on substitute parameter leaves."""

    def __init__(self, gen):
        super().__init__()
        self.gen = gen

    def forward(self, x):
        return self.gen.ssm(x)


class GraphedSsmStep:
    def __init__(self, gen, batch_shape, lr: float = 1e-3, optimizer: torch.optim.Optimizer | None = None,
                 warmup: int = 3, group=None, seed: int | None = None, graph_allreduce: bool = True,
                 p2p: bool | None = None):
        dev = torch.device(gen.deviceReverseSDE)
        if dev.type != "cuda":
            raise RuntimeError("sdeflow_light_b200 runs on CUDA only (no CPU fallback)")
        self.gen, self.dev, self.group = gen, dev, group
        self._names = [n for n, p in gen.named_parameters() if p.requires_grad]
        self.params = [p for n, p in gen.named_parameters() if p.requires_grad]
        self._mod = _SsmModule(gen)
        # Default optimiser: the reference driver's Adam(lr) (MSGM_higherDim.py:792) as ONE launch over the flat gradient
        # buffer (msgm_adam_step; lr and the update counter are device scalars, so schedulers can change lr between replays
        # with set_lr).  A caller-supplied torch optimiser (capturable=True) is recorded instead when given.
        self.opt = optimizer
        if optimizer is not None:
            for g in self.opt.param_groups:
                if not g.get("capturable", False):
                    raise ValueError("the optimizer must be built with capturable=True to be replayed inside a CUDA graph")
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.x = torch.zeros(*batch_shape, device=dev, dtype=torch.float32)
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)
        # every .grad is a view into ONE flat buffer: the all-reduce needs no gather/scatter, Adam reads it in place
        n_flat = sum(p.numel() for p in self.params)
        self._flat_storage = torch.zeros((n_flat + 3) // 4 * 4, device=dev)  # padded: the peer push moves 16-byte words
        self.flat = self._flat_storage[:n_flat]
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        self._grads = [p.grad for p in self.params]
        if self.opt is None:
            self.lr = torch.tensor(float(lr), device=dev, dtype=torch.float32)
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
            self.adam_step = torch.zeros(1, device=dev, dtype=torch.int64)
            tab, o = torch.zeros(len(self.params), 2, dtype=torch.int64), 0
            for i, p in enumerate(self.params):
                if not (p.is_contiguous() and p.dtype == torch.float32):
                    raise ValueError("msgm_adam_step needs contiguous fp32 parameters")
                tab[i, 0], tab[i, 1] = p.data_ptr(), o
                o += p.numel()
            self._seg_table = tab.to(dev)
        # Several ranks on one node: the gradient all-reduce is fused with Adam over NVLink peer memory (csrc/p2p.cu) unless
        # p2p=False / MSGM_NO_P2P=1 or the peers' buffers cannot be opened (then: NCCL all-reduce inside the graph).
        self.p2p = None
        import os
        if self.world > 1 and self.opt is None and (p2p if p2p is not None else os.environ.get("MSGM_NO_P2P") != "1"):
            try:
                from .dist import P2PAllreduceAdam
                self.p2p = P2PAllreduceAdam(dev, n_flat, group)
            except (RuntimeError, ValueError) as exc:
                if p2p:
                    raise
                self.p2p_error = str(exc)
        self.launches_per_iter = 0
        # Device-side random streams: a fixed seed plus a device counter bumped inside the graph, so every replay draws
        # fresh t / noise / v; the row offset makes a batch sharded over ranks draw what a single rank would.
        self._iter = torch.zeros(1, device=dev, dtype=torch.int64)
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._old_rng = getattr(gen, "_rng", None)
        self._seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        self._row_offset = rank * int(batch_shape[0])
        gen._rng = (self._seed, self._iter, self._row_offset)
        # MLP score nets: forward and backward kernels back to back, gradients written straight into the flat buffer
        from . import NN, SDEs
        self._direct = (isinstance(gen.a, NN.MLP) and gen.a.fused_ok() and len(batch_shape) == 2 and batch_shape[1] <= 32
                        and gen.vtype in SDEs._VTYPES and isinstance(gen.base_sde, (SDEs.MSGMsde, SDEs.SGMsde))
                        and all(n.startswith("a.main.") for n in self._names))
        self._gout = torch.full((int(batch_shape[0]),), 1.0 / int(batch_shape[0]), device=dev)

        gen.train()
        was = getattr(gen, "device_rng", False)
        gen.device_rng = True
        # Warm-up runs real iterations (allocator pools, lazy module loads, optimizer state creation) on stand-in data:
        # snapshot parameters and optimizer state, and put them back -- in place, the graphs hold the addresses.
        saved_p = [p.detach().clone() for p in self.params]
        saved_s = {} if self.opt is None else {id(q): {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
                                               for q, st in self.opt.state.items()}
        self.x.normal_()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        # ONE graph per iteration: prologue, forward, backward, [NCCL all-reduce of the flat gradient -- NCCL collectives are
        # capturable --], Adam.  `graph_allreduce = False` keeps the collective between two graphs (fwd/bwd | all-reduce |
        # Adam), the round-1 arrangement, for process groups that cannot be captured.
        l0 = _lib.launch_count(dev)
        self.g_fb, self.g_opt = torch.cuda.CUDAGraph(), None
        from . import unet_train
        unet_train.LEAF_CAPTURE = True  # this capture joins the weight-gradient branch (_fwd_bwd): it may be forked
        try:
            if self.world == 1 or graph_allreduce:
                with torch.cuda.graph(self.g_fb):
                    self._iteration()
            else:
                with torch.cuda.graph(self.g_fb):
                    self._fwd_bwd()
        finally:
            unet_train.LEAF_CAPTURE = False
        if not (self.world == 1 or graph_allreduce):
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt):
                self._opt_step()
        self.launches_per_iter = _lib.launch_count(dev) - l0
        with torch.no_grad():
            for p, q in zip(self.params, saved_p):
                p.copy_(q)
            if self.opt is None:
                self.exp_avg.zero_()
                self.exp_avg_sq.zero_()
                self.adam_step.zero_()
            else:
                for q, st in self.opt.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            old = saved_s.get(id(q), {}).get(k)
                            v.copy_(old) if old is not None else v.zero_()
        self.x.zero_()
        self._iter.zero_()
        gen.device_rng = was
        gen._rng = self._old_rng  # the graphs have the trainer's stream baked in; eager calls keep their own

    # -- the recorded part of the iteration ---------------------------------------------------------------------
    def _iteration(self):
        self._fwd_bwd()
        if self.p2p is not None:  # push to the peers + wait / sum / Adam: two launches, no collective call
            self.p2p.step(self._seg_table, len(self.params), self.flat.numel(), self.flat, self.exp_avg, self.exp_avg_sq,
                          self.lr, self.adam_step)
            return
        if self.world > 1:  # summed over ranks; the 1/world is folded into the Adam kernel's gradient scale
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self._opt_step()

    def _opt_step(self):
        if self.opt is None:
            import ctypes as C  # noqa: F401
            _lib.check(_lib.lib().msgm_adam_step(
                _lib.ctx(self.dev), _lib.ptr(self._seg_table), len(self.params), self.flat.numel(), _lib.ptr(self.flat),
                _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.lr), _lib.ptr(self.adam_step), 0.9, 0.999,
                1e-8, 1.0 / self.world, _lib.stream_ptr(self.dev)))
        else:
            if self.world > 1:
                self.flat /= self.world
            self.opt.step()

    def _fwd_bwd(self):
        if self._direct:
            from . import ssm_fused
            t_, y, v = self.gen._prepare(self.x)
            loss = ssm_fused.fused_loss_and_grads(self.gen, t_, y, v, self._gout, self.flat)
            torch.mean(loss, dim=0, out=self.loss)
            self._iter.add_(1)
            return
        # The loss runs on fresh leaves aliasing the parameters, and autograd.grad + one multi-tensor copy replaces
        # .backward(): a parameter's AccumulateGrad node remembers the stream it was first used on, and one kept alive
        # by an earlier eager iteration (a retained loss, a live optimizer) would tie this capture to the default stream.
        leaves = [p.detach().requires_grad_(True) for p in self.params]
        loss = torch.func.functional_call(self._mod, {"gen." + n: l for n, l in zip(self._names, leaves)},
                                          (self.x,)).mean()
        grads = torch.autograd.grad(loss, leaves)
        from . import unet_train
        unet_train.join_leaf_stream(self.dev)  # weight gradients of the U-Net convs run on a side branch of the graph
        torch._foreach_copy_(self._grads, list(grads))
        self.loss.copy_(loss.detach())
        self._iter.add_(1)

    def close(self):
        """Release the peer-memory all-reduce buffers (collective: call on every rank before the process group goes away)."""
        if self.p2p is not None:
            self.p2p.close()
            self.p2p = None

    # -- optimiser state in torch.optim.Adam's layout (what NN.save_checkpoint / the reference's loader expect) ---------
    def state_dict(self) -> dict:
        """``torch.optim.Adam(...).state_dict()``-compatible view of the optimiser state, so that
        ``NN.save_checkpoint(path, gen, step, it)`` writes a file the reference's ``load_checkpoint`` reads into its own
        ``torch.optim.Adam`` (NN.py:24-27)."""
        if self.opt is not None:
            return self.opt.state_dict()
        # torch indexes the state by position in the list the optimiser was built from; the reference builds its Adam from
        # gen_sde.parameters() (MSGM_higherDim.py:792), which also holds the two frozen horizons T (no state entries)
        allp = list(self.gen.parameters())
        pos = {id(p): i for i, p in enumerate(allp)}
        state, o = {}, 0
        t = self.adam_step.detach().to(torch.float32).reshape(()).cpu()
        for p in self.params:
            n = p.numel()
            state[pos[id(p)]] = {"step": t.clone(), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                                 "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
            o += n
        group = {"lr": float(self.lr.item()), "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(allp)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        if self.opt is not None:
            return self.opt.load_state_dict(sd)
        pos = {id(p): i for i, p in enumerate(self.gen.parameters())}
        o, steps = 0, []
        with torch.no_grad():
            for p in self.params:
                n, st = p.numel(), sd["state"].get(pos[id(p)])
                if st is not None:
                    self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.append(int(float(st["step"])))
                o += n
            self.adam_step.fill_(max(steps) if steps else 0)
            self.lr.fill_(float(sd["param_groups"][0]["lr"]))

    # -- resumable random stream ---------------------------------------------------------------------------------
    def rng_state(self) -> dict:
        """Philox seed and device iteration counter of the recorded draws (t, noise, v).  Store it next to the model
        checkpoint (``NN.save_checkpoint(..., trainer=step)``): a run resumed with ``load_rng_state`` continues the random
        stream instead of repeating it from iteration 0."""
        return {"seed": int(self._seed), "iter": int(self._iter.item()), "row_offset": int(self._row_offset)}

    def load_rng_state(self, state: dict) -> None:
        if int(state["seed"]) != int(self._seed):
            raise ValueError("the Philox seed is baked into the recorded graphs: build GraphedSsmStep(seed=state['seed'])")
        self._iter.fill_(int(state["iter"]))

    def set_lr(self, lr: float):
        """Change the learning rate seen by the recorded Adam update (in place, on the device)."""
        if self.opt is None:
            self.lr.fill_(float(lr))
            return
        for g in self.opt.param_groups:
            if isinstance(g["lr"], torch.Tensor):
                g["lr"].fill_(float(lr))
            else:
                raise ValueError("this optimizer was built with a Python-float lr, which the graph has baked in")

    # -- one iteration -----------------------------------------------------------------------------------------
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Run one iteration on batch ``x``; returns the (device, 0-dim) mean loss of this rank's shard."""
        if x.shape != self.x.shape:
            raise ValueError(f"batch shape {tuple(x.shape)} differs from the recorded {tuple(self.x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.g_fb.replay()
        if self.g_opt is not None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.g_opt.replay()
        # the replayed Adam wrote the parameters without bumping Tensor._version: invalidate every weight-derived cache
        # (packed tensor-core images, captured inference graphs) so that the next sample / evaluate sees live weights
        _lib.bump_weight_epoch()
        return self.loss

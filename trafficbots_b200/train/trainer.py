"""`TrainState` -- flat parameter / gradient / Adam buffers + one training step (BASELINE.json configs[3]).

Stands where Lightning's `training_step` -> `backward` -> DDP all-reduce -> `optimizer.step` stands in the reference
(pl_modules/waymo_motion.py:356-418,955-973; configs/trainer/default.yaml:12 gradient clipping at 5).  All parameters live in
ONE flat fp32 buffer (layout = the order of the reference's `named_parameters()`, goal-predictor parameters last: they form
the second Adam parameter group, :957-966), all gradients in another: the data-parallel reduction is a single NCCL
all-reduce of 13.6 MB and the optimizer a single fused kernel (`tb_tr_sq_norm` + `tb_tr_adam_step`).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping, Optional, Tuple

import torch
from torch import Tensor

from .. import weights
from . import graph
from .tape import Fn


def trainable_names(state_dict: Mapping[str, Tensor]):
    """the reference's `named_parameters()`: every state_dict key except buffers (`pre_processing.*`) and the aliases of the
    shared cross-attention blocks (latent_encoder.py:39-41)."""
    names = [k for k in state_dict if not k.startswith("pre_processing.") and weights._alias_of(k) is None]
    main = [k for k in names if "goal_predictor" not in k]
    goal = [k for k in names if "goal_predictor" in k]
    return main, goal


def build_params(state_dict: Mapping[str, Tensor], device) -> Tuple[graph.Params, Tensor, Tensor, "OrderedDict[str, Tuple[int, tuple]]", int]:
    """-> (Params over views, flat parameters, flat gradients, {name: (offset, shape)}, end of parameter group 0)."""
    main, goal = trainable_names(state_dict)
    layout: "OrderedDict[str, Tuple[int, tuple]]" = OrderedDict()
    off = 0
    for k in main + goal:
        layout[k] = (off, tuple(state_dict[k].shape))
        off += (state_dict[k].numel() + 3) // 4 * 4  # 16-byte aligned views
        if k == main[-1]:
            end_main = off
    flat_p = torch.zeros(off, dtype=torch.float32, device=device)
    flat_g = torch.zeros(off, dtype=torch.float32, device=device)
    tensors, grads = {}, {}
    for k, (o, shape) in layout.items():
        n = state_dict[k].numel()
        tensors[k] = flat_p[o:o + n].view(shape)
        grads[k] = flat_g[o:o + n].view(shape)
        tensors[k].copy_(state_dict[k])
    buffers = {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in state_dict.items() if k.startswith("pre_processing.")}
    return graph.Params(tensors, grads, buffers), flat_p, flat_g, layout, end_main


class TrainState:
    def __init__(self, state_dict: Mapping[str, Tensor], device="cuda", ops=None, lr: float = 3e-4, lr_goal: Optional[float] = None,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: float = 5.0, p_rollout_prior: float = 0.1,
                 dropout_p: float = 0.0, n_split: int = 1) -> None:
        if ops is None:
            from .cuda_ops import CudaOps  # raises without the CUDA library / a CUDA device: there is no CPU training path
            ops = CudaOps(device)
        self.ops = ops
        self.device = torch.device(device)
        self.params, self.flat_p, self.flat_g, self.layout, end_main = build_params(state_dict, self.device)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.group_end = torch.tensor([end_main, self.flat_p.numel()], dtype=torch.int32, device=self.device)
        self.lr = torch.tensor([lr, lr if lr_goal is None else lr_goal], dtype=torch.float32, device=self.device)
        self.betas, self.eps, self.max_grad_norm = betas, eps, max_grad_norm
        self.p_rollout_prior = p_rollout_prior
        # dropout of the reference's training mode (one probability for every site, traffic_bots.yaml); 0 = the parity configuration
        self.dropout_p = float(dropout_p)
        self.drop_seed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_split = int(n_split)  # concurrent sub-batch chains per step (see forward_backward)
        self.stack_backward = None  # None: step-stacked backward whenever the back end supports it (graph.training_forward)
        self._streams, self._keep = None, None
        self.n_step = 0
        self.last_ops = 0
        self._graphs, self._pool, self._static_batch, self._static_eps = {}, None, None, None
        self.replayed_kernels = 0  # kernels of this library launched through graph replays (tb_launch_count sees eager launches only)

    def state_dict(self) -> Dict[str, Tensor]:
        return {k: v.detach().clone() for k, v in self.params.t.items()}

    def grads(self) -> Dict[str, Tensor]:
        return self.params.g

    def draw_noise(self, n_scene: int, n_agent: int, latent_dim: int = 16) -> Tuple[bool, Tensor]:
        """the two host-side random draws of the reference step, in its order (global torch CPU generator):
        `torch.rand(1) < p_training_rollout_prior` (:384), then the `rsample` noise (distributions.py:30)."""
        use_prior = bool(torch.rand(1) < self.p_rollout_prior)
        eps = torch.empty(n_scene, n_agent, latent_dim).normal_()
        return use_prior, eps

    def new_dropout_seed(self, seed: Optional[int] = None) -> None:
        """a fresh seed for this step's dropout masks (host draw from the global torch generator unless given); the kernels
        read it from device memory, so a captured CUDA graph sees the new value."""
        if self.dropout_p > 0.0:
            self.drop_seed.fill_(int(torch.randint(0, 2 ** 31 - 1, (1,))) if seed is None else int(seed))

    def forward_backward(self, batch: Mapping[str, Tensor], eps: Optional[Tensor] = None, use_prior: Optional[bool] = None,
                         return_buffers: bool = False, new_seed: bool = True) -> Dict[str, Tensor]:
        """loss terms of `training_step` (device scalars) with the gradients of all parameters left in `self.flat_g`."""
        S, _, A = batch["agent/valid"].shape
        if eps is None or use_prior is None:
            use_prior, eps = self.draw_noise(S, A)
        if new_seed and not (self.device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            self.new_dropout_seed()
        self.flat_g.zero_()
        kw = dict(return_buffers=return_buffers, drop_seed=self.drop_seed if self.dropout_p > 0.0 else None, drop_p=self.dropout_p,
                  stack_backward=self.stack_backward)
        n_split = min(self.n_split, S) if self.device.type == "cuda" and not return_buffers else 1
        if n_split <= 1:
            fn = Fn(self.ops)
            out = graph.training_forward(fn, self.params, batch, eps.to(self.device), use_prior, **kw)
            self.last_ops = fn.n_fwd
            fn.backward()
            return out
        # ---- scenes are independent: n_split sub-batches run as concurrent chains on side streams (inside a graph capture they
        # become parallel branches).  At 16 scenes a launch has 1024 rows and is bound by latency, not by work: two chains of 512
        # rows overlap almost perfectly.  The loss normalisers are batch-wide counts, so the chains meet once between forward
        # and backward; parameter gradients of all chains accumulate atomically in the flat buffer.
        cur = torch.cuda.current_stream(self.device)
        if self._streams is None or len(self._streams) < n_split:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(n_split)]
        eps = eps.to(self.device)
        bounds = [(i * S) // n_split for i in range(n_split + 1)]
        fns, parts = [], []
        for i in range(n_split):
            st = self._streams[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                lo, hi = bounds[i], bounds[i + 1]
                sub = {k: v[lo:hi] for k, v in batch.items()}
                fn = Fn(self.ops)
                parts.append(graph.training_forward(fn, self.params, sub, eps[lo:hi], use_prior, defer_loss=True,
                                                    first_drop_site=i * 1000000, **kw))
                fns.append(fn)
        for st in self._streams[:n_split]:
            cur.wait_stream(st)
        counts = parts[0]["counts"]
        for p_ in parts[1:]:
            counts = counts + p_["counts"]
        outs = []
        for i in range(n_split):
            st = self._streams[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(parts[i]["finish"](counts))
                fns[i].backward()
        for st in self._streams[:n_split]:
            cur.wait_stream(st)
        self.last_ops = sum(fn.n_fwd for fn in fns)
        self._keep = (outs, counts)  # tensors that crossed streams stay referenced until the next step
        return {k: sum(o[k] for o in outs) for k in outs[0]}

    # ---- whole-step CUDA graph -------------------------------------------------------------------------------------
    def capture(self, batch: Mapping[str, Tensor]) -> None:
        """Captures forward + backward of the step (~40 k kernel launches for 90 decode steps; eagerly the step is bound by
        the host issuing them) into CUDA graphs -- one per latent choice (prior / posterior rollout, :384-387), created on
        first use, sharing one memory pool -- for batches of this shape.  `replay(batch, ...)` then copies the batch into
        the graph's static input buffers and launches the graph.  Everything data-dependent in the step is decided on the
        device (masks), so one capture serves every batch of the same shape."""
        self._static_batch = {k: v.to(self.device).clone() for k, v in batch.items()}
        S, _, A = batch["agent/valid"].shape
        self._static_eps = torch.zeros(S, A, 16, device=self.device)
        self._graphs = {}
        self._pool = None

    def _graph(self, use_prior: bool):
        if use_prior not in self._graphs:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # eager run on a side stream first (allocator / lazy-initialisation warm-up)
                self.forward_backward(self._static_batch, self._static_eps, use_prior)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            n0 = self.ops.L.tb_launch_count()
            with torch.cuda.graph(g, pool=self._pool):
                out = self.forward_backward(self._static_batch, self._static_eps, use_prior)
            if self._pool is None:
                self._pool = g.pool()
            self._graphs[use_prior] = (g, out, int(self.ops.L.tb_launch_count() - n0))  # kernels of this library in the graph
        return self._graphs[use_prior]

    def replay(self, batch: Mapping[str, Tensor], eps: Optional[Tensor] = None, use_prior: Optional[bool] = None) -> Dict[str, Tensor]:
        """forward + backward through the captured graph (see `capture`); `batch` may live in (pinned) host memory."""
        S, _, A = batch["agent/valid"].shape
        if eps is None or use_prior is None:
            use_prior, eps = self.draw_noise(S, A)
        g, out, n_kernel = self._graph(bool(use_prior))
        for k, dst in self._static_batch.items():
            dst.copy_(batch[k], non_blocking=True)
        self._static_eps.copy_(eps, non_blocking=True)
        self.new_dropout_seed()
        g.replay()
        self.replayed_kernels += n_kernel
        return out

    def all_reduce_grads(self) -> None:
        """DDP's gradient averaging (reference: Lightning DDP, src/run.py:51-53) as ONE collective on the flat buffer."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self.ops.scale_(self.flat_g, 1.0 / dist.get_world_size())

    def optimizer_step(self) -> Tensor:
        """clip_grad_norm_(max_grad_norm) + Adam; returns the squared gradient norm (device scalar)."""
        self.n_step += 1
        sq = self.ops.grad_sq_norm(self.flat_g)
        self.ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.group_end, self.betas[0], self.betas[1], self.eps,
                           self.n_step, sq, self.max_grad_norm)
        return sq

    def training_step(self, batch: Mapping[str, Tensor], eps: Optional[Tensor] = None, use_prior: Optional[bool] = None,
                      graph: bool = False) -> Dict[str, Tensor]:
        out = self.replay(batch, eps, use_prior) if graph else self.forward_backward(batch, eps, use_prior)
        self.all_reduce_grads()
        out["grad_sq_norm"] = self.optimizer_step()
        return out

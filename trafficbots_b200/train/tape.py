"""Reverse-mode tape over the training primitives of `csrc/tb_train.cu`.

The reference trains through `torch.autograd` (Lightning backward of `training_step`, pl_modules/waymo_motion.py:356-418).
Here every differentiable operation is one of ~20 hand-written CUDA primitives (forward kernel + hand-derived backward
kernel, bound in `cuda_ops.CudaOps`); this module only records which primitive produced which buffer and replays the
backward kernels in reverse order.  torch tensors are used as device memory; no torch autograd is involved.

`Var` = a 2-D fp32 buffer [rows, cols] (+ its gradient buffer).  `Fn` = the recording front end used by `graph.py`.
The primitive back end (`ops`) is injected: `CudaOps` in the product (the default; raises without the CUDA library),
a torch restatement in the tests (`oracle/train_ops_oracle.py`), which is how the composition is validated on CPU against
gradients of the unmodified reference.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from torch import Tensor


class Var:
    __slots__ = ("data", "grad", "req", "grad_fixed", "grad_shared")

    def __init__(self, data: Tensor, req: bool = False, grad: Optional[Tensor] = None) -> None:
        self.data = data
        self.req = req  # a gradient is wanted for this buffer
        self.grad = grad
        self.grad_fixed = grad is not None  # grad is a view into a parent / flat parameter buffer: accumulate in place only
        self.grad_shared = False  # grad tensor may be referenced by another Var: never modify it in place

    @property
    def rows(self) -> int:
        return self.data.shape[0]

    @property
    def cols(self) -> int:
        return self.data.shape[1]

    def detach(self) -> "Var":
        return Var(self.data, False)


class StepStack:
    """Allocation hook of the ops back end for a chain of forward primitives that runs once per decode step: the k-th output
    buffer of step t is row block t of ONE buffer [n_step, rows, cols].  The backward of that chain then runs ONCE over the stacked
    buffers (`ReplayOps`) instead of once per step: its launches see n_step x 1024 rows instead of 1024."""

    def __init__(self, ops, n_step: int) -> None:
        self.ops, self.n_step, self.bufs, self.t, self.k = ops, n_step, [], 0, 0

    def begin(self, t: int) -> None:
        self.t, self.k = t, 0
        self.ops.alloc_hook = self._alloc

    def end(self) -> None:
        self.ops.alloc_hook = None
        assert self.k == len(self.bufs), "every step must allocate the same sequence of buffers"

    def _alloc(self, shape, dtype) -> Tensor:
        if self.k == len(self.bufs):
            assert self.t == 0, "allocation sequence changed between steps"
            self.ops.alloc_hook = None
            self.bufs.append(self.ops.empty((self.n_step,) + tuple(shape), dtype=dtype))
            self.ops.alloc_hook = self._alloc
        buf = self.bufs[self.k]
        assert tuple(buf.shape[1:]) == tuple(shape) and buf.dtype == dtype, (buf.shape, shape)
        self.k += 1
        return buf[self.t]


class ReplayOps:
    """ops back end for RECORDING a chain whose forward already ran step by step into a `StepStack`: every forward primitive
    returns the next stacked buffer(s) instead of launching; everything else (the backward kernels) is the real back end."""

    def __init__(self, ops, stack: StepStack) -> None:
        self._ops, self._bufs, self._i = ops, stack.bufs, 0

    def __getattr__(self, name):
        return getattr(self._ops, name)

    def _pop(self) -> Tensor:
        b = self._bufs[self._i]
        self._i += 1
        return b.flatten(0, 1)

    def done(self) -> bool:
        return self._i == len(self._bufs)

    def linear_fwd(self, *a, **k):
        return self._pop()

    add_mask_fwd = select_rows_fwd = cat2_fwd = gather_rows_fwd = dropout = linear_fwd

    def layernorm_fwd(self, *a, **k):
        y = self._pop()
        return y, self._pop()

    def attention_fwd(self, *a, **k):
        o, p, alive = self._pop(), self._pop(), self._pop()
        return o, (p, o), alive


class Fn:
    """records primitives on a tape; `backward()` replays the backward kernels in reverse."""

    def __init__(self, ops, record: bool = True) -> None:
        self.ops = ops
        self.nodes: List[Callable[[], None]] = []
        self.n_fwd = 0
        self.record = record  # False: forward only (a step of a `StepStack` chain, recorded later through `ReplayOps`)
        self.stack_t: Optional[int] = None  # step index of a `StepStack` chain: offsets the dropout element indices

    def _d(self, drop, numel: int):
        """dropout site of a launch that fills row block `stack_t` of a step-stacked buffer: element indices continue across steps."""
        if drop is None or self.stack_t is None:
            return drop
        return (drop[0], drop[1], drop[2], self.stack_t * numel)

    # ------------------------------------------------------------------ tape mechanics
    def _acc(self, v: Var, g: Optional[Tensor], owned: bool = True) -> None:
        if g is None or not v.req:
            return
        if v.grad is None:
            v.grad = g
            v.grad_shared = not owned
        elif v.grad_fixed:
            self.ops.add_(v.grad, g)
        elif v.grad_shared:
            v.grad = self.ops.add_mask_fwd(v.grad, g, None)
            v.grad_shared = False
        else:
            self.ops.add_(v.grad, g)

    def _push(self, out: Var, fn: Callable[[Tensor], None]) -> None:
        if not self.record:
            return

        def node() -> None:
            g = out.grad
            if g is None:
                return
            fn(g)
            if not out.grad_fixed:
                out.grad = None  # the gradient of an intermediate buffer is dead once its producer has consumed it
        self.nodes.append(node)

    def backward(self) -> None:
        nodes, self.nodes = self.nodes, []
        while nodes:
            nodes.pop()()

    def const(self, t: Tensor) -> Var:
        return Var(t, False)

    def row_slice(self, v: Var, lo: int, hi: int) -> Var:
        """rows [lo, hi) of a buffer as a Var of its own; the gradient is accumulated into the parent's (zero-initialised)."""
        if not v.req:
            return Var(v.data[lo:hi], False)
        if v.grad is None:
            v.grad = self.ops.zeros(tuple(v.data.shape), v.data)
            v.grad_shared = False
        elif v.grad_shared:
            v.grad = v.grad.clone()  # memory copy: the slices below are accumulated into in place
            v.grad_shared = False
        out = Var(v.data[lo:hi], True, v.grad[lo:hi])
        return out

    def cat_rows(self, vs: Sequence[Var]) -> Var:
        req = any(v.req for v in vs)
        out = Var(torch.cat([v.data for v in vs], 0), req)
        self.n_fwd += 1
        if req:
            def bw(g: Tensor) -> None:
                lo = 0
                for v in vs:
                    self._acc(v, g[lo:lo + v.rows], owned=False)
                    lo += v.rows
            self._push(out, bw)
        return out

    # ------------------------------------------------------------------ primitives
    def linear(self, x: Var, w: Var, b: Optional[Var], relu: bool = False, keep_lin: Optional[Tensor] = None,
               res: Optional[Var] = None, keep_out: Optional[Tensor] = None, drop=None) -> Var:
        """y = (dropout(relu(x W^T + b) * keep_lin[row]) + res) * keep_out[row]: a Linear with the dropout / residual / row-mask
        tail of a transformer sub-layer fused into its epilogue (one kernel instead of four).  drop = (seed, site, p) or None."""
        assert not (relu and (res is not None or keep_lin is not None or keep_out is not None))
        bd = None if b is None else b.data.view(-1)
        drop = self._d(drop, x.rows * w.data.shape[0])
        y = self.ops.linear_fwd(x.data, w.data, bd, relu, keep_lin, None if res is None else res.data, keep_out, drop)
        self.n_fwd += 1
        req = x.req or w.req or (res is not None and res.req)
        out = Var(y, req)
        if req:
            def bw(g: Tensor) -> None:
                if x.req or w.req:
                    dx = self.ops.linear_bwd(g, x.data, w.data, bd, y, relu, w.grad if w.req else None,
                                             b.grad.view(-1) if (b is not None and b.req) else None, x.req, keep_lin, keep_out, drop)
                    self._acc(x, dx)
                if res is not None and res.req:
                    d = self.ops.add_mask_bwd(g, keep_out)
                    self._acc(res, d, owned=d is not g)
            self._push(out, bw)
        return out

    def layernorm(self, x: Var, w: Var, b: Var, relu: bool = False, drop=None) -> Var:
        wd, bd = w.data.view(-1), b.data.view(-1)
        drop = self._d(drop, x.data.numel())
        y, stats = self.ops.layernorm_fwd(x.data, wd, bd, relu, drop)
        self.n_fwd += 1
        req = x.req or w.req
        out = Var(y, req)
        if req:
            def bw(g: Tensor) -> None:
                dx = self.ops.layernorm_bwd(g, x.data, wd, bd, stats, y, relu, w.grad.view(-1) if w.req else None,
                                            b.grad.view(-1) if b.req else None, drop)
                self._acc(x, dx)
            self._push(out, bw)
        return out

    def attention(self, q: Var, kv: Var, key_valid: Tensor, n_batch: int, n_src: int, n_tgt: int, eye: bool, drop=None,
                  kv_shared: bool = False):
        """q [B*S, D], kv [B*T, 2D] -> (o [B*S, D], alive [B*S] u8: 0 for rows without any admissible key); drop: dropout on the
        attention probabilities.
        kv_shared (backward of a step-stacked chain only): kv holds fewer batch elements than q, element b uses kv[b % n_kv]."""
        D = q.cols
        n_kv = kv.rows // n_tgt
        drop = self._d(drop, n_batch * 4 * n_src * n_tgt)
        o, p, alive = self.ops.attention_fwd(q.data.view(n_batch, n_src, D), kv.data.view(n_kv, n_tgt, 2 * D), key_valid, eye, drop)
        self.n_fwd += 1
        req = q.req or kv.req
        out = Var(o.view(n_batch * n_src, D), req)
        if req:
            def bw(g: Tensor) -> None:
                dq, dkv = self.ops.attention_bwd(g.view(n_batch, n_src, D), q.data.view(n_batch, n_src, D),
                                                 kv.data.view(n_kv, n_tgt, 2 * D), key_valid, eye, p, drop, kv_shared)
                self._acc(q, dq.view(n_batch * n_src, D))
                self._acc(kv, dkv.view(n_kv * n_tgt, 2 * D))
            self._push(out, bw)
        return out, alive.view(-1)

    def add_mask(self, a: Var, b: Optional[Var], keep: Optional[Tensor], keep_a: Optional[Tensor] = None) -> Var:
        """(a * keep_a[row] + b) * keep[row] (masks are row masks; a zero entry zeroes the row)."""
        y = self.ops.add_mask_fwd(a.data, None if b is None else b.data, keep, keep_a)
        self.n_fwd += 1
        req = a.req or (b is not None and b.req)
        out = Var(y, req)
        if req:
            def bw(g: Tensor) -> None:
                d = self.ops.add_mask_bwd(g, keep)
                if b is not None:
                    self._acc(b, d, owned=False if (d is g or a.req) else True)
                if a.req:
                    if keep_a is None:
                        self._acc(a, d, owned=not (d is g or (b is not None and b.req)))
                    else:
                        self._acc(a, self.ops.add_mask_bwd(g, keep, keep_a))
            self._push(out, bw)
        return out

    def dropout(self, x: Var, drop) -> Var:
        """elementwise dropout (nn.GRU's inter-layer dropout); drop None = identity."""
        if drop is None:
            return x
        drop = self._d(drop, x.data.numel())
        y = self.ops.dropout(x.data, drop)
        self.n_fwd += 1
        out = Var(y, x.req)
        if x.req:
            def bw(g: Tensor) -> None:
                self._acc(x, self.ops.dropout_bwd(g, drop))
            self._push(out, bw)
        return out

    def select_rows(self, mask: Tensor, a: Var, b: Var) -> Var:
        y = self.ops.select_rows_fwd(mask, a.data, b.data)
        self.n_fwd += 1
        req = a.req or b.req
        out = Var(y, req)
        if req:
            def bw(g: Tensor) -> None:
                da, db = self.ops.select_rows_bwd(g, mask)
                self._acc(a, da)
                self._acc(b, db)
            self._push(out, bw)
        return out

    def cat2(self, a: Var, b: Var) -> Var:
        y = self.ops.cat2_fwd(a.data, b.data)
        self.n_fwd += 1
        req = a.req or b.req
        out = Var(y, req)
        if req:
            ka = a.cols

            def bw(g: Tensor) -> None:
                da, db = self.ops.cat2_bwd(g, ka)
                self._acc(a, da)
                self._acc(b, db)
            self._push(out, bw)
        return out

    def gru_gates(self, gi: Var, gh: Var, h: Var) -> Var:
        y = self.ops.gru_gates_fwd(gi.data, gh.data, h.data)
        self.n_fwd += 1
        out = Var(y, True)

        def bw(g: Tensor) -> None:
            dgi, dgh, dh = self.ops.gru_gates_bwd(g, gi.data, gh.data, h.data)
            self._acc(gi, dgi)
            self._acc(gh, dgh)
            self._acc(h, dh)
        self._push(out, bw)
        return out

    def masked_max(self, x: Var, valid: Tensor, n_outer: int, n_red: int, n_inner: int, fill: float) -> Var:
        D = x.cols
        y, idx = self.ops.masked_max_fwd(x.data.view(n_outer, n_red, n_inner, D), valid.view(n_outer, n_red, n_inner), fill)
        self.n_fwd += 1
        out = Var(y.view(n_outer * n_inner, D), x.req)
        if x.req:
            def bw(g: Tensor) -> None:
                dx = self.ops.masked_max_bwd(g.view(n_outer, n_inner, D), idx, n_red)
                self._acc(x, dx.view(-1, D))
            self._push(out, bw)
        return out

    def gather_rows(self, x: Var, idx: Tensor) -> Var:
        y = self.ops.gather_rows_fwd(x.data, idx)
        self.n_fwd += 1
        out = Var(y, x.req)
        if x.req:
            n = x.rows

            def bw(g: Tensor) -> None:
                self._acc(x, self.ops.gather_rows_bwd(g, idx, n))
            self._push(out, bw)
        return out

    def pair_add(self, u: Var, v: Var, n_scene: int, n_pl: int, n_agent: int) -> Var:
        D = u.cols
        y = self.ops.pair_add_fwd(u.data.view(n_scene, n_pl, D), v.data.view(n_scene, n_agent, D))
        self.n_fwd += 1
        req = u.req or v.req
        out = Var(y.view(-1, D), req)
        if req:
            def bw(g: Tensor) -> None:
                du, dv = self.ops.pair_add_bwd(g.view(n_scene, n_agent, n_pl, D))
                self._acc(u, du.reshape(-1, D))
                self._acc(v, dv.reshape(-1, D))
            self._push(out, bw)
        return out

    def rsample(self, mean: Var, log_std: Var, eps: Tensor) -> Var:
        z = self.ops.rsample_fwd(mean.data, log_std.data.view(-1), eps)
        self.n_fwd += 1
        out = Var(z, True)

        def bw(g: Tensor) -> None:
            dm = self.ops.rsample_bwd(g, eps, log_std.data.view(-1), log_std.grad.view(-1))
            self._acc(mean, dm, owned=dm is not g)
        self._push(out, bw)
        return out

    def dynamics(self, state: Var, mean: Var, a_type: Tensor, valid: Tensor) -> Var:
        y = self.ops.dynamics_fwd(state.data, mean.data, a_type, valid)
        self.n_fwd += 1
        out = Var(y, True)

        def bw(g: Tensor) -> None:
            ds, dm = self.ops.dynamics_bwd(g, state.data, mean.data, a_type, valid)
            self._acc(state, ds)
            self._acc(mean, dm)
        self._push(out, bw)
        return out

    def reward(self, pred: Var, gt: Tensor, rv: Tensor) -> Var:
        r = self.ops.reward_fwd(pred.data, gt, rv)
        self.n_fwd += 1
        out = Var(r.view(-1, 1), True)

        def bw(g: Tensor) -> None:
            self._acc(pred, self.ops.reward_bwd(g.view(-1), pred.data, gt, rv))
        self._push(out, bw)
        return out

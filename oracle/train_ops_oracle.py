"""TEST INFRASTRUCTURE ONLY -- torch (CPU, fp32) restatement of the training primitives of `csrc/tb_train.cu`.

`trafficbots_b200/train/graph.py` composes the reference's `training_step` (pl_modules/waymo_motion.py:356-418) from a small
set of primitives (forward + hand-derived backward CUDA kernels, `trafficbots_b200/train/cuda_ops.py`).  This module offers
the same primitive interface with torch ops; every backward here is `torch.autograd.grad` of the forward, i.e. NOT the
hand-derived formula -- which makes it the checker of the CUDA backward kernels (tests/test_gpu_train.py) and lets the
composition itself be validated on CPU against gradients of the UNMODIFIED reference (tests/test_train_cpu.py,
fixtures from oracle/make_golden_train.py).  Only tests/ may import it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

N_HEAD = 4
DT = 0.1
MAX_ACC = (5.0, 7.0, 6.0)  # veh, ped, cyc (configs/model/traffic_bots.yaml:142-155; order utils/dynamics.py:23-27)
MAX_YAW_RATE = (1.5, 7.0, 3.0)


def _grads(outs, ins, gouts):
    ins_ = [t for t in ins if t is not None]
    g = torch.autograd.grad(outs, ins_, gouts, allow_unused=True)
    g = [torch.zeros_like(t) if gi is None else gi for gi, t in zip(g, ins_)]
    it = iter(g)
    return [None if t is None else next(it) for t in ins]


def _mix32(x: Tensor) -> Tensor:
    """lowbias32 on int64 tensors holding uint32 values (the hash of csrc/tb_train.cu::mix32)."""
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x7feb352d) & m
    x = x ^ (x >> 15)
    x = (x * 0x846ca68b) & m
    x = x ^ (x >> 16)
    return x


def drop_factor(drop, shape, device) -> Tensor:
    """dropout factor tensor of a call site: drop = (seed tensor [1] int32/int64, site id, p); element index = flat index."""
    if drop is None:
        return None
    seed, site, p = drop[:3]
    offset = int(drop[3]) if len(drop) > 3 else 0
    n = 1
    for d in shape:
        n *= d
    key = _mix32(torch.tensor([(site * 0x9E3779B9) & 0xFFFFFFFF], dtype=torch.int64, device=device) ^ (seed.to(torch.int64).to(device) & 0xFFFFFFFF))
    h = _mix32(((torch.arange(n, dtype=torch.int64, device=device) + offset) & 0xFFFFFFFF) ^ key)
    thresh = int(float(torch.tensor(p, dtype=torch.float32)) * 4294967296.0)  # p as the fp32 value the kernel receives
    scale = torch.tensor(1.0, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(p, dtype=torch.float32))
    return ((h >= thresh).to(torch.float32) * scale.to(device)).view(*shape)


def _req(*ts):
    return [None if t is None else t.detach().clone().requires_grad_(True) for t in ts]


class OracleOps:
    """same method names / argument meaning as `trafficbots_b200.train.cuda_ops.CudaOps`."""

    name = "oracle"

    def __init__(self) -> None:
        # tape.StepStack (the step-stacked backward): like CudaOps, the forward primitives a decode step runs put their outputs into
        # buffers obtained from `empty`, in the SAME order as the CUDA wrappers allocate them (see `_out`)
        self.alloc_hook = None

    def _out(self, *tensors):
        """outputs of a forward primitive: copied into hook-provided buffers while a StepStack is recording."""
        if self.alloc_hook is None:
            return tensors if len(tensors) > 1 else tensors[0]
        outs = []
        for t in tensors:
            buf = self.alloc_hook(tuple(t.shape), t.dtype)
            buf.copy_(t)
            outs.append(buf)
        return tuple(outs) if len(outs) > 1 else outs[0]

    # ---- memory helpers (plumbing) ----
    def empty(self, shape, like: Tensor = None, dtype=torch.float32) -> Tensor:
        if self.alloc_hook is not None:
            return self.alloc_hook(tuple(shape), dtype)
        return torch.empty(shape, dtype=dtype, device=None if like is None else like.device)

    def zeros(self, shape, like: Tensor, dtype=torch.float32) -> Tensor:
        return torch.zeros(shape, dtype=dtype, device=like.device)

    def add_(self, dst: Tensor, src: Tensor) -> None:
        dst += src

    def scale_(self, x: Tensor, alpha: float) -> None:
        x *= alpha

    def _pure(self, fwd, *a, **k):
        """a forward primitive evaluated for autograd inside a backward: never through the StepStack hook."""
        hook, self.alloc_hook = self.alloc_hook, None
        try:
            return fwd(*a, **k)
        finally:
            self.alloc_hook = hook

    # ---- Linear (+ReLU) : models/modules/mlp.py, attention.py in/out projections ----
    def linear_fwd(self, x: Tensor, w: Tensor, b: Optional[Tensor], relu: bool, keep_lin=None, res=None, keep_out=None, drop=None) -> Tensor:
        """y = (dropout(relu(x W^T + b) * keep_lin[row]) + res) * keep_out[row] -- the optional tail is the residual / row-mask
        epilogue of a transformer sub-layer (transformer.py:203,220,236-237; attention.py:144-146)."""
        y = F.linear(x, w, b)
        y = torch.relu(y) if relu else y
        if keep_lin is not None:
            y = y * keep_lin.to(y.dtype).unsqueeze(-1)
        if drop is not None:
            y = y * drop_factor(drop, y.shape, y.device)
        if res is not None:
            y = y + res
        if keep_out is not None:
            y = y * keep_out.to(y.dtype).unsqueeze(-1)
        return self._out(y)

    def linear_bwd(self, dy, x, w, b, y, relu, dw, db, need_dx: bool, keep_lin=None, keep_out=None, drop=None):
        """gradients of the Linear part (the residual's gradient is dy * keep_out, formed by the caller)."""
        x_, w_, b_ = _req(x, w, b)
        gx, gw, gb = _grads(self._pure(self.linear_fwd, x_, w_, b_, relu, keep_lin, None, keep_out, drop), [x_, w_, b_], dy)
        if dw is not None:
            dw += gw
        if b is not None and db is not None:
            db += gb
        return gx if need_dx else None

    # ---- LayerNorm (+ReLU), eps 1e-5 ----
    def layernorm_fwd(self, x, w, b, relu: bool, drop=None):
        y = F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
        mean = x.mean(-1)
        rstd = (x.var(-1, unbiased=False) + 1e-5).rsqrt()
        y = torch.relu(y) if relu else y
        if drop is not None:
            y = y * drop_factor(drop, y.shape, y.device)
        return self._out(y, torch.stack([mean, rstd], -1))

    def layernorm_bwd(self, dy, x, w, b, stats, y, relu, dw, db, drop=None):
        x_, w_, b_ = _req(x, w, b)
        gx, gw, gb = _grads(self._pure(self.layernorm_fwd, x_, w_, b_, relu, drop)[0], [x_, w_, b_], dy)
        dw += gw
        db += gb
        return gx

    # ---- multi-head cross attention core (models/modules/attention.py:89-141), no projections ----
    def attention_fwd(self, q: Tensor, kv: Tensor, key_valid: Tensor, eye: bool, drop=None):
        """q [B,S,D], kv [B,T,2D] (K | V), key_valid [B,T]; eye: query i may not attend key i (agent_interaction.py:57-59).
        Rows without any admissible key (attention.py:101-107,144-146): o = 0, p = 0, alive = 0.  Returns (o, p, alive)."""
        B, S, D = q.shape
        T = kv.shape[1]
        dh = D // N_HEAD
        k, v = kv[..., :D], kv[..., D:]
        invalid = (~key_valid.bool()).unsqueeze(1).expand(-1, S, -1)
        if eye:
            invalid = invalid | torch.eye(S, dtype=torch.bool, device=q.device)[None]
        dead = invalid.all(-1)
        qh = q.view(B, S, N_HEAD, dh).transpose(1, 2)
        kh = k.reshape(B, T, N_HEAD, dh).transpose(1, 2)
        vh = v.reshape(B, T, N_HEAD, dh).transpose(1, 2)
        logits = torch.matmul(qh, kh.transpose(-2, -1)).masked_fill((invalid & ~dead.unsqueeze(-1)).unsqueeze(1), float("-inf"))
        p = torch.softmax(logits / math.sqrt(dh), dim=-1).masked_fill(dead[:, None, :, None], 0.0)
        pd = p if drop is None else p * drop_factor(drop, p.shape, p.device)  # attention.py:131-132
        o = torch.matmul(pd, vh).transpose(1, 2).flatten(2, 3)
        return self._out(o, p, (~dead).to(torch.uint8))

    def attention_bwd(self, do, q, kv, key_valid, eye, p, drop=None, kv_shared: bool = False):
        q_, kv_ = _req(q, kv)
        if kv_shared:  # batch element b of q uses the keys of b % kv.shape[0]
            rep = q.shape[0] // kv.shape[0]
            o = self._pure(self.attention_fwd, q_, kv_.repeat(rep, 1, 1), key_valid.repeat(rep, 1), eye, drop)[0]
        else:
            o = self._pure(self.attention_fwd, q_, kv_, key_valid, eye, drop)[0]
        gq, gkv = _grads(o, [q_, kv_], do)
        return gq, gkv

    def dropout(self, x, drop):
        """x * dropout factor (nn.GRU inter-layer dropout); applied to dy it is its own backward."""
        return self._out(x * drop_factor(drop, x.shape, x.device))

    def dropout_bwd(self, g, drop):
        return g * drop_factor(drop, g.shape, g.device)

    # ---- elementwise glue ----
    def add_mask_fwd(self, a, b, keep, keep_a=None):
        """(a * keep_a[row] + b) * keep[row]"""
        y = a if keep_a is None else a * keep_a.to(a.dtype).unsqueeze(-1)
        y = y if b is None else y + b
        return self._out(y if keep is None else y * keep.to(y.dtype).unsqueeze(-1))

    def add_mask_bwd(self, dy, keep, keep_a=None):
        d = dy if keep is None else dy * keep.to(dy.dtype).unsqueeze(-1)
        return d if keep_a is None else d * keep_a.to(dy.dtype).unsqueeze(-1)

    def select_rows_fwd(self, mask, a, b):
        return self._out(torch.where(mask.bool().unsqueeze(-1), a, b))

    def select_rows_bwd(self, dy, mask):
        m = mask.bool().unsqueeze(-1)
        return dy.masked_fill(~m, 0.0), dy.masked_fill(m, 0.0)

    def cat2_fwd(self, a, b):
        return self._out(torch.cat([a, b], -1))

    def cat2_bwd(self, dy, ka: int):
        return dy[:, :ka].contiguous(), dy[:, ka:].contiguous()

    # ---- one GRU layer's gate math (torch nn.GRU equations; gate order r, z, n) ----
    def gru_gates_fwd(self, gi, gh, h):
        i_r, i_z, i_n = gi.chunk(3, -1)
        h_r, h_z, h_n = gh.chunk(3, -1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        return (1.0 - z) * n + z * h

    def gru_gates_bwd(self, dhn, gi, gh, h):
        gi_, gh_, h_ = _req(gi, gh, h)
        return tuple(_grads(self.gru_gates_fwd(gi_, gh_, h_), [gi_, gh_, h_], dhn))

    # ---- masked max over a middle dim: map_encoder.py:95-97,105-106 (fill -inf), agent_temporal.py:31-32,43-44 (fill -1e3) ----
    def masked_max_fwd(self, x, valid, fill: float):
        """x [O,R,I,D], valid [O,R,I] -> y [O,I,D] (0 where no valid element), idx [O,I,D] int32 (-1: no gradient)."""
        v = valid.bool()
        xm = x.masked_fill(~v.unsqueeze(-1), fill)
        y, idx = xm.max(dim=1)
        any_v = v.any(1)
        picked_valid = torch.gather(v.unsqueeze(-1).expand_as(x), 1, idx.unsqueeze(1)).squeeze(1)
        idx = idx.masked_fill(~(picked_valid & any_v.unsqueeze(-1)), -1)
        return y.masked_fill(~any_v.unsqueeze(-1), 0.0), idx.to(torch.int32)

    def masked_max_bwd(self, dy, idx, n_r: int):
        O, I, D = dy.shape
        dx = torch.zeros(O, n_r, I, D, dtype=dy.dtype, device=dy.device)
        ok = idx >= 0
        dx.scatter_(1, idx.clamp(min=0).long().unsqueeze(1), (dy * ok).unsqueeze(1))
        return dx

    def gather_rows_fwd(self, x, idx):
        return self._out(x[idx])

    def gather_rows_bwd(self, dy, idx, n_row: int):
        dx = torch.zeros(n_row, dy.shape[1], dtype=dy.dtype, device=dy.device)
        dx.index_add_(0, idx, dy)
        return dx

    # ---- destination predictor (models/goal_manager.py:294-307,328-333) ----
    def pair_add_fwd(self, u, v):
        """u [S,P,D], v [S,A,D] -> [S,A,P,D]: first Linear of the pair MLP on cat[map_feature, agent] is separable."""
        return u.unsqueeze(1) + v.unsqueeze(2)

    def pair_add_bwd(self, dy):
        return dy.sum(1), dy.sum(2)

    def dest_nll(self, logits, pair_ok, row_valid, gt, loss_rows, scale):
        """logits [S,A,P] raw pair-MLP outputs; pair_ok [S,A,P]: not masked to -inf (goal_manager.py:228-246,328-329);
        row_valid [S,A] = dist_valid (:330); rows that end up all -inf are reset to 0 (:331).  NLL of `gt` summed over
        `loss_rows` (metrics/training.py:138-147) and d(sum)/d(logits)."""
        lg = logits.detach().clone().requires_grad_(True)
        x = lg.masked_fill(~pair_ok.bool(), float("-inf"))
        x = x.masked_fill(~row_valid.bool().unsqueeze(-1), 0.0)
        x = x.masked_fill((x == float("-inf")).all(-1).unsqueeze(-1), 0.0)
        logp = x - x.logsumexp(-1, keepdim=True)
        nll = -logp.gather(-1, gt.unsqueeze(-1)).squeeze(-1).masked_fill(~loss_rows.bool(), 0.0)
        total = nll.sum()
        (g,) = torch.autograd.grad(total, lg)
        return total.detach().reshape(1), g * scale

    # ---- latent (models/modules/distributions.py:40-50; metrics/loss.py:74-77) ----
    def rsample_fwd(self, mean, log_std, eps):
        return mean + eps * log_std.exp()

    def rsample_bwd(self, dz, eps, log_std, dlog_std):
        dlog_std += (dz * eps * log_std.exp()).sum(0)
        return dz

    def kl_fwd_bwd(self, mu_q, ls_q, mu_p, ls_p, valid, free_nats: float, scale, dls_q, dls_p):
        """KL(N(mu_q, e^ls_q) || N(mu_p, e^ls_p)) summed over the latent dim, clamped from below at free_nats, summed over the
        valid rows.  Returns (sum [1], scale * d/dmu_q, scale * d/dmu_p); scale * d/dls_q, d/dls_p are ADDED to dls_q, dls_p."""
        a, b, c, d = _req(mu_q, ls_q, mu_p, ls_p)
        from torch.distributions import Independent, Normal, kl_divergence
        kl = kl_divergence(Independent(Normal(a, b.exp().expand_as(a)), 1), Independent(Normal(c, d.exp().expand_as(c)), 1))
        if free_nats > 0:
            kl = torch.max(kl, kl.new_full(kl.size(), free_nats))
        total = kl.masked_fill(~valid.bool(), 0.0).sum()
        g = _grads(total, [a, b, c, d], None)
        dls_q += g[1] * scale
        dls_p += g[3] * scale
        return total.detach().reshape(1), g[0] * scale, g[2] * scale

    def masked_sum(self, x, mask):
        """sum of x [M,N] over the entries with mask [M,N] != 0 -> [1]."""
        return (x * mask.to(x.dtype)).sum().reshape(1)

    def mask_scale(self, mask, scale):
        """mask [M] u8, scale [1] -> [M,1] fp32 = mask * scale."""
        return (mask.to(torch.float32) * scale).reshape(-1, 1)

    # ---- positional encoding (utils/pose_pe.py:57-62, utils/pos_emb.py) : no gradient (inputs are data / detached) ----
    def pose_pe(self, xy, yaw, f_xy, f_yaw):
        def emb(v, freqs):
            e = v.unsqueeze(-1) * freqs
            return torch.cat([torch.cos(e[..., ::2]), torch.sin(e[..., 1::2])], dim=-1)
        return torch.cat([emb(xy[..., 0], f_xy), emb(xy[..., 1], f_xy), emb(yaw, f_yaw)], dim=-1)

    def dir_to_yaw(self, d):
        return torch.atan2(d[..., 1], d[..., 0])

    # ---- Dynamics.update (utils/dynamics.py:74-119,187-228) ----
    def dynamics_fwd(self, state, mean, a_type, valid):
        """state [M,4], mean [M,2] (deterministic action), a_type [M,3], valid [M] -> pred_state [M,4]."""
        tf_ = a_type.to(state.dtype)
        th = torch.tanh(mean)
        keep = valid.to(state.dtype)
        a_acc = th[:, 0] * (tf_ * torch.tensor(MAX_ACC, device=state.device)).sum(-1) * keep
        a_yr = th[:, 1] * (tf_ * torch.tensor(MAX_YAW_RATE, device=state.device)).sum(-1) * keep
        v_t = state[:, 3] + 0.5 * DT * a_acc
        th_t = state[:, 2] + 0.5 * DT * a_yr
        delta = torch.stack([v_t * torch.cos(th_t), v_t * torch.sin(th_t), a_yr, a_acc], -1)
        has_type = a_type.bool().any(-1, keepdim=True).to(state.dtype)
        return (state + DT * delta) * has_type * keep.unsqueeze(-1)

    def dynamics_bwd(self, dpred, state, mean, a_type, valid):
        s_, m_ = _req(state, mean)
        return tuple(_grads(self.dynamics_fwd(s_, m_, a_type, valid), [s_, m_], dpred))

    # ---- DifferentiableReward.get, IL part (utils/rewards.py:117-131) ----
    def reward_fwd(self, pred, gt, rv):
        keep = rv.to(pred.dtype).unsqueeze(-1)
        gs, ps = gt * keep, pred * keep
        e_pos = F.smooth_l1_loss(gs[:, :2], ps[:, :2], reduction="none").sum(-1)
        e_rot = 0.5 * (1 - torch.cos(gs[:, 2] - ps[:, 2]))
        e_spd = F.smooth_l1_loss(gs[:, 3], ps[:, 3], reduction="none")
        return -(0.1 * e_pos + 10.0 * e_rot + 0.1 * e_spd) * keep.squeeze(-1)

    def reward_bwd(self, dr, pred, gt, rv):
        (p_,) = _req(pred)
        return _grads(self.reward_fwd(p_, gt, rv), [p_], dr)[0]

    # ---- non-differentiable simulation bookkeeping of one step (pl_modules/waymo_motion.py:311-320) ----
    def sim_flags(self, state, valid, gt_valid_t, boundary, dest_pos, dest_dir, dest_valid, dest_is_lane, dest_is_edge,
                  killed, dest_reached, goal_valid):
        """state [B,A,4] post-override, valid [B,A] post-override.  Always-on checks that feed back into the simulation:
        outside_map (utils/traffic_rule_checker.py:101-119) -> kill (utils/dynamics.py:151-167), dest_reached (:364-410) ->
        goal_valid (models/goal_manager.py:155-161).  Returns (valid', killed', dest_reached', goal_valid')."""
        valid, killed, dest_reached, goal_valid = valid.bool(), killed.bool(), dest_reached.bool(), goal_valid.bool()
        dest_dir = dest_dir / torch.norm(dest_dir, dim=-1, keepdim=True)  # traffic_rule_checker.py:93
        dest_thresh = torch.ones_like(state[..., 0]) * 50 * (1 - dest_is_edge.to(state.dtype) * 0.8)  # :95-98
        px, py = state[..., 0], state[..., 1]
        out_t = ((px > boundary[:, [1]]) | (px < boundary[:, [0]]) | (py > boundary[:, [3]]) | (py < boundary[:, [2]])) & valid
        dist = torch.norm(state[..., :2].unsqueeze(2) - dest_pos, dim=-1).masked_fill(~dest_valid.bool(), 1e4)
        pos_reached = (dist < dest_thresh.unsqueeze(-1)).any(-1)
        head = torch.stack([torch.cos(state[..., 2]), torch.sin(state[..., 2])], dim=-1)
        rot = (head.unsqueeze(2) * dest_dir).sum(-1).masked_fill(~dest_valid.bool(), 0.0)
        rot_reached = (rot > math.cos(math.radians(30))).any(-1)
        dest_t = ~dest_reached & valid & ((dest_is_lane.bool() & pos_reached & rot_reached) | (dest_is_edge.bool() & pos_reached))
        dest_reached = dest_reached | dest_t
        kill = out_t & ~gt_valid_t.bool() if gt_valid_t is not None else out_t
        killed = killed | kill
        valid = valid & ~kill
        goal_valid = goal_valid & valid & ~dest_reached
        u8 = torch.uint8
        return valid.to(u8), killed.to(u8), dest_reached.to(u8), goal_valid.to(u8)

    # ---- optimizer: torch.optim.Adam (defaults) on a flat buffer + clip_grad_norm_ ----
    def grad_sq_norm(self, g):
        return (g.double() ** 2).sum().float().reshape(1)

    def adam_step(self, p, g, m, v, lr_by_group, group_end, beta1, beta2, eps, step: int, sq_norm, max_norm):
        """torch.optim.Adam (defaults) on the flat buffer; g is first scaled by min(1, max_norm / (||g|| + 1e-6))
        (torch.nn.utils.clip_grad_norm_); parameter group i covers [group_end[i-1], group_end[i])."""
        coef = 1.0
        if sq_norm is not None and max_norm > 0:
            coef = torch.clamp(max_norm / (sq_norm.sqrt() + 1e-6), max=1.0)
        g = g * coef
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        lo = 0
        for gi, hi in enumerate(group_end.tolist()):
            p[lo:hi].addcdiv_(m[lo:hi], (v[lo:hi].sqrt() / math.sqrt(bc2)).add_(eps), value=-float(lr_by_group[gi]) / bc1)
            lo = hi

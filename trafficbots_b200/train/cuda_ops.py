"""`CudaOps` -- the training primitives of `csrc/tb_train.cu` behind the interface `tape.Fn` expects.

Thin ctypes wrappers: allocate the outputs (torch tensors as device memory), launch on the current CUDA stream, return.
There is no CPU implementation: constructing `CudaOps` without the built library or without a CUDA device raises `TbError`.
Reference operations each primitive stands for: see the header block in include/trafficbots_b200.h.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .. import _native as nt

F32 = torch.float32
U8 = torch.uint8


def _rows_ld(t: Tensor):
    """(data_ptr, leading dimension) of a 2-D fp32 tensor whose rows are dense (column stride 1); row slices of a larger
    parameter and column slices (`W[:, lo:hi]`) qualify."""
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != F32:
        raise nt.TbError(f"expected a row-strided 2-D fp32 tensor, got shape {tuple(t.shape)} stride {t.stride()} {t.dtype}")
    return t.data_ptr(), t.stride(0)


class CudaOps:
    name = "cuda"

    def __init__(self, device: Optional[torch.device] = None, check: bool = True) -> None:
        if not torch.cuda.is_available():
            raise nt.TbError("the training primitives have no CPU implementation: a CUDA device is required")
        self.L = nt.lib()
        self.check = check
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.alloc_hook = None  # tape.StepStack: outputs of the forward kernels land in slices of step-stacked buffers

    # ---- helpers ----
    def _st(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def _c(self, t: Tensor, dtype=F32) -> int:
        if self.check and (not t.is_cuda or not t.is_contiguous() or t.dtype != dtype):
            raise nt.TbError(f"expected a contiguous CUDA {dtype} tensor, got {tuple(t.shape)} {t.dtype} {t.device} "
                             f"contiguous={t.is_contiguous()}")
        return t.data_ptr()

    def _u8(self, t: Optional[Tensor]) -> Optional[int]:
        if t is None:
            return None
        if t.dtype == torch.bool:
            t = t.view(U8)
        return self._c(t, U8)

    def _drop(self, drop):
        """(seed pointer, site id, p, element-index offset) of a dropout call site; (NULL, 0, 0, 0) switches it off."""
        if drop is None:
            return None, 0, 0.0, 0
        seed, site, p = drop[:3]
        return self._c(seed, torch.int32), int(site) & 0xFFFFFFFF, float(p), int(drop[3]) if len(drop) > 3 else 0

    def dropout(self, x, drop):
        y = self.empty(tuple(x.shape))
        self._run(self.L.tb_tr_dropout(self._c(x), x.numel(), y.data_ptr(), *self._drop(drop), self._st()), "tb_tr_dropout")
        return y

    dropout_bwd = dropout  # the same mask applied to the gradient

    def _run(self, rc: int, what: str) -> None:
        if rc != nt.TB_OK:
            raise nt.TbError(f"{what} failed: {nt.STATUS.get(rc, rc)}")

    def empty(self, shape, like: Tensor = None, dtype=F32) -> Tensor:
        if self.alloc_hook is not None:
            return self.alloc_hook(tuple(shape), dtype)
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def zeros(self, shape, like: Tensor = None, dtype=F32) -> Tensor:
        return torch.zeros(shape, dtype=dtype, device=self.dev)  # cudaMemsetAsync

    def add_(self, dst: Tensor, src: Tensor) -> None:
        if dst.dim() == 1:
            dst, src = dst.view(1, -1), src.view(1, -1)
        if src.dim() != 2:
            src = src.reshape(dst.shape)
        pd, ld = _rows_ld(dst)
        ps, ls = _rows_ld(src)
        self._run(self.L.tb_tr_axpy(pd, ld, ps, ls, dst.shape[0], dst.shape[1], self._st()), "tb_tr_axpy")

    def scale_(self, x: Tensor, alpha: float) -> None:
        self._run(self.L.tb_tr_scale(self._c(x), x.numel(), float(alpha), self._st()), "tb_tr_scale")

    # ---- Linear ----
    def linear_fwd(self, x, w, b, relu, keep_lin=None, res=None, keep_out=None, drop=None):
        M, K = x.shape
        N = w.shape[0]
        pw, ldw = _rows_ld(w)
        y = self.empty((M, N))
        self._run(self.L.tb_tr_linear_fwd(self._c(x), M, K, pw, ldw, N, None if b is None else self._c(b), int(relu),
                                          self._u8(keep_lin), None if res is None else self._c(res), self._u8(keep_out), y.data_ptr(),
                                          *self._drop(drop), self._st()), "tb_tr_linear_fwd")
        return y

    def linear_bwd(self, dy, x, w, b, y, relu, dw, db, need_dx, keep_lin=None, keep_out=None, drop=None):
        M, K = x.shape
        N = w.shape[0]
        pw, ldw = _rows_ld(w)
        dx = self.empty((M, K)) if need_dx else None
        pdw, lddw = (None, 0) if dw is None else _rows_ld(dw)
        self._run(self.L.tb_tr_linear_bwd(self._c(dy), self._c(x), pw, ldw, self._c(y), int(relu), self._u8(keep_lin), self._u8(keep_out),
                                          M, K, N, None if dx is None else dx.data_ptr(), pdw, lddw, None if db is None else self._c(db),
                                          *self._drop(drop), self._st()), "tb_tr_linear_bwd")
        return dx

    # ---- LayerNorm ----
    def layernorm_fwd(self, x, w, b, relu, drop=None):
        M, D = x.shape
        y = self.empty((M, D))
        stats = self.empty((M, 2))
        self._run(self.L.tb_tr_layernorm_fwd(self._c(x), self._c(w), self._c(b), int(relu), M, D, y.data_ptr(), stats.data_ptr(),
                                             *self._drop(drop), self._st()), "tb_tr_layernorm_fwd")
        return y, stats

    def layernorm_bwd(self, dy, x, w, b, stats, y, relu, dw, db, drop=None):
        M, D = x.shape
        dx = self.empty((M, D))
        self._run(self.L.tb_tr_layernorm_bwd(self._c(dy), self._c(x), self._c(w), self._c(stats), self._c(y), int(relu), M, D,
                                             dx.data_ptr(), None if dw is None else self._c(dw), None if db is None else self._c(db),
                                             *self._drop(drop), self._st()), "tb_tr_layernorm_bwd")
        return dx

    # ---- attention ----
    def attention_fwd(self, q, kv, key_valid, eye, drop=None):
        B, S, D = q.shape
        T = kv.shape[1]
        o = self.empty((B, S, D))
        p = self.empty((B, 4, S, T))
        alive = self.empty((B, S), dtype=U8)
        self._run(self.L.tb_tr_attention_fwd(self._c(q), self._c(kv), self._u8(key_valid), int(eye), B, S, T, o.data_ptr(),
                                             p.data_ptr(), alive.data_ptr(), *self._drop(drop), self._st()), "tb_tr_attention_fwd")
        return o, (p, o), alive

    def attention_bwd(self, do, q, kv, key_valid, eye, p, drop=None, kv_shared: bool = False):
        """kv_shared: q / do / p hold B batch elements, kv only kv.shape[0] (element b uses the keys of b % kv.shape[0])."""
        p, o = p
        B, S, D = q.shape
        Bk, T = kv.shape[0], kv.shape[1]
        dq = self.zeros((B, S, D))
        dkv = self.zeros((Bk, T, 2 * D)) if kv_shared else self.empty((B, T, 2 * D))
        self._run(self.L.tb_tr_attention_bwd(self._c(do), self._c(q), self._c(kv), self._c(p), self._c(o), B, S, T,
                                             Bk if kv_shared else 0, dq.data_ptr(),
                                             dkv.data_ptr(), *self._drop(drop), self._st()), "tb_tr_attention_bwd")
        return dq, dkv

    # ---- glue ----
    def add_mask_fwd(self, a, b, keep, keep_a=None):
        M, N = a.shape
        y = self.empty((M, N))
        self._run(self.L.tb_tr_add_mask(self._c(a), self._u8(keep_a), None if b is None else self._c(b), self._u8(keep), M, N,
                                        y.data_ptr(), self._st()), "tb_tr_add_mask")
        return y

    def add_mask_bwd(self, dy, keep, keep_a=None):
        if keep is None and keep_a is None:
            return dy
        return self.add_mask_fwd(dy, None, keep, keep_a)

    def select_rows_fwd(self, mask, a, b):
        M, N = a.shape
        y = self.empty((M, N))
        self._run(self.L.tb_tr_select_rows(self._u8(mask), self._c(a), self._c(b), M, N, y.data_ptr(), self._st()),
                  "tb_tr_select_rows")
        return y

    def select_rows_bwd(self, dy, mask):
        M, N = dy.shape
        da, db = self.empty((M, N)), self.empty((M, N))
        self._run(self.L.tb_tr_select_rows_bwd(self._u8(mask), self._c(dy), M, N, da.data_ptr(), db.data_ptr(), self._st()),
                  "tb_tr_select_rows_bwd")
        return da, db

    def cat2_fwd(self, a, b):
        M, ka = a.shape
        kb = b.shape[1]
        y = self.empty((M, ka + kb))
        self._run(self.L.tb_tr_cat2(self._c(a), ka, self._c(b), kb, M, y.data_ptr(), self._st()), "tb_tr_cat2")
        return y

    def cat2_bwd(self, dy, ka):
        M, n = dy.shape
        da, db = self.empty((M, ka)), self.empty((M, n - ka))
        self._run(self.L.tb_tr_cat2_bwd(self._c(dy), ka, n - ka, M, da.data_ptr(), db.data_ptr(), self._st()), "tb_tr_cat2_bwd")
        return da, db

    def gru_gates_fwd(self, gi, gh, h):
        M = h.shape[0]
        hn = self.empty((M, h.shape[1]))
        self._run(self.L.tb_tr_gru_gates_fwd(self._c(gi), self._c(gh), self._c(h), M, hn.data_ptr(), self._st()), "tb_tr_gru_gates_fwd")
        return hn

    def gru_gates_bwd(self, dhn, gi, gh, h):
        M, D = h.shape
        dgi, dgh, dh = self.empty((M, 3 * D)), self.empty((M, 3 * D)), self.empty((M, D))
        self._run(self.L.tb_tr_gru_gates_bwd(self._c(dhn), self._c(gi), self._c(gh), self._c(h), M, dgi.data_ptr(), dgh.data_ptr(),
                                             dh.data_ptr(), self._st()), "tb_tr_gru_gates_bwd")
        return dgi, dgh, dh

    def masked_max_fwd(self, x, valid, fill):
        O, R, I, D = x.shape
        y = self.empty((O, I, D))
        idx = self.empty((O, I, D), dtype=torch.int32)
        self._run(self.L.tb_tr_masked_max_fwd(self._c(x), self._u8(valid), O, R, I, D, float(fill), y.data_ptr(), idx.data_ptr(),
                                              self._st()), "tb_tr_masked_max_fwd")
        return y, idx

    def masked_max_bwd(self, dy, idx, n_r):
        O, I, D = dy.shape
        dx = self.empty((O, n_r, I, D))
        self._run(self.L.tb_tr_masked_max_bwd(self._c(dy), self._c(idx, torch.int32), O, n_r, I, D, dx.data_ptr(), self._st()),
                  "tb_tr_masked_max_bwd")
        return dx

    def gather_rows_fwd(self, x, idx):
        M, D = idx.shape[0], x.shape[1]
        y = self.empty((M, D))
        self._run(self.L.tb_tr_gather_rows(self._c(x), self._c(idx, torch.int64), M, D, y.data_ptr(), self._st()), "tb_tr_gather_rows")
        return y

    def gather_rows_bwd(self, dy, idx, n_row):
        M, D = dy.shape
        dx = self.zeros((n_row, D))
        self._run(self.L.tb_tr_scatter_add_rows(self._c(dy), self._c(idx, torch.int64), M, D, dx.data_ptr(), self._st()),
                  "tb_tr_scatter_add_rows")
        return dx

    def pair_add_fwd(self, u, v):
        S, P, D = u.shape
        A = v.shape[1]
        y = self.empty((S, A, P, D))
        self._run(self.L.tb_tr_pair_add(self._c(u), self._c(v), S, P, A, y.data_ptr(), self._st()), "tb_tr_pair_add")
        return y

    def pair_add_bwd(self, dy):
        S, A, P, D = dy.shape
        du, dv = self.empty((S, P, D)), self.zeros((S, A, D))
        self._run(self.L.tb_tr_pair_add_bwd(self._c(dy), S, P, A, du.data_ptr(), dv.data_ptr(), self._st()), "tb_tr_pair_add_bwd")
        return du, dv

    def dest_nll(self, logits, pair_ok, row_valid, gt, loss_rows, scale):
        S, A, P = logits.shape
        total = self.zeros((1,))
        dlogits = self.empty((S, A, P))
        self._run(self.L.tb_tr_dest_nll(self._c(logits), self._u8(pair_ok), self._u8(row_valid), self._c(gt, torch.int64),
                                        self._u8(loss_rows), self._c(scale), S * A, P, total.data_ptr(), dlogits.data_ptr(),
                                        self._st()), "tb_tr_dest_nll")
        return total, dlogits

    def rsample_fwd(self, mean, log_std, eps):
        M, E = mean.shape
        z = self.empty((M, E))
        self._run(self.L.tb_tr_rsample(self._c(mean), self._c(log_std), self._c(eps), M, E, z.data_ptr(), self._st()), "tb_tr_rsample")
        return z

    def rsample_bwd(self, dz, eps, log_std, dlog_std):
        M, E = dz.shape
        self._run(self.L.tb_tr_rsample_bwd(self._c(dz), self._c(eps), self._c(log_std), M, E, self._c(dlog_std), self._st()),
                  "tb_tr_rsample_bwd")
        return dz

    def kl_fwd_bwd(self, mu_q, ls_q, mu_p, ls_p, valid, free_nats, scale, dls_q, dls_p):
        M, E = mu_q.shape
        total = self.zeros((1,))
        dmq, dmp = self.empty((M, E)), self.empty((M, E))
        self._run(self.L.tb_tr_kl(self._c(mu_q), self._c(ls_q), self._c(mu_p), self._c(ls_p), self._u8(valid), float(free_nats),
                                  self._c(scale), M, E, total.data_ptr(), dmq.data_ptr(), dmp.data_ptr(), self._c(dls_q),
                                  self._c(dls_p), self._st()), "tb_tr_kl")
        return total, dmq, dmp

    def masked_sum(self, x, mask):
        out = self.zeros((1,))
        self._run(self.L.tb_tr_masked_sum(self._c(x), self._u8(mask), x.numel(), out.data_ptr(), self._st()), "tb_tr_masked_sum")
        return out

    def mask_scale(self, mask, scale):
        n = mask.numel()
        out = self.empty((n, 1))
        self._run(self.L.tb_tr_mask_scale(self._u8(mask), self._c(scale), n, out.data_ptr(), self._st()), "tb_tr_mask_scale")
        return out

    # ---- data-side encodings (no gradient) ----
    def pose_pe(self, xy, yaw, f_xy, f_yaw):
        M = yaw.numel()
        n_xy, n_yaw = f_xy.numel(), f_yaw.numel()
        pe = self.empty((M, 2 * n_xy + n_yaw))
        self._run(self.L.tb_tr_pose_pe(self._c(xy.contiguous()), self._c(yaw.contiguous()), self._c(f_xy), n_xy, self._c(f_yaw), n_yaw, M,
                                       pe.data_ptr(), self._st()), "tb_tr_pose_pe")
        return pe

    def dir_to_yaw(self, d):
        d = d.contiguous()
        M = d.shape[0]
        yaw = self.empty((M,))
        self._run(self.L.tb_tr_dir_to_yaw(self._c(d), M, yaw.data_ptr(), self._st()), "tb_tr_dir_to_yaw")
        return yaw

    # ---- simulation ----
    def dynamics_fwd(self, state, mean, a_type, valid):
        M = state.shape[0]
        pred = self.empty((M, 4))
        self._run(self.L.tb_tr_dynamics(self._c(state), self._c(mean), self._u8(a_type), self._u8(valid), M, pred.data_ptr(), None,
                                        None, None, self._st()), "tb_tr_dynamics")
        return pred

    def dynamics_bwd(self, dpred, state, mean, a_type, valid):
        M = state.shape[0]
        ds, dm = self.empty((M, 4)), self.empty((M, 2))
        self._run(self.L.tb_tr_dynamics(self._c(state), self._c(mean), self._u8(a_type), self._u8(valid), M, None, self._c(dpred),
                                        ds.data_ptr(), dm.data_ptr(), self._st()), "tb_tr_dynamics (bwd)")
        return ds, dm

    def reward_fwd(self, pred, gt, rv):
        M = pred.shape[0]
        r = self.empty((M,))
        self._run(self.L.tb_tr_reward(self._c(pred), self._c(gt), self._u8(rv), M, r.data_ptr(), None, None, self._st()), "tb_tr_reward")
        return r

    def reward_bwd(self, dr, pred, gt, rv):
        M = pred.shape[0]
        dp = self.empty((M, 4))
        self._run(self.L.tb_tr_reward(self._c(pred), self._c(gt), self._u8(rv), M, None, self._c(dr), dp.data_ptr(), self._st()),
                  "tb_tr_reward (bwd)")
        return dp

    def sim_flags(self, state, valid, gt_valid_t, boundary, dest_pos, dest_dir, dest_valid, dest_is_lane, dest_is_edge, killed,
                  dest_reached, goal_valid):
        B, A = valid.shape
        o = [self.empty((B, A), dtype=U8) for _ in range(4)]
        self._run(self.L.tb_tr_sim_flags(self._c(state.contiguous()), self._u8(valid), self._u8(gt_valid_t), self._c(boundary),
                                         self._c(dest_pos), self._c(dest_dir), self._u8(dest_valid), self._u8(dest_is_lane),
                                         self._u8(dest_is_edge), self._u8(killed), self._u8(dest_reached), self._u8(goal_valid), B, A,
                                         o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), self._st()),
                  "tb_tr_sim_flags")
        return tuple(o)

    # ---- optimizer ----
    def grad_sq_norm(self, g):
        out = self.zeros((1,))
        self._run(self.L.tb_tr_sq_norm(self._c(g), g.numel(), out.data_ptr(), self._st()), "tb_tr_sq_norm")
        return out

    def adam_step(self, p, g, m, v, lr_by_group, group_end, beta1, beta2, eps, step, sq_norm, max_norm):
        self._run(self.L.tb_tr_adam_step(self._c(p), self._c(g), self._c(m), self._c(v), p.numel(), self._c(lr_by_group),
                                         self._c(group_end, torch.int32), group_end.numel(), beta1, beta2, eps, int(step),
                                         None if sq_norm is None else self._c(sq_norm), float(max_norm), self._st()),
                  "tb_tr_adam_step")

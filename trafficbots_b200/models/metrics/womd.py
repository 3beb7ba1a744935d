"""`WOMDMetrics` -- mirror of the packing half of the reference metric (`src/models/metrics/womd.py:14-164`): `update(batch,
pred_traj, pred_score)` builds the six tensors the Waymo motion-metrics op consumes.  The reference fills them with a
per-scene Python loop (:124-138) and lets torchmetrics all-gather six list states (`dist_sync_on_step=True`, :23,44-49); here
one CTA per scene writes ONE fixed-size record per scene (`tb_womd_pack`), the six tensors are strided views into the record
buffer, and the reduction over the GPUs is a single NCCL all-gather of `[n_scene, record_bytes]` on a side stream
(`gather()`), which never blocks the compute streams.

Evaluating the metrics themselves needs `waymo_open_dataset`'s TensorFlow op (`py_metrics_ops.motion_metrics`, :196-206),
which is outside the hot path: `compute()` returns the op's inputs exactly like the reference's `compute()`."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from ... import _native as nt
from ...config import UnsupportedConfig

STATES = ("prediction_trajectory", "prediction_score", "ground_truth_trajectory", "ground_truth_is_valid",
          "prediction_ground_truth_indices_mask", "object_type")


class WOMDMetrics:
    def __init__(self, prefix: str = "val", step_gt: int = 90, step_current: int = 10, interactive_challenge: bool = False) -> None:
        if interactive_challenge:
            raise UnsupportedConfig("WOMDMetrics: interactive_challenge=True is not implemented")
        self.prefix, self.step_gt, self.step_current = prefix, step_gt, step_current
        self.track_future_samples = step_gt - step_current
        self.m_joint, self.n_pred = 8, 1
        self.records: List[Tensor] = []  # device: one uint8 [n_scene, record_bytes] buffer per update() since the last reset()
        self.ops_inputs_cpu: List[Tensor] = []  # host: what aggregate_on_cpu() collected over the epoch
        self._shape = None  # (n_agent, K)
        self.overflow: Optional[Tensor] = None
        self._comm: Optional[torch.cuda.Stream] = None

    # ------------------------------------------------------------------------------------------------
    def record_layout(self, n_agent: int, K: int):
        off = (C.c_int64 * 6)()
        nbytes = nt.lib().tb_womd_record_bytes(n_agent, K, self.step_gt, self.step_current, self.m_joint, off)
        return int(nbytes), [int(x) for x in off]

    def views(self, rec: Tensor, n_agent: int, K: int) -> Dict[str, Tensor]:
        """the six tensors of the reference (womd.py:112-122) as strided views into a `[n_scene, record_bytes]` uint8 buffer."""
        nbytes, off = self.record_layout(n_agent, K)
        S = rec.shape[0]
        n_ds, n_gt = self.track_future_samples // 5, self.step_gt + 1
        f32 = rec.view(torch.float32)  # [S, nbytes / 4]
        row = nbytes // 4

        def fview(o, shape):
            st, acc = [], 1
            for d in reversed(shape):
                st.append(acc)
                acc *= d
            return torch.as_strided(f32, (S, *shape), (row, *reversed(st)), o // 4)

        def bview(o, shape):
            st, acc = [], 1
            for d in reversed(shape):
                st.append(acc)
                acc *= d
            return torch.as_strided(rec, (S, *shape), (nbytes, *reversed(st)), o).view(torch.bool)

        return {"prediction_trajectory": fview(off[0], (self.m_joint, K, self.n_pred, n_ds, 2)),
                "prediction_score": fview(off[1], (self.m_joint, K)),
                "ground_truth_trajectory": fview(off[2], (n_agent, n_gt, 7)),
                "ground_truth_is_valid": bview(off[3], (n_agent, n_gt)),
                "prediction_ground_truth_indices_mask": bview(off[4], (self.m_joint, self.n_pred)),
                "object_type": fview(off[5], (n_agent,))}

    def update(self, batch: Mapping[str, Tensor], pred_traj: Tensor, pred_score: Optional[Tensor] = None) -> Tensor:
        """batch: `agent/{role,valid,pos,size,yaw_bbox,vel,type}`; pred_traj [S, Tf, A, K, 2] (future steps
        step_current+1 .. step_gt, i.e. `pred_dict["waymo_trajs"]`); pred_score [S,A,K] normalised or None.  Returns the
        record buffer of this batch (also appended to `self.records`)."""
        S, Tf, A, K, _ = pred_traj.shape
        dev = pred_traj.device
        nbytes, off = self.record_layout(A, K)
        rec = torch.empty(S, nbytes, dtype=torch.uint8, device=dev)
        if self.overflow is None:
            self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        p = nt.dev_ptr
        Tg = batch["agent/valid"].shape[1]
        win = nt.TbWomdIn(p(batch["agent/role"], "u8", (S, A, 3), "agent/role"), p(batch["agent/valid"], "u8", (S, Tg, A), "agent/valid"),
                          p(batch["agent/pos"], "f32", (S, Tg, A, 2), "agent/pos"), p(batch["agent/size"], "f32", (S, A, 3), "agent/size"),
                          p(batch["agent/yaw_bbox"], "f32", (S, Tg, A, 1), "agent/yaw_bbox"),
                          p(batch["agent/vel"], "f32", (S, Tg, A, 2), "agent/vel"), p(batch["agent/type"], "u8", (S, A, 3), "agent/type"),
                          p(pred_traj.contiguous(), "f32", (S, Tf, A, K, 2), "pred_traj"),
                          p(pred_score.contiguous(), "f32", (S, A, K), "pred_score") if pred_score is not None else None,
                          A, K, Tf, Tg, self.step_gt, self.step_current, self.m_joint)
        base = rec.data_ptr()
        wout = nt.TbWomdOut(base + off[0], base + off[1], base + off[2], base + off[3], base + off[4], base + off[5], nbytes,
                            self.overflow.data_ptr())
        with torch.cuda.device(dev):
            nt.check(nt.lib().tb_womd_pack(S, C.byref(win), C.byref(wout), nt.current_stream_ptr()), "tb_womd_pack")
        self.records.append(rec)
        self._shape = (A, K)
        return rec

    __call__ = update

    # ------------------------------------------------------------------------------------------------
    def gather(self, rec: Tensor, group=None, side_stream: bool = True) -> Tensor:
        """the metrics reduction of one batch: ONE all-gather of the packed records, `[world * n_scene, record_bytes]`, issued
        on a dedicated side stream behind the current stream's position (the compute streams never wait for NCCL)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return rec
        world = dist.get_world_size(group)
        out = torch.empty(world * rec.shape[0], rec.shape[1], dtype=torch.uint8, device=rec.device)
        if not side_stream or not rec.is_cuda:
            dist.all_gather_into_tensor(out, rec, group=group)
            return out
        if self._comm is None:
            self._comm = torch.cuda.Stream(rec.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(rec.device))
        self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            rec.record_stream(self._comm)
            out.record_stream(self._comm)
            dist.all_gather_into_tensor(out, rec, group=group)
        return out

    def wait_gathers(self) -> None:
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)

    def aggregate_on_cpu(self, records: Optional[Tensor] = None) -> None:
        """reference womd.py:165-175: after every step the (rank-synchronised, `dist_sync_on_step=True`) states are moved to
        the host and kept there until the end of the epoch.  Here: ONE all-gather of the packed records (if distributed), one
        device -> host copy; the device buffer can then be dropped with `reset()` like in the reference's loop
        (waymo_motion.py:655-660), so an epoch does not accumulate device memory."""
        if records is None:
            if not self.records:
                return
            records = self.records[-1]
        if records.is_cuda:
            records = self.gather(records, side_stream=False)
        self.ops_inputs_cpu.append(records.cpu())

    def compute(self) -> Dict[str, List[Tensor]]:
        """the motion-metrics op inputs, as the reference's `compute()` returns them (:153-163): one entry per update()."""
        A, K = self._shape
        out: Dict[str, List[Tensor]] = {k: [] for k in STATES}
        for rec in (self.records if self.records else self.ops_inputs_cpu):
            for k, v in self.views(rec, A, K).items():
                out[k].append(v)
        return out

    def reset(self) -> None:
        """drops the per-step device records (the reference resets its torchmetrics states after `aggregate_on_cpu`,
        waymo_motion.py:659-660); what was aggregated on the host stays until `clear()`."""
        self.records = []

    def clear(self) -> None:
        self.records = []
        self.ops_inputs_cpu = []

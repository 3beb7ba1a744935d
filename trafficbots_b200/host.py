"""Host-side tensor plumbing around the CUDA hot path (torch ops on small bool / index tensors, no arithmetic of
the model): the teacher-forcing / spawn mask, per-mode sampling bookkeeping and batch movement."""
from __future__ import annotations

from typing import Dict, Mapping

import torch
from torch import Tensor


def teacher_forcing_mask(valid: Tensor, step_spawn_agent: int, step_warm_start: int) -> Tensor:
    """`TeacherForcing.get` with the (default-off) schedules (reference utils/teacher_forcing.py:43-56).
    valid [S,T,A] bool -> mask [S,T,A] bool: frame 0, spawns (invalid -> valid transitions) up to
    `step_spawn_agent`, and every valid agent during the warm start `0..step_warm_start`."""
    m = torch.zeros_like(valid)
    m[:, 0] |= valid[:, 0]
    if step_spawn_agent > 0:
        spawn = (~valid[:, :-1]) & valid[:, 1:]
        spawn[:, step_spawn_agent:] = False
        m[:, 1:] |= spawn
    if step_warm_start >= 0:
        m[:, : step_warm_start + 1] |= valid[:, : step_warm_start + 1]
    return m


def batch_to_device(batch: Mapping[str, Tensor], device, non_blocking: bool = True) -> Dict[str, Tensor]:
    return {k: v.to(device, non_blocking=non_blocking) for k, v in batch.items()}


def pin_batch(batch: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    return {k: v.pin_memory() for k, v in batch.items()}


class SceneStager:
    """Double-buffered staging of pinned host batches and results on a side stream: the copy of batch i+1 and the read-back
    of step i-1's results overlap the kernels of step i (the `DataLoader(pin_memory=True)` + non-blocking transfer idiom of
    the reference's Lightning loop, `data_h5_womd.py:21-55` / `run.py:51-53`).  `submit` -> `get` hand a batch over with an
    event; `read_back` queues device -> pinned-host copies behind the compute stream's current position."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def submit(self, host_batch: Mapping[str, Tensor]) -> None:
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def get(self) -> Dict[str, Tensor]:
        if self._pending is None:
            raise RuntimeError("SceneStager.get() without a submitted batch")
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev.values():
            t.record_stream(cur)
        return dev

    def read_back(self, pairs) -> None:
        """pairs: iterable of (pinned host tensor, device tensor)."""
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            for dst, src in pairs:
                src.record_stream(self.stream)
                dst.copy_(src, non_blocking=True)

    def join(self) -> None:
        """make the compute stream wait for everything queued on the staging stream."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

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

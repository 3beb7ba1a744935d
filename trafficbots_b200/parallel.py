"""Scene sharding over the GPUs of one box: one process per GPU, scenes are independent, so the data
path has NO collective -- rank r owns a contiguous block of scenes (the K joint futures of a scene stay together
because they share the scene's map / traffic-light K|V caches).  The only exchange is the metrics reduction: one
fixed-layout all-gather of a packed per-scene buffer (replaces torchmetrics' per-state all-gathers, reference
`src/models/metrics/womd.py:23,44-49`, `metrics/logging.py:15-18`)."""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

METRIC_FIELDS = ("n_agent_valid", "min_ade", "min_fde", "ade_mode0", "outside_map", "goal_reached", "dest_reached",
                 "mean_reward")


def scene_shard(n_scene: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the scenes owned by `rank`: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(n_scene, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch: Mapping[str, Tensor], rank: int, world: int) -> Dict[str, Tensor]:
    n = next(iter(batch.values())).shape[0]
    b, e = scene_shard(n, rank, world)
    return {k: v[b:e] for k, v in batch.items()}


def pack_scene_metrics(preds: Tensor, valid: Tensor, violations: Mapping[str, Tensor], rewards: Tensor, gt_pos: Tensor,
                       gt_valid: Tensor, step_current: int = 10) -> Tensor:
    """per-scene summary of a K-mode rollout, [n_scene, len(METRIC_FIELDS)] fp32.
    preds [S,A,K,T,4], valid [S,A,K,T], violations[k] [S,A,K,T], rewards [S,A,K,T]; gt_pos [S,T+1,A,2], gt_valid [S,T+1,A].
    Displacement errors use the future steps (t > step_current) where both prediction and GT are valid."""
    S, A, K, T = valid.shape
    g_pos = gt_pos[:, 1:T + 1].transpose(1, 2).unsqueeze(2)  # [S,A,1,T,2]
    g_val = gt_valid[:, 1:T + 1].transpose(1, 2).unsqueeze(2)  # [S,A,1,T]
    fut = torch.zeros(T, dtype=torch.bool, device=valid.device)
    fut[step_current:] = True
    m = valid & g_val & fut
    err = (preds[..., :2] - g_pos).norm(dim=-1) * m
    n = m.sum(-1).clamp(min=1)
    ade = err.sum(-1) / n  # [S,A,K]
    last = (m.float() * torch.arange(1, T + 1, device=valid.device)).argmax(-1, keepdim=True)
    fde = err.gather(-1, last).squeeze(-1)
    has = m.any(-1)  # [S,A,K]
    agent_has = has.any(-1)
    big = torch.finfo(ade.dtype).max
    min_ade = torch.where(has, ade, torch.full_like(ade, big)).amin(-1)
    min_fde = torch.where(has, fde, torch.full_like(fde, big)).amin(-1)
    n_ag = agent_has.sum(-1).clamp(min=1).float()
    out = torch.stack([
        agent_has.sum(-1).float(),
        (min_ade * agent_has).sum(-1) / n_ag,
        (min_fde * agent_has).sum(-1) / n_ag,
        (ade[:, :, 0] * has[:, :, 0]).sum(-1) / has[:, :, 0].sum(-1).clamp(min=1),
        violations["outside_map"][..., -1].float().mean(dim=(1, 2)),
        violations["goal_reached"][..., -1].float().mean(dim=(1, 2)),
        violations["dest_reached"][..., -1].float().mean(dim=(1, 2)),
        (rewards * valid).sum(dim=(1, 2, 3)) / valid.sum(dim=(1, 2, 3)).clamp(min=1),
    ], dim=-1)
    return out.contiguous()


def all_gather_scenes(local: Tensor, n_scene_total: int, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """ONE collective for the whole metrics reduction: every rank contributes its [n_local, M] block (padded to the
    largest shard so the message has a fixed size) and receives the [n_scene_total, M] table in scene order."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    width = local.shape[1]
    cap = -(-n_scene_total // world)
    pad = torch.zeros(cap, width, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty(world * cap, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    rows = []
    for r in range(world):
        b, e = scene_shard(n_scene_total, r, world)
        rows.append(gathered[r * cap: r * cap + (e - b)])
    return torch.cat(rows, dim=0)

"""Scene sharding over the GPUs of one box: one process per GPU, scenes are independent, so the data
path has NO collective -- rank r owns a contiguous block of scenes (the K joint futures of a scene stay together
because they share the scene's map / traffic-light K|V caches).  The only exchange is the metrics reduction: one
fixed-layout all-gather of a packed per-scene buffer -- the WOMD records written by `tb_womd_pack`
(`models/metrics/womd.py::WOMDMetrics.update / gather`; replaces torchmetrics' six per-state all-gathers, reference
`src/models/metrics/womd.py:23,44-49`).  `all_gather_scenes` is the same collective for callers whose shards are uneven."""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

def scene_shard(n_scene: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the scenes owned by `rank`: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(n_scene, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch: Mapping[str, Tensor], rank: int, world: int) -> Dict[str, Tensor]:
    n = next(iter(batch.values())).shape[0]
    b, e = scene_shard(n, rank, world)
    return {k: v[b:e] for k, v in batch.items()}


def all_gather_scenes(local: Tensor, n_scene_total: int, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """ONE collective for the whole metrics reduction: every rank contributes its [n_local, M] block of per-scene records
    (any dtype; padded to the largest shard so the message has a fixed size) and receives the [n_scene_total, M] table in
    scene order."""
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

"""Seeded synthetic scenes in the reference's batch schema.

Schema = `DataH5womd.tensor_size_val` (reference `src/data_modules/data_h5_womd.py:85-173`), i.e. what the
reference's DataLoader hands to `WaymoMotion.validation_step` / `test_step`.  Scene i of a run is drawn from
`torch.Generator().manual_seed(seed + i)` on the CPU, so the same call reproduces bit-identical scenes in the
build container (where golden vectors are generated from the reference) and on the GPU box.

Everything is SDC-centred like `pack_h5.center_at_sdc` (`src/utils/pack_h5.py:348-416`): agent 0 sits near the origin.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

N_PL_NODE = 20
N_PL_TYPE = 11
DT = 0.1


def _scene(seed: int, n_agent: int, n_pl: int, n_step: int, n_tl: int, zero_tl: bool, single_agent: bool,
           area_scale: float = 1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def U(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    def RI(shape, lo, hi):  # integers in [lo, hi)
        return torch.randint(lo, hi, shape, generator=g)

    out: Dict[str, torch.Tensor] = {}
    # ---------------- map: straight polylines, 20 nodes 1 m apart ----------------
    start = U((n_pl, 1, 2), -100.0, 100.0) * area_scale
    heading = U((n_pl, 1), -math.pi, math.pi)
    step = torch.stack([heading.cos(), heading.sin()], dim=-1)  # [P,1,2] unit step
    k = torch.arange(N_PL_NODE, dtype=torch.float32).view(1, N_PL_NODE, 1)
    out["map/pos"] = (start + k * step).contiguous()
    out["map/dir"] = step.expand(n_pl, N_PL_NODE, 2).contiguous()
    n_valid_node = RI((n_pl,), 1, N_PL_NODE + 1)
    pl_valid = torch.rand((n_pl,), generator=g) < 0.9
    n_fix = min(n_pl, 10)
    pl_valid[:n_fix] = True  # make sure every destination type exists
    out["map/valid"] = (torch.arange(N_PL_NODE).view(1, -1) < n_valid_node.view(-1, 1)) & pl_valid.view(-1, 1)
    pl_type = RI((n_pl,), 0, N_PL_TYPE)
    pl_type[:n_fix] = torch.arange(n_fix) % 5
    out["map/type"] = torch.nn.functional.one_hot(pl_type, N_PL_TYPE).bool()
    vpos = out["map/pos"][out["map/valid"]]
    out["map/boundary"] = torch.stack([vpos[:, 0].min(), vpos[:, 0].max(), vpos[:, 1].min(), vpos[:, 1].max()])

    # ---------------- agents: constant acc / yaw-rate ground truth ----------------
    p0 = U((n_agent, 2), -80.0, 80.0) * area_scale
    p0[0] = U((2,), -1.0, 1.0)
    yaw0 = U((n_agent,), -math.pi, math.pi)
    spd0 = U((n_agent,), 0.0, 10.0)
    acc = U((n_agent,), -0.5, 0.5)
    yaw_rate = U((n_agent,), -0.1, 0.1)
    t = torch.arange(n_step, dtype=torch.float32).view(-1, 1) * DT  # [T,1]
    yaw = yaw0.view(1, -1) + yaw_rate.view(1, -1) * t
    spd = spd0.view(1, -1) + acc.view(1, -1) * t
    vel = torch.stack([spd * yaw.cos(), spd * yaw.sin()], dim=-1)  # [T,A,2]
    pos = p0.view(1, n_agent, 2) + torch.cumsum(torch.cat([torch.zeros(1, n_agent, 2), vel[:-1] * DT], 0), 0)
    a_valid = torch.rand((n_agent,), generator=g) < 0.9
    a_valid[0] = True
    if single_agent:
        a_valid[:] = False
        a_valid[0] = True
    valid = a_valid.view(1, -1).expand(n_step, -1).clone()
    # late spawn (valid from t in [1,10]) and early disappearance (invalid from t in [30, 80)) for a few agents
    n_special = max(1, n_agent // 8)
    for j in range(n_special):
        a = 1 + 2 * j
        if a < n_agent and not single_agent:
            valid[: int(RI((1,), 1, 11)), a] = False
        b = 2 + 2 * j
        if b < n_agent and n_step > 30 and not single_agent:
            valid[int(RI((1,), 30, min(80, n_step))):, b] = False
    out["agent/valid"] = valid
    out["agent/pos"] = pos.contiguous()
    out["agent/z"] = torch.zeros(n_step, n_agent, 1)
    out["agent/vel"] = vel.contiguous()
    out["agent/spd"] = spd.unsqueeze(-1).contiguous()
    out["agent/acc"] = acc.view(1, -1, 1).expand(n_step, -1, -1).contiguous()
    out["agent/yaw_bbox"] = yaw.unsqueeze(-1).contiguous()
    out["agent/yaw_rate"] = yaw_rate.view(1, -1, 1).expand(n_step, -1, -1).contiguous()
    a_type = RI((n_agent,), 0, 3)
    a_type[0] = 0
    out["agent/type"] = torch.nn.functional.one_hot(a_type, 3).bool()
    role = torch.zeros(n_agent, 3, dtype=torch.bool)
    role[0, 0] = True
    role[: min(8, n_agent), 2] = True
    out["agent/role"] = role
    size = torch.tensor([4.5, 2.0, 1.6]).view(1, 3).expand(n_agent, 3) * U((n_agent, 1), 0.8, 1.2)
    out["agent/size"] = size.contiguous()
    out["agent/cmd"] = torch.nn.functional.one_hot(RI((n_agent,), 0, 8), 8).bool()
    out["agent/goal"] = torch.cat([pos[-1], yaw[-1].unsqueeze(-1), spd[-1].unsqueeze(-1)], dim=-1).contiguous()
    # destination: a valid polyline whose type is compatible with the agent type (goal_manager.py:235-244)
    compat = {0: (0, 1, 2, 4), 1: (4,), 2: (3, 4)}
    dest = torch.zeros(n_agent, dtype=torch.int64)
    for a in range(n_agent):
        ok = torch.zeros(n_pl, dtype=torch.bool)
        for c in compat[int(a_type[a])]:
            ok |= pl_type == c
        ok &= pl_valid
        cand = ok.nonzero().flatten()
        dest[a] = cand[int(RI((1,), 0, len(cand)))]
    out["agent/dest"] = dest
    out["agent/object_id"] = torch.arange(n_agent, dtype=torch.int64)

    # ---------------- traffic lights ----------------
    tl_valid = torch.rand((n_step, n_tl), generator=g) < 0.3
    if zero_tl:
        tl_valid[:] = False
    out["tl_stop/valid"] = tl_valid
    out["tl_stop/state"] = torch.nn.functional.one_hot(RI((n_step, n_tl), 0, 5), 5).bool()
    out["tl_stop/pos"] = (U((1, n_tl, 2), -80.0, 80.0) * area_scale).expand(n_step, -1, -1).contiguous()
    tl_yaw = U((1, n_tl), -math.pi, math.pi).expand(n_step, -1)
    out["tl_stop/dir"] = torch.stack([tl_yaw.cos(), tl_yaw.sin()], dim=-1).contiguous()
    return out


def _plant_red_light_events(batch: Dict[str, torch.Tensor], frame: int = 5, n_event: int = 3) -> None:
    """moves the first `n_event` stop points of every scene 1 m behind the centre of a fast vehicle at `frame` (a teacher-forced
    frame, so the simulated agent is exactly there) and makes them valid + red at that frame: the vehicle's front box contains
    the stop point now and not 0.1 s later = a run-red-light event (traffic_rule_checker.py:199-258).  No RNG involved."""
    S = batch["agent/valid"].shape[0]
    for s in range(S):
        ok = batch["agent/valid"][s, frame] & batch["agent/type"][s, :, 0] & (batch["agent/spd"][s, frame, :, 0] >= 4.0)
        for j, a in enumerate(ok.nonzero().flatten()[:n_event].tolist()):
            yaw = batch["agent/yaw_bbox"][s, frame, a, 0]
            back = torch.stack([yaw.cos(), yaw.sin()]) * 1.0
            batch["tl_stop/pos"][s, :, j] = batch["agent/pos"][s, frame, a] - back
            batch["tl_stop/valid"][s, frame, j] = True
            batch["tl_stop/state"][s, frame, j] = torch.tensor([False, True, False, False, False])


def make_batch(n_scene: int, n_agent: int = 64, n_pl: int = 1024, seed: int = 0, n_step: int = 91,
               n_step_hist: int = 11, n_tl: int = 40, special_scenes: bool = True, area_scale: float = 1.0,
               plant_red_light: bool = False) -> Dict[str, torch.Tensor]:
    """A validation-style batch (`agent/*` = 91 frames of GT, `history/*` = first 11 frames) of CPU tensors.

    With `special_scenes`, scene 1 has no valid traffic light at all and scene 2 has exactly one valid agent
    (the two masking corner cases called out in SURVEY.md §8a "parity hazards").  `area_scale` < 1 packs agents, map and
    traffic lights into a smaller area (dense scenes: collisions, road-edge crossings and red-light events occur).
    """
    scenes = [
        _scene(seed + i, n_agent, n_pl, n_step, n_tl, zero_tl=special_scenes and i == 1,
               single_agent=special_scenes and i == 2)
        for i in range(n_scene)
    ]
    batch = {k: torch.stack([s[k] for s in scenes], dim=0) for k in scenes[0]}
    if plant_red_light:
        _plant_red_light_events(batch)
    for k in ("valid", "pos", "z", "vel", "spd", "acc", "yaw_bbox", "yaw_rate"):
        batch[f"history/agent/{k}"] = batch[f"agent/{k}"][:, :n_step_hist].contiguous()
    for k in ("type", "role", "size", "object_id"):
        batch[f"history/agent/{k}"] = batch[f"agent/{k}"]
    for k in ("valid", "state", "pos", "dir"):
        batch[f"history/tl_stop/{k}"] = batch[f"tl_stop/{k}"][:, :n_step_hist].contiguous()
    # agents that are not simulated: a minimal dummy block so that the reference pre-processing runs
    n_ns = 4
    zf = lambda *s: torch.zeros(n_scene, *s)  # noqa: E731
    batch["history/agent_no_sim/valid"] = torch.zeros(n_scene, n_step_hist, n_ns, dtype=torch.bool)
    batch["history/agent_no_sim/pos"] = zf(n_step_hist, n_ns, 2)
    batch["history/agent_no_sim/z"] = zf(n_step_hist, n_ns, 1)
    batch["history/agent_no_sim/vel"] = zf(n_step_hist, n_ns, 2)
    batch["history/agent_no_sim/spd"] = zf(n_step_hist, n_ns, 1)
    batch["history/agent_no_sim/yaw_bbox"] = zf(n_step_hist, n_ns, 1)
    batch["history/agent_no_sim/type"] = torch.zeros(n_scene, n_ns, 3, dtype=torch.bool)
    batch["history/agent_no_sim/size"] = zf(n_ns, 3)
    batch["history/agent_no_sim/object_id"] = torch.zeros(n_scene, n_ns, dtype=torch.int64)
    return batch


def make_mode_trajectories(n_scene: int, n_agent: int, n_pred: int, seed: int, n_step: int = 80, n_cluster: int = 4):
    """Synthetic multi-modal predictions for the post-processing tests: every agent has `n_cluster` distinct futures and each
    of its `n_pred` modes is one of them plus a small perturbation, so that NMS thresholds of ~1-3 m really merge modes.
    Returns (valid [S,A] bool, scores [S,A,n_pred] unnormalised > 0, trajs [S,A,n_pred,n_step,4] = x, y, yaw, spd)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    S, A, n, T = n_scene, n_agent, n_pred, n_step
    p0 = (torch.rand(S, A, 1, 1, 2, generator=g) - 0.5) * 100.0
    yaw0 = (torch.rand(S, A, n_cluster, 1, generator=g) - 0.5) * 2 * math.pi
    spd = torch.rand(S, A, n_cluster, 1, generator=g) * 12.0
    yaw_rate = (torch.rand(S, A, n_cluster, 1, generator=g) - 0.5) * 0.3
    pick = torch.randint(0, n_cluster, (S, A, n), generator=g)
    t = torch.arange(1, T + 1, dtype=torch.float32).view(1, 1, 1, T) * DT
    take = lambda x: torch.gather(x.expand(S, A, n_cluster, 1), 2, pick.unsqueeze(-1))  # noqa: E731
    yaw = take(yaw0) + take(yaw_rate) * t  # [S,A,n,T]
    v = take(spd).expand(-1, -1, -1, T)
    step = torch.stack([v * yaw.cos(), v * yaw.sin()], -1) * DT
    xy = p0 + torch.cumsum(step, dim=3) + torch.randn(S, A, n, 1, 2, generator=g) * 0.4
    trajs = torch.cat([xy, yaw.unsqueeze(-1), v.unsqueeze(-1)], dim=-1).contiguous()
    scores = torch.rand(S, A, n, generator=g) + 0.05
    valid = torch.rand(S, A, generator=g) < 0.85
    return valid, scores, trajs

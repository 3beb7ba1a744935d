"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch CPU ops, fp32) of the reference's OPTIONAL traffic-rule checks and of the collision term of the
differentiable reward -- SURVEY.md 8f-2 -- as the reference implements them, one call per decode step:

  collided        utils/traffic_rule_checker.py:122-160   (separating-axis test on the 4 corners, 1.1 x size, no ped/cyc pairs)
  run_road_edge   :163-196, :574-589, ccw :609-610        (vehicles only; bbox edges x road-edge segments of types 4, 5, 7)
  run_red_light   :199-258                                 (vehicles only; stop point inside the front box now, not 0.1 s later)
  passive         :261-335                                 (vehicles only; near a lane centre, slow, nothing ahead, > 20 steps)
  collision term  utils/rewards.py:49-115                  (5 circles per agent, relaxed overlap in [0, 1])

`RuleState` carries what `TrafficRuleChecker.__init__` precomputes (:30-75) and the sticky flags / passive counter of
`check` (:413-472).  Pinned against the unmodified reference by `tests/test_oracle_vs_reference.py` and by the golden case
`*_rules` of `oracle/make_golden.py`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
from torch import Tensor

OPTIONAL_KEYS = ("collided", "run_road_edge", "run_red_light", "passive")


def agent_bbox(state: Tensor, size_lw: Tensor) -> Tensor:
    """4 corners (rear-right, front-right, front-left, rear-left) of every agent box (:518-545). state [B,A,4], size [B,A,2]."""
    c, s = torch.cos(state[..., 2]), torch.sin(state[..., 2])
    fwd = 0.5 * size_lw[..., [0]].expand(-1, -1, 2) * torch.stack([c, s], -1)
    right = 0.5 * size_lw[..., [1]].expand(-1, -1, 2) * torch.stack([s, -c], -1)
    off = torch.stack([-fwd + right, fwd + right, fwd - right, -fwd - right], dim=2)
    return state[:, :, None, :2].expand(-1, -1, 4, -1) + off


def _ccw(a: Tensor, b: Tensor, c: Tensor) -> Tensor:
    return (c[..., 1] - a[..., 1]) * (b[..., 0] - a[..., 0]) > (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0])


@dataclass
class RuleState:
    size_lw: Tensor  # [B,A,2] length, width x collision_size_scale
    pair_off: Tensor  # [B,A,A] pairs that never collide: self, ped/cyc flag pairs (:56-61)
    veh: Tensor  # [B,A]
    edge: Tensor  # [B,P*20,2,2] road-edge segments (start, end)
    edge_valid: Tensor  # [B,P*20]
    lane: Tensor  # [B,P*20,2] lane-centre nodes
    lane_valid: Tensor  # [B,P*20]
    rl_len: Tensor  # [B,A,1]
    rl_wid: Tensor  # [B,A,1]
    tl_valid: Tensor  # [B,T_tl,TL]
    tl_pos: Tensor  # [B,T_tl,TL,2]
    tl_state: Tensor  # [B,T_tl,TL,5]
    enable: Dict[str, bool]
    sticky: Dict[str, Tensor] = field(default_factory=dict)
    passive_counter: Optional[Tensor] = None


def init_rules(agent_type: Tensor, agent_size: Tensor, map_valid: Tensor, map_type: Tensor, map_pos: Tensor, map_dir: Tensor,
               tl_valid: Tensor, tl_pos: Tensor, tl_state: Tensor, enable: Dict[str, bool], collision_size_scale: float = 1.1
               ) -> RuleState:
    B, A = agent_type.shape[:2]
    eye = torch.eye(A, dtype=torch.bool)[None].expand(B, -1, -1)
    pc = agent_type[:, :, 1]
    rs = RuleState(
        size_lw=agent_size[..., :2] * collision_size_scale,
        pair_off=eye | (pc.unsqueeze(1) & pc.unsqueeze(2)),
        veh=agent_type[:, :, 0],
        edge=torch.stack([map_pos, map_pos + map_dir], dim=-2).flatten(1, 2),
        edge_valid=(map_valid & map_type[:, :, [4, 5, 7]].any(-1, keepdim=True)).flatten(1, 2),
        lane=map_pos.flatten(1, 2),
        lane_valid=(map_valid & map_type[:, :, :3].any(-1, keepdim=True)).flatten(1, 2),
        rl_len=agent_size[:, :, [0]] * 0.5 * 0.6,
        rl_wid=agent_size[:, :, [1]] * 0.5 * 1.8,
        tl_valid=tl_valid, tl_pos=tl_pos, tl_state=tl_state, enable=dict(enable))
    rs.sticky = {k: torch.zeros(B, A, dtype=torch.bool) for k in OPTIONAL_KEYS}
    rs.passive_counter = torch.zeros(B, A)
    return rs


def check_collided(valid: Tensor, bbox: Tensor, pair_off: Tensor) -> Tensor:
    nxt = bbox.roll(-1, dims=2)
    line = torch.cat([nxt[..., [1]] - bbox[..., [1]], bbox[..., [0]] - nxt[..., [0]],
                      nxt[..., [0]] * bbox[..., [1]] - nxt[..., [1]] * bbox[..., [0]]], dim=-1)  # [B,A,4,3]: ax + by + c = 0
    pt = torch.cat([bbox, torch.ones_like(bbox[..., [0]])], dim=-1)  # [B,A,4,3]
    A = bbox.shape[1]
    line = line[:, :, None, :, None, :].expand(-1, -1, A, -1, 4, -1)
    pt = pt[:, None, :, None, :, :].expand(-1, A, -1, 4, -1, -1)
    outside = torch.sum(line * pt, dim=-1) > 0  # [B, A(lines of i), A(points of j), 4 lines, 4 points]
    sep = torch.any(torch.all(outside, dim=-1), dim=-1)  # an edge of i has all corners of j outside
    sep = sep | sep.transpose(1, 2)
    sep = sep | pair_off | ~(valid[:, :, None] & valid[:, None, :])
    return ~sep.all(-1)


def check_run_road_edge(valid: Tensor, bbox: Tensor, veh: Tensor, edge: Tensor, edge_valid: Tensor) -> Tensor:
    nxt = bbox.roll(-1, dims=2)
    a, b = bbox.unsqueeze(2), nxt.unsqueeze(2)  # [B,A,1,4,2]
    c, d = edge[:, None, :, None, 0], edge[:, None, :, None, 1]  # [B,1,E,1,2]
    hit = (_ccw(a, c, d) != _ccw(b, c, d)) & (_ccw(a, b, c) != _ccw(a, b, d))  # [B,A,E,4]
    hit = hit.any(-1) & edge_valid.unsqueeze(1)
    return hit.any(-1) & valid & veh


def check_run_red_light(valid: Tensor, state: Tensor, tl_valid: Tensor, tl_pos: Tensor, tl_state: Tensor, rl_len: Tensor,
                        rl_wid: Tensor, veh: Tensor) -> Tensor:
    c, s = torch.cos(state[..., 2]), torch.sin(state[..., 2])
    hf = torch.stack([c, s], -1).unsqueeze(2)
    hr = torch.stack([s, -c], -1).unsqueeze(2)
    p0 = state[..., :2].unsqueeze(2)
    p1 = p0 + 0.1 * state[..., [3]].unsqueeze(2) * hf
    tp = tl_pos.unsqueeze(1)

    def inside(p):
        return torch.logical_and(torch.abs(torch.sum((tp - p) * hf, dim=-1)) < rl_len,
                                 torch.abs(torch.sum((tp - p) * hr, dim=-1)) < rl_wid)

    m_agent = (valid & veh).unsqueeze(2)
    m_tl = (tl_valid & tl_state[:, :, 1]).unsqueeze(1)
    return (inside(p0) & ~inside(p1) & m_agent & m_tl).any(-1)


def check_passive(valid: Tensor, state: Tensor, counter: Tensor, tl_valid: Tensor, tl_pos: Tensor, tl_state: Tensor,
                  lane: Tensor, lane_valid: Tensor, veh: Tensor):
    A = valid.shape[1]
    near = torch.norm(state[:, :, :2].unsqueeze(2) - lane.unsqueeze(1), dim=-1) < 2
    near = (near & lane_valid.unsqueeze(1)).any(-1)
    slow = state[:, :, 3] < 5
    hf = torch.stack([torch.cos(state[..., 2]), torch.sin(state[..., 2])], -1).unsqueeze(2)
    m_tl = (tl_valid & tl_state[:, :, [0, 1, 2, 4]].any(-1)).unsqueeze(1)
    v = tl_pos.unsqueeze(1) - state[:, :, :2].unsqueeze(2)
    n = torch.norm(v, dim=-1)
    red_ahead = ((n < 10) & (((hf * v).sum(-1) / n) > 0.95) & m_tl).any(-1)
    av = state[:, :, :2].unsqueeze(1) - state[:, :, :2].unsqueeze(2)  # [B, i, j]: j relative to i
    an = torch.norm(av, dim=-1)
    eye = torch.eye(A, dtype=torch.bool)[None]
    agent_ahead = ((an < 10) & (((hf * av).sum(-1) / an) > 0.95) & valid.unsqueeze(1) & valid.unsqueeze(2) & ~eye).any(-1)
    raw = valid & veh & near & slow & ~red_ahead & ~agent_ahead
    counter = (counter + raw) * raw
    return counter > 20, counter


def check_optional(rs: RuleState, step: int, valid: Tensor, state: Tensor) -> Dict[str, Tensor]:
    """the optional part of `TrafficRuleChecker.check` (:420-472) for one step; disabled checks report their sticky state."""
    bbox = agent_bbox(state, rs.size_lw)
    out: Dict[str, Tensor] = {}
    tl_step = min(step, rs.tl_valid.shape[1] - 1)
    this: Dict[str, Tensor] = {}
    if rs.enable.get("collided"):
        this["collided"] = check_collided(valid, bbox, rs.pair_off)
    if rs.enable.get("run_road_edge"):
        this["run_road_edge"] = check_run_road_edge(valid, bbox, rs.veh, rs.edge, rs.edge_valid)
    if rs.enable.get("run_red_light"):
        this["run_red_light"] = check_run_red_light(valid, state, rs.tl_valid[:, tl_step], rs.tl_pos[:, tl_step],
                                                    rs.tl_state[:, tl_step], rs.rl_len, rs.rl_wid, rs.veh)
    if rs.enable.get("passive"):
        if not rs.enable.get("run_red_light"):  # the reference reads `tl_step` that only the red-light branch defines (:441,:457)
            raise RuntimeError("enable_check_passive needs enable_check_run_red_light (reference bug, SURVEY 8a)")
        this["passive"], rs.passive_counter = check_passive(valid, state, rs.passive_counter, rs.tl_valid[:, tl_step],
                                                             rs.tl_pos[:, tl_step], rs.tl_state[:, tl_step], rs.lane,
                                                             rs.lane_valid, rs.veh)
    for k in OPTIONAL_KEYS:
        if k in this:
            rs.sticky[k] = rs.sticky[k] | this[k]
            out[k + "_this_step"] = this[k]
        else:
            out[k + "_this_step"] = rs.sticky[k]
        out[k] = rs.sticky[k]
    return out


def collision_term(valid: Tensor, state: Tensor, agent_size: Tensor, reduce_with_max: bool) -> Tensor:
    """relaxed collision in [0,1] per agent (rewards.py:49-114), before the `-w_collision *` and the validity mask."""
    B, A = valid.shape
    eps = torch.finfo(state.dtype).eps
    xy, yaw = state[..., :2], state[..., 2]
    h = torch.stack([torch.cos(yaw), torch.sin(yaw)], -1)
    w = agent_size[:, :, :2].amin(-1)
    l = agent_size[:, :, :2].amax(-1)
    d = ((l - w) / 4.0).unsqueeze(-1).expand(-1, -1, 2)
    cen = xy.unsqueeze(2).expand(-1, -1, 5, -1) + torch.stack([-2 * h * d, -1 * h * d, 0 * h * d, 1 * h * d, 2 * h * d], dim=2)
    c0 = cen.unsqueeze(2).expand(-1, -1, A, -1, -1)
    c1 = c0.transpose(1, 2)
    r = w.unsqueeze(-1).expand(-1, -1, A) / 2.0 + eps
    r_sum = r.transpose(1, 2) + r
    dist = torch.zeros(B, A, A, 5, 5)
    for i in range(5):
        for j in range(5):
            dist[:, :, :, i, j] = torch.norm(c0[:, :, :, i] - c1[:, :, :, j], dim=-1) + eps
    dist = dist.flatten(3, 4).min(-1)[0]
    col = torch.clamp(1 - dist / r_sum, min=0)
    off = torch.eye(A, dtype=torch.bool)[None].expand(B, -1, -1) | ~valid[:, :, None] | ~valid[:, None, :]
    col = col.masked_fill(off, 0.0)
    if reduce_with_max:
        return col.amax(2)
    return torch.clamp(col, max=1).sum(-1) / valid.sum(-1, keepdim=True)

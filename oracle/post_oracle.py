"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the step right after the rollout:

  post_process   `WaymoPostProcessing.forward` (data_modules/waymo_post_processing.py:33-81) with `mtr_nms` (:126-171),
                 `traj_topk` (:173-193) and `mpa_nms` (:83-124); `traj_aggr` (k-means EM, :195-295) is not restated
                 (`aggr_thresh` is empty in configs/model/traffic_bots.yaml:185).
  womd_pack      `WOMDMetrics.update` (models/metrics/womd.py:60-145), `interactive_challenge=False`: per scene the agents to
                 predict (`agent/role[..., 2]`) first, then the other fully observed agents, packed into the six tensors the
                 Waymo metrics op consumes.

Written per (scene, agent) with explicit loops where the reference uses batched indexing; pinned against the reference's
own classes executed live (`tests/test_oracle_vs_reference.py`) and against `tests/golden/post_*.npz`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor


def _type_thresh(agent_type: Tensor, thresh: Sequence[float]) -> Tensor:
    t = torch.zeros(agent_type.shape[:2])
    for i, v in enumerate(thresh):
        t = t + agent_type[:, :, i] * v
    return t  # [S,A]


def _within(xy: Tensor, thresh: Tensor, use_ade: bool) -> Tensor:
    """[S,A,n,n] bool: trajectories i, j closer than the per-agent threshold (ADE over the steps or final displacement)."""
    if use_ade:
        d = torch.norm(xy.unsqueeze(2) - xy.unsqueeze(3), dim=-1).mean(-1)
    else:
        d = torch.norm(xy[:, :, :, -1].unsqueeze(2) - xy[:, :, :, -1].unsqueeze(3), dim=-1)
    return d < thresh[:, :, None, None]


def mtr_nms(trajs: Tensor, scores: Tensor, k_pred: int, thresh: Sequence[float], use_ade: bool, agent_type: Tensor):
    S, A, n = scores.shape
    near = _within(trajs[..., :2], _type_thresh(agent_type, thresh), use_ade)
    idx = torch.zeros(S, A, k_pred, dtype=torch.int64)
    for s in range(S):
        for a in range(A):
            sc = scores[s, a].clone()
            for k in range(k_pred):
                j = int(sc.max(-1)[1])
                sc = sc * ((~near[s, a, j]) * 0.99 + 0.01)  # suppress everything close to the pick
                sc[j] = -1
                idx[s, a, k] = j
    si = torch.arange(S)[:, None, None]
    ai = torch.arange(A)[None, :, None]
    sk = scores[si, ai, idx]
    return trajs[si, ai, idx], sk / sk.sum(-1, keepdim=True), idx


def traj_topk(trajs: Tensor, scores: Tensor, k_pred: int):
    """the k highest scores; the reference's `topk(sorted=False)` leaves the order unspecified -- here descending."""
    idx = scores.topk(k_pred, dim=-1, sorted=True)[1]
    si = torch.arange(scores.shape[0])[:, None, None]
    ai = torch.arange(scores.shape[1])[None, :, None]
    sk = scores[si, ai, idx]
    return trajs[si, ai, idx], sk / sk.sum(-1, keepdim=True), idx


def mpa_nms(valid: Tensor, trajs: Tensor, scores: Tensor, thresh: Sequence[float], use_ade: bool, agent_type: Tensor) -> Tensor:
    near = _within(trajs[..., :2], _type_thresh(agent_type, thresh), use_ade)
    scores = scores.clone()
    S, A, n = scores.shape
    for s in range(S):
        for a in range(A):
            if not valid[s, a]:
                continue
            for k in scores[s, a].argsort(descending=True).tolist():  # order fixed before the in-place edits
                if bool((near[s, a, k] & (scores[s, a] > scores[s, a, k])).any()):
                    scores[s, a, k] = 1e-3
    return scores / scores.sum(-1, keepdim=True)


def post_process(valid: Tensor, scores: Tensor, trajs: Tensor, agent_type: Tensor, k_pred: int = 6, score_temperature: float = 1e2,
                 mpa_nms_thresh: Sequence[float] = (), mtr_nms_thresh: Sequence[float] = (), use_ade: bool = True
                 ) -> Dict[str, Optional[Tensor]]:
    """valid [S,A]; scores [S,A,n] unnormalised; trajs [S,A,n,Tf,4] -> the reference's pred_dict (+ `mode_idx` when modes were
    selected)."""
    scores = scores / scores.sum(-1, keepdim=True)
    n = trajs.shape[2]
    idx = None
    if n > k_pred:
        if len(mtr_nms_thresh) > 0:
            trajs, scores, idx = mtr_nms(trajs, scores, k_pred, mtr_nms_thresh, use_ade, agent_type)
        else:
            trajs, scores, idx = traj_topk(trajs, scores, k_pred)
    if len(mpa_nms_thresh) > 0:
        scores = mpa_nms(valid, trajs, scores, mpa_nms_thresh, use_ade, agent_type)
    if score_temperature > 0:
        scores = torch.softmax(torch.log(scores) / score_temperature, dim=-1)
    t = trajs.movedim(3, 1)  # [S,Tf,A,k,4]
    return {"waymo_trajs": t[..., :2], "waymo_yaw_bbox": t[..., 2:3], "waymo_spd": t[..., 3:4], "waymo_scores": scores,
            "waymo_valid": valid.unsqueeze(1).expand(-1, t.shape[1], -1), "mode_idx": idx}


def womd_pack(batch: Dict[str, Tensor], pred_traj: Tensor, pred_score: Optional[Tensor], step_gt: int = 90, step_current: int = 10,
              m_joint: int = 8) -> Dict[str, Tensor]:
    """pred_traj [S, Tf (= steps step_current+1 .. step_gt), A, K, 2], pred_score [S,A,K] or None."""
    track_future = step_gt - step_current
    S, A = batch["agent/type"].shape[:2]
    mask_pred = batch["agent/role"][..., 2]
    mask_other = (~mask_pred) & batch["agent/valid"][:, : step_current + 1].all(1)
    n_frame = batch["agent/pos"].shape[1]
    gt = torch.cat([batch["agent/pos"], batch["agent/size"][..., :2].unsqueeze(1).expand(-1, n_frame, -1, -1),
                    batch["agent/yaw_bbox"], batch["agent/vel"]], dim=-1).transpose(1, 2)[:, :, : step_gt + 1]
    gt_valid = batch["agent/valid"].transpose(1, 2)[:, :, : step_gt + 1]
    obj_type = batch["agent/type"].float().argmax(-1) + 1.0
    p = pred_traj[:, 4:track_future:5].permute(0, 2, 3, 1, 4)  # [S,A,K,16,2]
    K, n_step = p.shape[2], p.shape[3]
    if pred_score is None:
        pred_score = torch.full((S, A, K), 1.0 / K)
    out = {
        "prediction_trajectory": torch.zeros(S, m_joint, K, 1, n_step, 2),
        "prediction_score": torch.zeros(S, m_joint, K),
        "ground_truth_trajectory": torch.zeros(S, A, gt.shape[2], 7),
        "ground_truth_is_valid": torch.zeros(S, A, gt.shape[2], dtype=torch.bool),
        "prediction_ground_truth_indices_mask": torch.zeros(S, m_joint, 1, dtype=torch.bool),
        "object_type": torch.zeros(S, A),
    }
    for s in range(S):
        first: List[int] = mask_pred[s].nonzero().flatten().tolist()
        rest: List[int] = mask_other[s].nonzero().flatten().tolist()
        for slot, a in enumerate(first):
            out["prediction_trajectory"][s, slot, :, 0] = p[s, a]
            out["prediction_score"][s, slot] = pred_score[s, a]
            out["prediction_ground_truth_indices_mask"][s, slot] = True
        for slot, a in enumerate(first + rest):
            out["ground_truth_trajectory"][s, slot] = gt[s, a]
            out["ground_truth_is_valid"][s, slot] = gt_valid[s, a]
            out["object_type"][s, slot] = obj_type[s, a]
    return out

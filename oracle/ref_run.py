"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference on a synthetic batch (build container only).

Mirrors what `WaymoMotion.validation_step` does up to the rollouts (reference
`src/pl_modules/waymo_motion.py:574-601,683-690`) and returns every tensor the parity tests compare.
"""
from __future__ import annotations

from typing import Dict

import torch

VIOLATION_KEYS = (
    "outside_map", "outside_map_this_step", "collided", "collided_this_step", "run_road_edge",
    "run_road_edge_this_step", "run_red_light", "run_red_light_this_step", "passive", "passive_this_step",
    "goal_reached", "goal_reached_this_step", "dest_reached", "dest_reached_this_step",
)


def buffer_to_dict(buf) -> Dict[str, torch.Tensor]:
    out = {
        "preds": buf.preds, "valid": buf.valid, "override_masks": buf.override_masks,
        "diffbar_rewards": buf.diffbar_rewards, "diffbar_rewards_valid": buf.diffbar_rewards_valid,
        "latent_log_probs": buf.latent_log_probs, "action_log_probs": buf.action_log_probs,
    }
    for k in VIOLATION_KEYS:
        out[f"violations/{k}"] = buf.violations[k]
    return {k: v.detach().clone() for k, v in out.items()}


@torch.no_grad()
def run_reference(model, batch: Dict[str, torch.Tensor], k_futures: int = 1, sample_seed: int = 7,
                  do_reactive_replay: bool = True) -> Dict[str, torch.Tensor]:
    """encode -> prior/posterior latent -> dest prediction -> reactive_replay -> joint_future_pred(K)."""
    model.eval()
    model.hparams.n_joint_future = k_futures
    batch = {k: v.clone() for k, v in batch.items()}
    batch = model.pre_processing(batch)
    input_dict = {k.split("input/")[-1]: v for k, v in batch.items() if "input/" in k}
    latent_post_dict = {k.split("latent_post/")[-1]: v for k, v in batch.items() if "latent_post/" in k}
    latent_prior_dict = {k.split("latent_prior/")[-1]: v for k, v in batch.items() if "latent_prior/" in k}
    feat = model.model.encode_input_features(**input_dict)
    feat_post = model.model.encode_input_features(**latent_post_dict)
    feat_prior = model.model.encode_input_features(**latent_prior_dict)
    res: Dict[str, torch.Tensor] = {f"enc/{k}": v.detach().clone() for k, v in feat.items()}

    goal_gt, goal_valid = model.model.goal_manager.get_gt_goal(
        agent_valid=input_dict["agent_valid"], gt_dest=batch["gt/dest"], gt_goal=batch["gt/goal"])
    goal_pred = model.model.goal_manager.pred_goal(
        agent_type=batch["ref/agent_type"], map_type=batch["ref/map_type"], agent_state=batch["ref/agent_state"],
        **feat)
    res["dest/probs"] = goal_pred.probs.detach().clone()
    res["dest/valid"] = goal_pred.valid.clone()
    latent_post = model.model.latent_encoder(posterior=True, **feat_post)
    latent_prior = model.model.latent_encoder(**feat_prior)
    res["latent_prior/mean"] = latent_prior.mean.clone()
    res["latent_prior/stddev"] = latent_prior.stddev.clone()
    res["latent_post/mean"] = latent_post.mean.clone()
    res["goal_valid"] = goal_valid.clone()

    if do_reactive_replay:
        buf = model.reactive_replay(
            batch=batch, input_feature_dict=feat,
            mask_teacher_forcing=model.teacher_forcing_reactive_replay.get(batch["gt/valid"], 0),
            latent=latent_post, goal=goal_gt, goal_valid=goal_valid, deterministic_latent=True,
            deterministic_action=True, require_vis_dict=False)
        for k, v in buffer_to_dict(buf).items():
            res[f"replay/{k}"] = v
        res["replay/hidden"] = model.model.hidden.detach().clone()

    torch.manual_seed(sample_seed)
    buf, goal_sample, goal_log_probs = model.joint_future_pred(
        batch=batch, input_feature_dict=feat, latent=latent_prior, goal=goal_pred, goal_valid=goal_valid,
        require_vis_dict=False)
    for k, v in buffer_to_dict(buf).items():
        res[f"jfp/{k}"] = v
    res["jfp/goal_sample"] = goal_sample.clone()
    res["jfp/goal_log_probs"] = goal_log_probs.clone()
    res["jfp/latent_sample"] = model.model.latent_sample.detach().clone()
    res["jfp/hidden"] = model.model.hidden.detach().clone()
    return res

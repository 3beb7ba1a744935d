"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/*.npz` from the UNMODIFIED reference.

Run in the build container (needs `/root/reference`):   python oracle/make_golden.py

For each case: seeded synthetic scenes (`trafficbots_b200.synthetic.make_batch`) and seeded parameters
(`trafficbots_b200.weights.init_state_dict`) are loaded into the reference `WaymoMotion` (strict
`load_state_dict`), the reference's own `pre_processing -> encode_input_features -> latent_encoder -> pred_goal ->
reactive_replay -> joint_future_pred(K)` is executed (`oracle/ref_run.py`), and its outputs are stored.  Inputs and
parameters are NOT stored -- they are regenerated from the seeds (a checksum of both is stored and verified).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402
import ref_run  # noqa: E402
from trafficbots_b200 import synthetic, weights  # noqa: E402

CASES = {
    # name: (n_scene, n_agent, n_pl, K, scene_seed, weight_seed, sample_seed)
    "cfg1_s1_a8_p64_k1": (1, 8, 64, 1, 11, 2023, 7),  # BASELINE.json configs[0] shape (parity gate)
    "s3_a8_p64_k2": (3, 8, 64, 2, 100, 2023, 7),  # + zero-TL scene, single-valid-agent scene, K>1 sampling
    "s1_a64_p1024_k1": (1, 64, 1024, 1, 500, 2023, 7),  # BASELINE.json configs[1] per-scene shape
}
# the four optional traffic-rule checks and the collision reward switched ON in the reference, dense scenes
# (area_scale 0.25) so that collisions / road-edge crossings / red-light / passive events really occur
RULE_CASES = {
    # name: (n_scene, n_agent, n_pl, K, scene_seed, weight_seed, sample_seed, area_scale, w_collision, reduce_with_max)
    "s2_a16_p96_k2_rules": (2, 16, 96, 2, 900, 2023, 7, 0.25, 0.5, True),
    "s2_a12_p64_k1_rules_sum": (2, 12, 64, 1, 950, 2023, 7, 0.2, 1.0, False),
}
# WaymoPostProcessing.forward + WOMDMetrics.update of the reference on seeded multi-modal trajectories
POST_CASES = {
    # name: (n_scene, n_agent, n_pred, scene_seed, traj_seed, mtr_nms_thresh, mpa_nms_thresh, use_ade)
    "default": (3, 12, 6, 77, 5, [], [], True),
    "mtr_nms": (3, 12, 14, 77, 5, [2.5, 1.0, 1.5], [], True),
    "mpa_nms": (3, 12, 6, 77, 5, [], [2.5, 1.0, 1.5], True),
    "mtr_mpa_fde": (3, 12, 10, 77, 5, [3.0, 1.5, 2.0], [4.0, 2.0, 3.0], False),
    "a64_k6": (1, 64, 6, 300, 9, [], [2.5, 1.0, 1.5], True),
}
RULES_ON = {"enable_check_collided": True, "enable_check_run_road_edge": True, "enable_check_run_red_light": True,
            "enable_check_passive": True}
RULE_BUF = tuple(f"violations/{k}{s}" for k in ("collided", "run_road_edge", "run_red_light", "passive") for s in ("", "_this_step"))

KEEP = (
    "enc/map_feature", "enc/map_feature_valid", "enc/agent_feature", "enc/tl_feature", "dest/probs",
    "latent_prior/mean", "latent_post/mean", "jfp/goal_sample", "jfp/goal_log_probs", "jfp/latent_sample",
    "jfp/hidden", "replay/hidden",
)
BUF = ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid", "latent_log_probs",
       "action_log_probs", "violations/outside_map", "violations/outside_map_this_step", "violations/goal_reached",
       "violations/goal_reached_this_step", "violations/dest_reached", "violations/dest_reached_this_step")


def checksum(tensors) -> float:
    """order-dependent fp64 checksum of a dict of tensors (inputs / parameters are regenerated, not stored)."""
    acc = 0.0
    for i, (k, v) in enumerate(sorted(tensors.items())):
        acc += (i + 1) * float(v.double().sum()) + 1e-3 * float(v.double().abs().sum())
    return acc


def main() -> None:
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--post-only" in sys.argv:
        make_post_goldens(out_dir)
        return
    only_rules = "--rules-only" in sys.argv  # keeps the existing fixtures byte-identical
    for name, (S, A, P, K, seed, wseed, sseed) in ({} if only_rules else CASES).items():
        model = ref_loader.build_reference(n_agent=A, n_pl=P, n_joint_future=K)
        sd = weights.init_state_dict(wseed)
        model.load_state_dict(sd, strict=True)
        batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed)
        res = ref_run.run_reference(model, batch, k_futures=K, sample_seed=sseed)
        keep = {k: res[k] for k in KEEP}
        for leg in ("jfp", "replay"):
            for b in BUF:
                keep[f"{leg}/{b}"] = res[f"{leg}/{b}"]
        arrays = {k.replace("/", "__"): v.numpy() for k, v in keep.items()}
        arrays["meta__case"] = np.array([S, A, P, K, seed, wseed, sseed], dtype=np.int64)
        arrays["meta__checksum_batch"] = np.array(checksum(batch))
        arrays["meta__checksum_weights"] = np.array(checksum(sd))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")
    for name, (S, A, P, K, seed, wseed, sseed, area, wcol, rmax) in RULE_CASES.items():
        model = ref_loader.build_reference(n_agent=A, n_pl=P, n_joint_future=K, overrides={
            "traffic_rule_checker": dict(RULES_ON),
            "differentiable_reward": {"w_collision": wcol, "reduce_collsion_with_max": rmax}})
        sd = weights.init_state_dict(wseed)
        model.load_state_dict(sd, strict=True)
        batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed, special_scenes=False, area_scale=area, plant_red_light=True)
        res = ref_run.run_reference(model, batch, k_futures=K, sample_seed=sseed)
        keep = {k: res[k] for k in ("jfp/goal_sample", "jfp/latent_sample", "latent_post/mean")}
        for leg in ("jfp", "replay"):
            for b in BUF + RULE_BUF:
                keep[f"{leg}/{b}"] = res[f"{leg}/{b}"]
        arrays = {k.replace("/", "__"): v.numpy() for k, v in keep.items()}
        arrays["meta__case"] = np.array([S, A, P, K, seed, wseed, sseed], dtype=np.int64)
        arrays["meta__rules"] = np.array([area, wcol, float(rmax)], dtype=np.float64)
        arrays["meta__checksum_batch"] = np.array(checksum(batch))
        arrays["meta__checksum_weights"] = np.array(checksum(sd))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        events = {k.split("/")[-1]: int(res[f"replay/{k}"].sum()) for k in RULE_BUF if not k.endswith("_this_step")}
        print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", "sticky-flag counts (replay):", events)
    make_post_goldens(out_dir)


def make_post_goldens(out_dir: str) -> None:
    ref_loader.install_stubs()
    from data_modules.waymo_post_processing import WaymoPostProcessing  # noqa: the reference's own classes
    from models.metrics.womd import WOMDMetrics
    arrays = {}
    for name, (S, A, n, seed, tseed, mtr, mpa, ade) in POST_CASES.items():
        batch = synthetic.make_batch(S, n_agent=A, n_pl=16, seed=seed)
        valid, scores, trajs = synthetic.make_mode_trajectories(S, A, n, seed=tseed)
        pp = WaymoPostProcessing(k_pred=6, score_temperature=1e2, mpa_nms_thresh=mpa, mtr_nms_thresh=mtr, aggr_thresh=[], n_iter_em=3,
                                 use_ade=ade)
        ref = pp(valid=valid, scores=scores.clone(), trajs=trajs.clone(), agent_type=batch["agent/type"])
        m = WOMDMetrics("val", step_gt=90, step_current=10, interactive_challenge=False)
        m.update(batch, ref["waymo_trajs"], ref["waymo_scores"])
        for k in ("waymo_trajs", "waymo_yaw_bbox", "waymo_spd", "waymo_scores"):
            arrays[f"{name}__{k}"] = ref[k].contiguous().numpy()
        for k in ("prediction_trajectory", "prediction_score", "ground_truth_trajectory", "ground_truth_is_valid",
                  "prediction_ground_truth_indices_mask", "object_type"):
            arrays[f"{name}__{k}"] = getattr(m, k + "_gpu")[0].numpy()
        arrays[f"{name}__meta"] = np.array([S, A, n, seed, tseed, int(ade)], dtype=np.int64)
        arrays[f"{name}__mtr"] = np.array(mtr, dtype=np.float64)
        arrays[f"{name}__mpa"] = np.array(mpa, dtype=np.float64)
        arrays[f"{name}__checksum"] = np.array(checksum({"s": scores, "t": trajs, "v": valid.float()}))
        uniform = torch.softmax(torch.log(scores / scores.sum(-1, keepdim=True)) / 1e2, -1)
        print("post", name, "scores changed by NMS:", bool(n != 6 or (ref["waymo_scores"] - uniform).abs().max() > 1e-4))
    path = os.path.join(out_dir, "post_cases.npz")
    np.savez_compressed(path, **arrays)
    print("post ->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

"""CPU: the oracle restatement reproduces the golden vectors that the UNMODIFIED reference produced
(`oracle/make_golden.py`).  Tolerances: encoders are the same ATen kernels -> 1e-6; the 90-step closed loop
accumulates GRU-kernel rounding differences -> 5e-4 m (reference fp32-vs-fp64 noise floor is 1.4e-4 m)."""
import pytest
import torch

import trafficbots_oracle as orc
from golden_util import CASES, RULE_CASES, RULE_KEYS, RULES_ON, load_case

BOOL_KEYS = ("valid", "override_masks", "diffbar_rewards_valid", "outside_map", "outside_map_this_step",
             "goal_reached", "goal_reached_this_step", "dest_reached", "dest_reached_this_step")


def _g(gold, leg, k):
    return gold[f"{leg}/{k}"] if f"{leg}/{k}" in gold else gold[f"{leg}/violations/{k}"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    gold, sd, batch, meta = load_case(case)
    feat = orc.encode_scene(sd, batch)
    assert torch.equal(feat["map_feature_valid"], gold["enc/map_feature_valid"])
    for k in ("map_feature", "agent_feature", "tl_feature"):
        assert (feat[k] - gold[f"enc/{k}"]).abs().max() <= 1e-6, k
    jfp = orc.joint_future_pred(sd, batch, k=meta["K"], sample_seed=meta["sseed"], feat=feat)
    rep = orc.reactive_replay(sd, batch, feat=feat)
    assert (jfp["dest_probs"] - gold["dest/probs"]).abs().max() <= 1e-6
    assert (jfp["latent_prior_mean"] - gold["latent_prior/mean"]).abs().max() <= 1e-6
    assert (rep["latent_post_mean"] - gold["latent_post/mean"]).abs().max() <= 1e-6
    assert torch.equal(jfp["goal_sample"], gold["jfp/goal_sample"])
    assert (jfp["latent_sample"] - gold["jfp/latent_sample"]).abs().max() <= 1e-6
    for leg, res in (("jfp", jfp), ("replay", rep)):
        for k in BOOL_KEYS:
            assert torch.equal(res[k], _g(gold, leg, k)), (leg, k)
        assert (res["preds"] - gold[f"{leg}/preds"]).abs().max() <= 5e-4, leg
        assert (res["diffbar_rewards"] - gold[f"{leg}/diffbar_rewards"]).abs().max() <= 5e-4, leg
        assert (res["action_log_probs"] - gold[f"{leg}/action_log_probs"]).abs().max() <= 1e-5, leg
        assert (res["latent_log_probs"] - gold[f"{leg}/latent_log_probs"]).abs().max() <= 1e-4, leg
        assert (res["hidden"] - gold[f"{leg}/hidden"]).abs().max() <= 5e-4, leg


def test_golden_covers_corner_cases():
    """the fixtures really contain the masking corner cases (SURVEY.md §8a parity hazards)."""
    gold, sd, batch, meta = load_case("s3_a8_p64_k2")
    assert not batch["tl_stop/valid"][1].any()  # scene without any valid traffic light
    assert int(batch["agent/valid"][2, 0].sum()) == 1  # scene with exactly one valid agent
    assert gold["jfp/override_masks"][..., 10:].any() is not None
    assert gold["replay/violations/goal_reached"].any() and gold["jfp/violations/dest_reached"].any()
    # late spawn: some agent invalid at t=0 becomes valid through an override
    v = gold["jfp/valid"][0, :, 0]
    assert (~v[:, 0] & v[:, -1]).any()


@pytest.mark.parametrize("case", RULE_CASES)
def test_oracle_optional_rule_checks_match_reference_golden(case):
    """the reference ran with the four optional checks and the collision reward ON (dense scenes with
    collisions, road-edge crossings, planted red-light events, passive vehicles): the oracle's restatement
    (`oracle/rule_checks_oracle.py`) reproduces its 8 extra violation maps bit for bit and the reward."""
    gold, sd, batch, meta = load_case(case)
    kw = dict(rules_enable=RULES_ON, w_collision=meta["w_collision"], reduce_collision_with_max=meta["reduce_with_max"])
    jfp = orc.joint_future_pred(sd, batch, k=meta["K"], sample_seed=meta["sseed"], **kw)
    rep = orc.reactive_replay(sd, batch, **kw)
    seen = set()
    for leg, res in (("jfp", jfp), ("replay", rep)):
        for k in BOOL_KEYS + RULE_KEYS:
            assert torch.equal(res[k], _g(gold, leg, k)), (leg, k)
            if k in RULE_KEYS and res[k].any():
                seen.add(k.replace("_this_step", ""))
        assert (res["preds"] - gold[f"{leg}/preds"]).abs().max() <= 5e-4, leg
        assert (res["diffbar_rewards"] - gold[f"{leg}/diffbar_rewards"]).abs().max() <= 5e-4, leg
    assert {"collided", "run_red_light"} <= seen  # the fixtures really contain events


@pytest.mark.parametrize("name", ["default", "mtr_nms", "mpa_nms", "mtr_mpa_fde", "a64_k6"])
def test_post_oracle_matches_reference_golden(name):
    """`oracle/post_oracle.py` reproduces what the reference's `WaymoPostProcessing.forward` and
    `WOMDMetrics.update` produced for the seeded multi-modal trajectories (`oracle/make_golden.py`, POST_CASES)."""
    import os

    import numpy as np

    import post_oracle as po
    from golden_util import GOLDEN_DIR
    from trafficbots_b200 import synthetic
    z = np.load(os.path.join(GOLDEN_DIR, "post_cases.npz"))
    S, A, n, seed, tseed, ade = [int(x) for x in z[f"{name}__meta"]]
    batch = synthetic.make_batch(S, n_agent=A, n_pl=16, seed=seed)
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, n, seed=tseed)
    got = po.post_process(valid, scores, trajs, batch["agent/type"], 6, 1e2, list(z[f"{name}__mpa"]), list(z[f"{name}__mtr"]), bool(ade))
    for k in ("waymo_trajs", "waymo_yaw_bbox", "waymo_spd"):
        assert torch.equal(got[k], torch.from_numpy(z[f"{name}__{k}"])), k
    assert (got["waymo_scores"] - torch.from_numpy(z[f"{name}__waymo_scores"])).abs().max() <= 1e-7
    packed = po.womd_pack(batch, got["waymo_trajs"], got["waymo_scores"])
    for k, v in packed.items():
        w = torch.from_numpy(z[f"{name}__{k}"])
        assert (torch.equal(v, w) if v.dtype == torch.bool else (v - w).abs().max() <= 1e-7), k

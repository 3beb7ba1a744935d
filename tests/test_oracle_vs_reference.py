"""Build container only: pins the oracle against the UNMODIFIED reference executed live (fresh seeds, i.e. inputs
that are not in the golden fixtures), including test-mode (11 GT frames) rollouts."""
import pytest
import torch

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("seed,k", [(321, 1), (654, 3)])
def test_oracle_vs_live_reference(seed, k):
    import ref_loader
    import ref_run
    import trafficbots_oracle as orc
    from trafficbots_b200 import synthetic, weights

    A, P, S = 8, 48, 2
    model = ref_loader.build_reference(n_agent=A, n_pl=P, n_joint_future=k)
    sd = weights.init_state_dict(seed)
    model.load_state_dict(sd, strict=True)
    batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed, special_scenes=False)
    ref = ref_run.run_reference(model, batch, k_futures=k, sample_seed=3)
    feat = orc.encode_scene(sd, batch)
    jfp = orc.joint_future_pred(sd, batch, k=k, sample_seed=3, feat=feat)
    assert (feat["map_feature"] - ref["enc/map_feature"]).abs().max() <= 1e-6
    assert torch.equal(jfp["goal_sample"], ref["jfp/goal_sample"])
    assert torch.equal(jfp["valid"], ref["jfp/valid"])
    assert (jfp["preds"] - ref["jfp/preds"]).abs().max() <= 5e-4
    for key in ("outside_map", "goal_reached", "dest_reached"):
        assert torch.equal(jfp[key], ref[f"jfp/violations/{key}"])


def test_state_dict_schema_matches_reference():
    import ref_loader
    from trafficbots_b200 import weights

    model = ref_loader.build_reference(n_agent=8, n_pl=32)
    ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    spec = weights.state_dict_spec()
    assert list(ref) == list(spec)
    assert all(ref[k] == tuple(spec[k]) for k in ref)
    assert weights.count_parameters(weights.init_state_dict(0)) == sum(p.numel() for p in model.parameters())


@pytest.mark.parametrize("variant", ["default", "mtr_nms", "mpa_nms", "mtr+mpa_fde"])
def test_post_processing_oracle_vs_live_reference(variant):
    """`oracle/post_oracle.py` against the reference's own `WaymoPostProcessing` and `WOMDMetrics.update`."""
    import post_oracle as po
    import ref_loader
    from trafficbots_b200 import synthetic
    ref_loader.install_stubs()
    from data_modules.waymo_post_processing import WaymoPostProcessing
    from models.metrics.womd import WOMDMetrics

    cfg = {"default": dict(n=6, mtr=[], mpa=[], ade=True), "mtr_nms": dict(n=14, mtr=[2.5, 1.0, 1.5], mpa=[], ade=True),
           "mpa_nms": dict(n=6, mtr=[], mpa=[2.5, 1.0, 1.5], ade=True), "mtr+mpa_fde": dict(n=10, mtr=[3.0, 1.5, 2.0], mpa=[4.0, 2.0, 3.0], ade=False)}[variant]
    S, A = 3, 12
    batch = synthetic.make_batch(S, n_agent=A, n_pl=16, seed=77)
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, cfg["n"], seed=5)
    pp = WaymoPostProcessing(k_pred=6, score_temperature=1e2, mpa_nms_thresh=cfg["mpa"], mtr_nms_thresh=cfg["mtr"], aggr_thresh=[],
                             n_iter_em=3, use_ade=cfg["ade"])
    ref = pp(valid=valid, scores=scores.clone(), trajs=trajs.clone(), agent_type=batch["agent/type"])
    got = po.post_process(valid, scores, trajs, batch["agent/type"], 6, 1e2, cfg["mpa"], cfg["mtr"], cfg["ade"])
    for k in ("waymo_trajs", "waymo_yaw_bbox", "waymo_spd", "waymo_valid"):
        assert torch.equal(got[k], ref[k]), k
    assert (got["waymo_scores"] - ref["waymo_scores"]).abs().max() <= 1e-7
    m = WOMDMetrics("val", step_gt=90, step_current=10, interactive_challenge=False)
    m.update(batch, ref["waymo_trajs"], ref["waymo_scores"])
    mine = po.womd_pack(batch, got["waymo_trajs"], got["waymo_scores"])
    for k, v in mine.items():
        assert torch.equal(v, getattr(m, k + "_gpu")[0]), k
    m2 = WOMDMetrics("val", step_gt=90, step_current=10, interactive_challenge=False)
    m2.update(batch, ref["waymo_trajs"], None)
    assert torch.equal(po.womd_pack(batch, got["waymo_trajs"], None)["prediction_score"], m2.prediction_score_gpu[0])

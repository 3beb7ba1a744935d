"""GPU: the eval-loop drop-in.

* `validation_step` / `test_step` of the b200 shell against the golden vectors of the reference.
* marked `reference` (needs the reference sources: /root/reference here, baseline/_ref/src on the GPU box): the reference's OWN
  `WaymoMotion.validation_step` / `test_step` function bodies are executed with the b200 module as `self` -- the exact
  call sequence of the reference's training / eval loop (pre_processing -> 3 x encode_input_features(**dict) -> get_gt_goal /
  pred_goal -> latent_encoder x 2 -> reactive_replay -> metrics -> waymo_post_processing -> womd_metrics ->
  joint_future_pred -> ...) drives the CUDA library without a line of the reference being changed.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(sd, K):
    from trafficbots_b200 import config
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    m = WaymoMotion(**config.default_config(n_joint_future=K))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _cuda_batch(batch):
    cb = {k: v.cuda() for k, v in batch.items()}
    S = batch["map/valid"].shape[0]
    cb["scenario_center"] = torch.zeros(S, 2, device="cuda")
    cb["scenario_yaw"] = torch.zeros(S, device="cuda")
    cb["scenario_id"] = torch.arange(S, device="cuda")
    return cb


@pytest.mark.parametrize("case", ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2"])
def test_validation_step_matches_reference_golden(case):
    from golden_util import load_case
    gold, sd, batch, meta = load_case(case)
    m = _module(sd, meta["K"])
    torch.manual_seed(0)
    out = m.validation_step(_cuda_batch(batch), 0)
    rep, jfp = out["reactive_replay"], out["joint_future_pred"]
    assert torch.equal(rep.valid.squeeze(2).cpu(), gold["replay/valid"])
    assert float((rep.preds.squeeze(2).cpu() - gold["replay/preds"]).abs().max()) <= 2e-3  # closed-loop tolerance (test_gpu_parity.py)
    assert float((out["latent_post"].mean.cpu() - gold["latent_post/mean"]).abs().max()) <= 1e-4
    K = meta["K"]  # joint_future_pred repeats the prior / destination distributions per mode IN PLACE, like the reference (:493-496)
    assert float((out["latent_prior"].mean[::K].cpu() - gold["latent_prior/mean"]).abs().max()) <= 1e-4
    assert float((out["goal_pred"].probs[::K].cpu() - gold["dest/probs"]).abs().max()) <= 1e-4
    assert torch.equal(jfp.valid[:, :, 0].cpu(), gold["jfp/valid"][:, :, 0])  # mode 0 is deterministic
    assert float((jfp.preds[:, :, 0].cpu() - gold["jfp/preds"][:, :, 0]).abs().max()) <= 2e-3
    # post-processing + WOMD records of both legs exist and decode
    S, A = batch["agent/type"].shape[:2]
    six = m.womd_metrics_joint_future_pred.compute()
    assert six["prediction_trajectory"][0].shape == (S, 8, meta["K"], 1, 16, 2)
    assert six["ground_truth_trajectory"][0].shape == (S, A, 91, 7)
    sc = out["pred_dict_joint_future_pred"]["waymo_scores"]
    assert float((sc.sum(-1) - 1).abs().max()) <= 1e-5
    # the aliased encodes were served from the first one (one map encode per step, not three)
    assert out["reactive_replay"] is not None and m.engine().lib.tb_launch_count() > 0


def test_test_step_history_only():
    from golden_util import load_case
    import trafficbots_oracle as orc
    gold, sd, batch, meta = load_case("s3_a8_p64_k2")
    m = _module(sd, meta["K"])
    hist = {k: v for k, v in batch.items() if k.startswith(("history/", "map/"))}  # what a test-set batch holds
    torch.manual_seed(0)
    out = m.test_step(_cuda_batch(hist), 0)
    ref = orc.joint_future_pred(sd, batch, k=meta["K"], sample_seed=0, test_mode=True)
    buf = out["joint_future_pred"]
    assert torch.equal(buf.valid[:, :, 0].cpu(), ref["valid"][:, :, 0])
    assert float((buf.preds[:, :, 0].cpu() - ref["preds"][:, :, 0]).abs().max()) <= 2e-3  # closed-loop tolerance
    assert out["pred_dict"]["waymo_trajs"].shape == (meta["S"], 80, meta["A"], meta["K"], 2)


@pytest.mark.reference
@pytest.mark.parametrize("case", ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2"])
def test_reference_validation_step_body_drives_the_library(case):
    import ref_loader
    from golden_util import load_case
    ref_loader.install_stubs()
    from pl_modules.waymo_motion import WaymoMotion as RefWaymoMotion  # the reference's own class (unmodified source)
    gold, sd, batch, meta = load_case(case)
    m = _module(sd, meta["K"])
    torch.manual_seed(0)
    mine = m.validation_step(_cuda_batch(batch), 99)
    m2 = _module(sd, meta["K"])
    torch.manual_seed(0)
    RefWaymoMotion.validation_step(m2, _cuda_batch(batch), 99)  # batch_idx >= n_video_batch: no videos
    # the reference's body pushed its keyword tensors into the module's recorders: identical to the shell's own step
    args, kw = m2.err_metrics_joint_future_pred.last
    assert torch.equal(kw["pred_states"], mine["joint_future_pred"].preds)
    assert torch.equal(kw["pred_valid"], mine["joint_future_pred"].valid)
    args, kw = m2.err_metrics_reactive_replay.last
    assert torch.equal(kw["pred_states"], mine["reactive_replay"].preds)
    assert float((kw["pred_states"].squeeze(2).cpu() - gold["replay/preds"]).abs().max()) <= 2e-3  # closed-loop tolerance
    args, kw = m2.rule_metrics_joint_future_pred.last
    assert torch.equal(kw["dest_reached"][:, :, 0].cpu(), gold["jfp/violations/dest_reached"][:, :, 0])
    args, kw = m2.sub_womd_joint_future_pred.last
    assert torch.equal(kw["waymo_trajs"], mine["pred_dict_joint_future_pred"]["waymo_trajs"])
    assert torch.equal(m2.womd_metrics_joint_future_pred.ops_inputs_cpu[0], mine["womd_records_joint_future_pred"].cpu())
    assert m2.womd_metrics_joint_future_pred.records == []  # the reference's loop reset the per-step device state
    assert m2.train_metrics_reactive_replay.n_call == 1


@pytest.mark.reference
def test_reference_test_step_body_drives_the_library():
    import ref_loader
    from golden_util import load_case
    ref_loader.install_stubs()
    from pl_modules.waymo_motion import WaymoMotion as RefWaymoMotion
    gold, sd, batch, meta = load_case("s3_a8_p64_k2")
    m = _module(sd, meta["K"])
    cb = _cuda_batch({k: v for k, v in batch.items() if k.startswith(("history/", "map/"))})
    torch.manual_seed(0)
    mine = m.test_step(dict(cb), 0)
    m2 = _module(sd, meta["K"])
    torch.manual_seed(0)
    RefWaymoMotion.test_step(m2, dict(cb), 0)
    args, kw = m2.sub_womd_joint_future_pred.last
    assert torch.equal(kw["waymo_trajs"], mine["pred_dict"]["waymo_trajs"])
    assert torch.equal(kw["waymo_scores"], mine["pred_dict"]["waymo_scores"])

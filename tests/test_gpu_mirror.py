"""GPU: the reference-shaped Python surface (`WaymoMotion.joint_future_pred / reactive_replay / forward`) drives the CUDA
library and reproduces the oracle / golden vectors.  The first tests take the prior latent and the destination distribution
from the oracle (as the distribution objects the reference's methods expect); the last one runs the library's own
pre-rollout heads (`model.latent_encoder`, `model.goal_manager.pred_goal`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(sd, K):
    from trafficbots_b200 import config
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    m = WaymoMotion(**config.default_config(n_joint_future=K))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _heads(sd, batch):
    import trafficbots_oracle as orc
    from trafficbots_b200.models.distributions import DestCategorical, DiagGaussian
    feat = orc.encode_scene(sd, batch)
    prior = orc.latent_encoder(sd, feat)
    dest = orc.dest_predictor(sd, feat, batch["agent/type"], batch["map/type"])
    log_std = sd["model.latent_encoder.latent_prior_dist.log_std"].cuda()
    return (DiagGaussian(prior["mean"].cuda(), log_std, prior["valid"].cuda()),
            DestCategorical(probs=dest["probs"].cuda(), valid=dest["valid"].cuda()))


@pytest.mark.parametrize("case", ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2"])
def test_joint_future_pred_surface(case):
    from golden_util import load_case
    gold, sd, batch, meta = load_case(case)
    K, S = meta["K"], meta["S"]
    m = _module(sd, K)
    cb = {k: v.cuda() for k, v in batch.items()}
    feat = m.model.encode_input_features(cb)
    latent, goal = _heads(sd, batch)
    goal_valid = cb["history/agent/valid"].any(1)
    torch.manual_seed(0)
    buf, goal_sample, goal_logp = m.joint_future_pred(cb, feat, latent, goal, goal_valid, require_vis_dict=False)
    A = batch["agent/type"].shape[1]
    assert buf.preds.shape == (S, A, K, 90, 4) and buf.valid.shape == (S, A, K, 90)
    assert goal_sample.shape == (S, A, K) and goal_logp.shape == (S, A, K)
    assert len(buf.violations) == 14 and not buf.violations["collided"].any()
    # mode 0 is deterministic (prior mean, arg-max destination): identical to the reference's mode 0
    assert torch.equal(goal_sample[:, :, 0].cpu(), gold["jfp/goal_sample"][:, :, 0])
    assert torch.equal(buf.valid[:, :, 0].cpu(), gold["jfp/valid"][:, :, 0])
    assert float((buf.preds[:, :, 0].cpu() - gold["jfp/preds"][:, :, 0]).abs().max()) <= 2e-3  # closed-loop tolerance, see tests/test_gpu_parity.py
    for k in ("outside_map", "goal_reached", "dest_reached"):
        assert torch.equal(buf.violations[k][:, :, 0].cpu(), gold[f"jfp/violations/{k}"][:, :, 0]), k
    assert float((m.model.hidden.view(3, S, K, A, 128)[:, :, 0].cpu()
                  - gold["jfp/hidden"].view(3, S, K, A, 128)[:, :, 0]).abs().max()) <= 2e-3  # closed-loop tolerance, see tests/test_gpu_parity.py


def test_reactive_replay_surface_and_stepwise_forward():
    import trafficbots_oracle as orc
    from golden_util import load_case
    from trafficbots_b200 import host
    gold, sd, batch, meta = load_case("s3_a8_p64_k2")
    m = _module(sd, 1)
    cb = {k: v.cuda() for k, v in batch.items()}
    feat = m.model.encode_input_features(cb)
    post = gold["latent_post/mean"].cuda()  # posterior mean from the reference (deterministic latent)
    goal_valid = cb["history/agent/valid"].any(1)
    tf = m.teacher_forcing_reactive_replay.get(cb["agent/valid"], 0)
    assert torch.equal(tf.cpu(), orc.teacher_forcing_mask(batch["agent/valid"], 90, 10))
    buf = m.reactive_replay(cb, feat, tf, post, cb["agent/dest"], goal_valid, deterministic_latent=True,
                            deterministic_action=True, require_vis_dict=False)
    assert torch.equal(buf.valid.cpu(), gold["replay/valid"])
    assert torch.equal(buf.override_masks.cpu(), gold["replay/override_masks"])
    assert float((buf.preds.cpu() - gold["replay/preds"]).abs().max()) <= 2e-3  # closed-loop tolerance, see tests/test_gpu_parity.py
    assert float((buf.diffbar_rewards.cpu() - gold["replay/diffbar_rewards"]).abs().max()) <= 2e-3  # closed-loop tolerance, see tests/test_gpu_parity.py
    for k in ("outside_map", "goal_reached", "dest_reached", "dest_reached_this_step"):
        assert torch.equal(buf.violations[k].cpu(), gold[f"replay/violations/{k}"]), k
    one_shot = buf.preds.clone()
    # the same rollout driven step by step through forward()
    from trafficbots_b200.pl_modules.waymo_motion import TrafficRuleChecker
    rc = TrafficRuleChecker(cb["map/boundary"], cb["map/valid"], cb["map/type"], cb["map/pos"], cb["map/dir"],
                            agent_goal=cb["agent/goal"], agent_dest=cb["agent/dest"])
    feats = m._features(cb, feat, 1)
    feats["_stepwise"] = True
    m.rollout(feats, post, cb["agent/dest"], goal_valid, tf, rc, True, True, step_end=90, step_start=1)
    for t in range(1, 91):
        state, valid, train, vis = m.forward()
        assert state.shape == (3, 8, 4) and valid.shape == (3, 8)
        if t == 37:
            assert torch.equal(train["pred_state"], one_shot[:, :, 36])
    buf2 = m.finish_rollout()
    assert torch.equal(buf2.preds, one_shot)


def test_repack_after_parameter_update():
    from golden_util import load_case
    gold, sd, batch, meta = load_case("cfg1_s1_a8_p64_k1")
    m = _module(sd, 1)
    cb = {k: v.cuda() for k, v in batch.items()}
    f1 = m.model.encode_input_features(cb)["map_feature"].clone()
    with torch.no_grad():
        getattr(m.model.map_encoder.transformer_self_attn.layers, "0").linear2.bias.add_(1.0)
    f2 = m.model.encode_input_features(cb)["map_feature"]
    valid = gold["enc/map_feature_valid"].cuda()
    assert float(((f2 - f1)[valid] - 1.0).abs().max()) <= 1e-5  # the bias shift shows up: weights were re-packed


@pytest.mark.parametrize("case", ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2"])
def test_full_pipeline_with_library_heads(case):
    """validation_step's joint_future_pred leg entirely on the library: encode_input_features -> model.latent_encoder ->
    model.goal_manager.pred_goal -> joint_future_pred (reference src/pl_modules/waymo_motion.py:581-598); the deterministic
    mode 0 (prior mean, arg-max destination) reproduces the unmodified reference."""
    from golden_util import load_case
    gold, sd, batch, meta = load_case(case)
    K, S = meta["K"], meta["S"]
    m = _module(sd, K)
    cb = {k: v.cuda() for k, v in batch.items()}
    feat = m.model.encode_input_features(cb)
    latent = m.model.latent_encoder(**feat)
    goal = m.model.goal_manager.pred_goal(agent_type=cb["agent/type"], map_type=cb["map/type"], agent_state=None, **feat)
    assert float((latent.mean.cpu() - gold["latent_prior/mean"]).abs().max()) <= 1e-4
    assert float((goal.probs.cpu() - gold["dest/probs"]).abs().max()) <= 2e-5
    goal_valid = cb["history/agent/valid"].any(1)
    torch.manual_seed(0)
    buf, goal_sample, _ = m.joint_future_pred(cb, feat, latent, goal, goal_valid, require_vis_dict=False)
    assert torch.equal(goal_sample[:, :, 0].cpu(), gold["jfp/goal_sample"][:, :, 0])
    assert torch.equal(buf.valid[:, :, 0].cpu(), gold["jfp/valid"][:, :, 0])
    assert float((buf.preds[:, :, 0].cpu() - gold["jfp/preds"][:, :, 0]).abs().max()) <= 2e-3  # closed-loop tolerance


@pytest.mark.gpu
def test_scene_stager_round_trip():
    """host.SceneStager: a staged batch equals a direct copy, results read back behind the compute stream arrive intact."""
    from trafficbots_b200 import host
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    batch = host.pin_batch({"a": torch.randn(5, 7, generator=g), "b": torch.randint(0, 2, (4, 3), generator=g).bool()})
    st = host.SceneStager(dev)
    with pytest.raises(RuntimeError):
        st.get()
    st.submit(batch)
    cb = st.get()
    st.submit(batch)  # next step's batch travels while this one is used
    y = cb["a"] * 2.0 + cb["b"].float().sum()
    out = torch.empty(5, 7).pin_memory()
    st.read_back([(out, y)])
    st.join()
    torch.cuda.synchronize()
    assert torch.equal(out, batch["a"] * 2.0 + batch["b"].float().sum())
    cb2 = st.get()
    torch.cuda.synchronize()
    assert torch.equal(cb2["a"].cpu(), batch["a"]) and torch.equal(cb2["b"].cpu(), batch["b"])


def test_sampled_modes_through_the_public_surface_match_reference():
    """All K modes -- not only the deterministic mode 0 -- of `joint_future_pred` on the public surface against the reference's
    golden rollout: the CUDA generator cannot reproduce the reference's CPU draws, so the prior / destination objects handed to
    `joint_future_pred` return the reference's own samples (`jfp/latent_sample`, `jfp/goal_sample` of the fixture) from
    `.sample()`; everything downstream (log-probs, K-mode replication sharing the scene's key blocks, rollout, buffers) is the
    product path."""
    from golden_util import load_case
    from trafficbots_b200.models.distributions import DestCategorical, DiagGaussian
    gold, sd, batch, meta = load_case("s3_a8_p64_k2")
    K, S = meta["K"], meta["S"]
    A = batch["agent/type"].shape[1]
    m = _module(sd, K)
    cb = {k: v.cuda() for k, v in batch.items()}
    feat = m.model.encode_input_features(cb)
    latent, goal = _heads(sd, batch)

    class FixedLatent(DiagGaussian):
        def sample(self, deterministic):
            return gold["jfp/latent_sample"].cuda()  # [S*K, A, 16], scene-major like the reference's repeat_interleave

    class FixedGoal(DestCategorical):
        def sample(self, deterministic):
            return gold["jfp/goal_sample"].transpose(1, 2).reshape(S * K, A).cuda()  # fixture layout [S, A, K]

    latent.__class__, goal.__class__ = FixedLatent, FixedGoal
    goal_valid = cb["history/agent/valid"].any(1)
    buf, goal_sample, goal_logp = m.joint_future_pred(cb, feat, latent, goal, goal_valid, require_vis_dict=False)
    assert torch.equal(goal_sample.cpu(), gold["jfp/goal_sample"])
    assert float((goal_logp.cpu() - gold["jfp/goal_log_probs"]).abs().max()) <= 1e-4
    for k in range(K):  # every mode, sampled ones included
        assert torch.equal(buf.valid[:, :, k].cpu(), gold["jfp/valid"][:, :, k]), k
        assert torch.equal(buf.override_masks[:, :, k].cpu(), gold["jfp/override_masks"][:, :, k]), k
        assert float((buf.preds[:, :, k].cpu() - gold["jfp/preds"][:, :, k]).abs().max()) <= 2e-3, k  # closed-loop tolerance
        assert float((buf.latent_log_probs[:, :, k].cpu() - gold["jfp/latent_log_probs"][:, :, k]).abs().max()) <= 1e-3, k
        for name in ("outside_map", "goal_reached", "dest_reached"):
            assert torch.equal(buf.violations[name][:, :, k].cpu(), gold[f"jfp/violations/{name}"][:, :, k]), (name, k)
    # the sampled mode really differs from the deterministic one
    assert float((buf.preds[:, :, 1] - buf.preds[:, :, 0]).abs().max()) > 1e-2

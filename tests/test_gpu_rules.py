"""GPU: the optional traffic-rule checks and the collision reward (`tb_rule_checks`) against the golden vectors
the UNMODIFIED reference produced with all four `enable_check_*` flags and `w_collision` switched on, and against the oracle.

Two levels:
  * kernel level: `tb_rule_checks` fed with the REFERENCE's own rollout outputs (preds / valid / override_masks /
    outside_map_this_step from the fixture) must reproduce the reference's 8 optional violation maps bit for bit -- the
    post-pass reconstruction of the post-override state is exact, and so is the thresholded geometry;
  * end to end through the `WaymoMotion` surface on the library's own rollout: boolean maps compared entry by entry (a closed
    loop that differs by <= 2e-3 m can flip a threshold test that is decided by less than that, so a mismatch budget of 0.2 %
    of the entries is allowed and the measured count is printed), rewards within the closed-loop tolerance.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(sd, K, meta):
    from trafficbots_b200 import config
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    m = WaymoMotion(**config.default_config(n_joint_future=K, rule_checks=True, w_collision=meta["w_collision"],
                                            reduce_collision_with_max=meta["reduce_with_max"]))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _flat(x, S, K):
    """[S,A,K,T,..] (flatten_repeat layout of the fixtures) -> [S*K,A,T,..] contiguous."""
    return x.transpose(1, 2).reshape(S * K, *x.shape[1:2], *x.shape[3:]).contiguous()


@pytest.mark.parametrize("case", ["s2_a16_p96_k2_rules", "s2_a12_p64_k1_rules_sum"])
@pytest.mark.parametrize("leg", ["jfp", "replay"])
def test_rule_kernels_on_reference_rollout_bit_exact(case, leg):
    from golden_util import RULE_KEYS, RULES_ON, load_case
    from trafficbots_b200 import engine as E, weights
    gold, sd, batch, meta = load_case(case)
    S, K = meta["S"], (meta["K"] if leg == "jfp" else 1)
    eng = E.Engine(sd, "cuda")
    cb = {k: v.cuda() for k, v in batch.items()}
    if leg == "jfp":
        g = lambda k: _flat(gold[f"jfp/{k}"], S, K).cuda()  # noqa: E731
        tl = {k: cb[f"history/tl_stop/{k}"] for k in ("valid", "pos", "state")}
    else:
        g = lambda k: gold[f"replay/{k}"].contiguous().cuda()  # noqa: E731
        tl = {k: cb[f"tl_stop/{k}"] for k in ("valid", "pos", "state")}
    # the reference's reward WITHOUT its collision term is not in the fixture: check the booleans here (reward: next test)
    out = {"preds": g("preds"), "valid": g("valid"), "override_masks": g("override_masks"),
           "violations/outside_map_this_step": g("violations/outside_map_this_step"),
           "diffbar_rewards": torch.zeros_like(g("diffbar_rewards")), "diffbar_rewards_valid": g("diffbar_rewards_valid")}
    res = eng.rule_checks(out, E.gt_from_batch(cb), cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), tl, RULES_ON,
                          n_mode=K)
    torch.cuda.synchronize()
    n_event = 0
    for k in RULE_KEYS:
        want = g(f"violations/{k}")
        got = res[f"violations/{k}"]
        assert torch.equal(got, want), (k, int((got != want).sum()), int(want.sum()))
        n_event += int(want.sum())
    assert n_event > 0


@pytest.mark.parametrize("case", ["s2_a16_p96_k2_rules", "s2_a12_p64_k1_rules_sum"])
def test_rule_checks_end_to_end_surface(case):
    from golden_util import RULE_KEYS, load_case
    gold, sd, batch, meta = load_case(case)
    S, K = meta["S"], meta["K"]
    m = _module(sd, 1, meta)
    cb = {k: v.cuda() for k, v in batch.items()}
    feat = m.model.encode_input_features(cb)
    goal_valid = cb["history/agent/valid"].any(1)
    tf = m.teacher_forcing_reactive_replay.get(cb["agent/valid"], 0)
    buf = m.reactive_replay(cb, feat, tf, gold["latent_post/mean"].cuda(), cb["agent/dest"], goal_valid, deterministic_latent=True,
                            deterministic_action=True, require_vis_dict=False)
    torch.cuda.synchronize()
    assert torch.equal(buf.valid.cpu(), gold["replay/valid"])
    assert float((buf.preds.cpu() - gold["replay/preds"]).abs().max()) <= 2e-3  # closed-loop tolerance (tests/test_gpu_parity.py)
    assert float((buf.diffbar_rewards.cpu() - gold["replay/diffbar_rewards"]).abs().max()) <= 5e-3  # incl. the collision term
    total = mism = events = 0
    for k in RULE_KEYS:
        want = gold[f"replay/violations/{k}"]
        got = buf.violations[k].cpu()
        mism += int((got != want).sum())
        total += want.numel()
        events += int(want.sum())
    print(f"{case}: optional-check maps: {mism} of {total} entries differ from the reference ({events} events)")
    assert events > 0 and mism <= 0.002 * total
    # the collision term is really in the reward: it differs from an IL-only run
    m0 = _module(sd, 1, dict(meta, w_collision=0.0))
    f0 = m0.model.encode_input_features(cb)
    buf0 = m0.reactive_replay(cb, f0, tf, gold["latent_post/mean"].cuda(), cb["agent/dest"], goal_valid, deterministic_latent=True,
                              deterministic_action=True, require_vis_dict=False)
    assert float((buf0.diffbar_rewards - buf.diffbar_rewards).abs().max()) > 1e-3


def test_rule_checks_vs_oracle_k6_shape():
    """config-2-like K = 3 joint futures on a denser, larger scene: CUDA post-pass == oracle's per-step evaluation on the
    oracle's own rollout (kernel level, bit exact), history traffic lights frozen after frame 10."""
    import trafficbots_oracle as orc
    from trafficbots_b200 import engine as E, synthetic, weights
    sd = weights.init_state_dict(5)
    S, A, P, K = 2, 24, 160, 3
    batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=77, special_scenes=False, area_scale=0.3, plant_red_light=True)
    on = {"collided": True, "run_road_edge": True, "run_red_light": True, "passive": True}
    ref = orc.joint_future_pred(sd, batch, k=K, sample_seed=3, rules_enable=on, w_collision=0.7)
    ref0 = orc.joint_future_pred(sd, batch, k=K, sample_seed=3)  # IL-only reward
    eng = E.Engine(sd, "cuda")
    cb = {k: v.cuda() for k, v in batch.items()}
    g = lambda r, k: _flat(r[k], S, K).cuda()  # noqa: E731
    out = {"preds": g(ref, "preds"), "valid": g(ref, "valid"), "override_masks": g(ref, "override_masks"),
           "violations/outside_map_this_step": g(ref, "outside_map_this_step"),
           "diffbar_rewards": g(ref0, "diffbar_rewards"), "diffbar_rewards_valid": g(ref, "diffbar_rewards_valid")}
    tl = {k: cb[f"history/tl_stop/{k}"] for k in ("valid", "pos", "state")}
    res = eng.rule_checks(out, E.gt_from_batch(cb), cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), tl, on, n_mode=K,
                          w_collision=0.7)
    torch.cuda.synchronize()
    for k in ("collided", "run_road_edge", "run_red_light", "passive"):
        for sfx in ("", "_this_step"):
            assert torch.equal(res[f"violations/{k}{sfx}"], g(ref, k + sfx)), k + sfx
    assert float((res["diffbar_rewards"] - g(ref, "diffbar_rewards")).abs().max()) <= 1e-5


def test_passive_without_red_light_is_rejected():
    from trafficbots_b200 import config
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    cfg = config.default_config()
    cfg["traffic_rule_checker"] = {"enable_check_passive": True, "enable_check_run_red_light": False}
    with pytest.raises(config.UnsupportedConfig):
        WaymoMotion(**cfg)

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

"""GPU parity of the pre-rollout heads through the C ABI: prior / posterior latent encoder
(`tb_xlayer` + `tb_kv_project` + `tb_gru_sequence` + `tb_mlp_head`) and the destination predictor (`tb_gru_sequence` +
`tb_dest_logits`) against the CPU oracle and the golden vectors of the unmodified reference.  fp32 kernels: 1e-4 on the
latent mean (O(1) values), 2e-5 on the destination probabilities; masked-out destinations are exactly zero."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2", "s1_a64_p1024_k1"]


def _setup(case):
    import trafficbots_oracle as orc
    from golden_util import load_case
    from trafficbots_b200.engine import Engine
    gold, sd, batch, meta = load_case(case)
    eng = Engine(sd, "cuda")
    cb = {k: v.cuda() for k, v in batch.items()}
    return orc, gold, sd, batch, cb, eng


@pytest.mark.parametrize("case", CASES)
def test_latent_prior_matches_oracle_and_golden(case):
    orc, gold, sd, batch, cb, eng = _setup(case)
    feat = eng.encode_scene(cb)
    mean, valid = eng.latent_encoder(feat)
    ref = orc.latent_encoder(sd, orc.encode_scene(sd, batch))
    assert torch.equal(valid.cpu(), ref["valid"])
    assert float((mean.cpu() - ref["mean"]).abs().max()) <= 1e-4
    assert float((mean.cpu() - gold["latent_prior/mean"]).abs().max()) <= 1e-4


@pytest.mark.parametrize("case", CASES[:2])
def test_latent_posterior_matches_golden(case):
    orc, gold, sd, batch, cb, eng = _setup(case)
    feat_post = eng.encode_scene(cb, prefix="")  # the full 91-frame episode (data_modules/sc_latent.py:166-168,211-236)
    mean, valid = eng.latent_encoder(feat_post, posterior=True)
    assert torch.equal(valid.cpu(), batch["agent/valid"].any(1))
    assert float((mean.cpu() - gold["latent_post/mean"]).abs().max()) <= 1e-4


@pytest.mark.parametrize("case", CASES)
def test_dest_predictor_matches_oracle_and_golden(case):
    orc, gold, sd, batch, cb, eng = _setup(case)
    feat = eng.encode_scene(cb)
    probs, logp, valid = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])
    ref = orc.dest_predictor(sd, orc.encode_scene(sd, batch), batch["agent/type"], batch["map/type"])
    assert torch.equal(valid.cpu(), ref["valid"])
    p, lp = probs.cpu(), logp.cpu()
    assert torch.equal(p == 0, ref["probs"] == 0)  # type masks: the same destinations are excluded
    assert float((p - ref["probs"]).abs().max()) <= 2e-5
    assert float((p - gold["dest/probs"]).abs().max()) <= 2e-5
    fin = torch.isfinite(ref["logp"])
    assert torch.equal(torch.isfinite(lp), fin)
    assert float((lp[fin] - ref["logp"][fin]).abs().max()) <= 2e-4
    assert float((p.sum(-1) - 1).abs().max()) <= 1e-5

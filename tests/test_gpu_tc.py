"""GPU: the tcgen05 building blocks.  bf16x3 split GEMM on the tensor pipe vs an fp64 product: relative error of the
split (dropped lo*lo term, rounded lo parts) is bounded by ~2^-16 of |a|.|w| per output."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
def test_tc_selftest_gemm_matches_fp64(mode):
    from trafficbots_b200 import _native as nt, weights
    from trafficbots_b200.engine import Engine
    sd = weights.init_state_dict(4)
    eng = Engine(sd, "cuda")
    lib = eng.lib
    names = [lib.tb_weight_name(i).decode() for i in range(lib.tb_weight_count())]
    g = torch.Generator().manual_seed(0)
    for key, nb, kb in (("model.transformer_as2pl.layers.0.linear1.weight", 0, 0),
                        ("model.transformer_as2pl.layers.1.attn.in_proj_weight", 2, 0),
                        ("model.add_goal.mlp_out.fc_layers.0.weight", 0, 1),
                        ("model.agent_temporal.rnn.weight_hh_l2", 1, 0)):
        wi = names.index(key)
        first = lib.tb_tc_first_block(wi)
        assert first >= 0
        w = sd[key]
        blk = first + nb * (w.shape[1] // 128) + kb
        a = (torch.randn(128, 128, generator=g) * 3).cuda()
        d = torch.empty(128, 128, device="cuda")
        nt.check(lib.tb_tc_selftest(a.data_ptr(), blk, eng.packed.data_ptr(), d.data_ptr(), mode, nt.current_stream_ptr()), "selftest")
        torch.cuda.synchronize()
        wsub = w[nb * 128:(nb + 1) * 128, kb * 128:(kb + 1) * 128].double()
        ref = a.cpu().double() @ wsub.t()
        bound = (a.cpu().double().abs() @ wsub.abs().t()) * 2.0 ** -15 + 1e-6
        err = (d.cpu().double() - ref).abs()
        assert bool((err <= bound).all()), (key, float(err.max()), float((err / bound).max()))
        assert float(err.max()) <= 2e-4

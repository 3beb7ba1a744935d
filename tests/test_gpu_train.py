"""Training step (BASELINE.json configs[3]) -- GPU parity tests, through the C ABI (`tb_tr_*`).

(1) every primitive: forward kernel vs the torch restatement, hand-derived backward kernel vs torch.autograd of that
    restatement (`oracle/train_ops_oracle.py`), fp32, tolerance 2e-5 relative to the tensor's max;
(2) the whole step: loss terms and gradients of all 385 parameters vs fingerprints of the UNMODIFIED reference's
    `training_step` + backward (`tests/golden/train_*.npz`), tolerance 2e-3 of each gradient's scale;
(3) a few optimizer steps reduce the loss and match the same steps done with the torch restatement on CPU.
"""
import pytest
import torch

from train_ops_oracle import OracleOps
from train_util import TRAIN_CASES, compare_grads, load_train_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from trafficbots_b200.train.cuda_ops import CudaOps
    return CudaOps(DEV)


ORC = OracleOps()


def close(a, b, tol=2e-5, what="", atol=2e-6):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= tol * scale + atol, f"{what}: max|diff| {err:.3e} vs scale {scale:.3e}"


def g(*t):
    return [x.to(DEV) if x is not None else None for x in t]


@pytest.mark.parametrize("M,K,N,relu,strided", [(37, 11, 32, True, False), (300, 128, 256, False, False), (1000, 256, 128, True, True),
                                                (129, 128, 1, False, False), (64, 16, 128, True, False), (5000, 128, 384, False, False)])
def test_linear(ops, M, K, N, relu, strided):
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K)
    wfull = torch.randn(N, 2 * K if strided else K) / K ** 0.5
    w = wfull[:, K // 2: K // 2 + K] if strided else wfull
    b = torch.randn(N) * 0.1
    y_ref = ORC.linear_fwd(x, w, b, relu)
    dy = torch.randn(M, N)
    dw_ref, db_ref = torch.zeros_like(w) + 0.5, torch.zeros(N) - 0.25  # accumulate on top of existing content
    dx_ref = ORC.linear_bwd(dy, x, w, b, y_ref, relu, dw_ref, db_ref, True)
    xg, bg, dyg = g(x, b, dy)
    wfg = wfull.to(DEV)
    wg = wfg[:, K // 2: K // 2 + K] if strided else wfg
    y = ops.linear_fwd(xg, wg, bg, relu)
    close(y, y_ref, what="y")
    dwf = torch.zeros_like(wfg) + 0.5
    dwg = dwf[:, K // 2: K // 2 + K] if strided else dwf
    dbg = torch.zeros(N, device=DEV) - 0.25
    dx = ops.linear_bwd(dyg, xg, wg, bg, y, relu, dwg, dbg, True)
    close(dx, dx_ref, what="dx")
    close(dwg, dw_ref, tol=5e-5, what="dw")
    close(dbg, db_ref, tol=5e-5, what="db")
    if strided:  # columns outside the slice untouched
        assert float((dwf[:, : K // 2] - 0.5).abs().max()) == 0.0


def test_linear_fused_epilogue(ops):
    """y = (x W^T + b) * keep_lin + res) * keep_out and the row-masked backward (dead attention rows, residual, invalid rows)."""
    torch.manual_seed(11)
    M, K, N = 203, 128, 128
    x, w, b, res, dy = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N) * 0.1, torch.randn(M, N), torch.randn(M, N)
    kl, ko = (torch.rand(M) < 0.7).to(torch.uint8), (torch.rand(M) < 0.8).to(torch.uint8)
    for keep_lin, r, keep_out in ((kl, res, None), (None, res, ko), (kl, res, ko), (kl, None, None)):
        y_ref = ORC.linear_fwd(x, w, b, False, keep_lin, r, keep_out)
        dw_ref, db_ref = torch.zeros_like(w), torch.zeros(N)
        dx_ref = ORC.linear_bwd(dy, x, w, b, y_ref, False, dw_ref, db_ref, True, keep_lin, keep_out)
        xg, wg, bg, dyg, klg, rg, kog = g(x, w, b, dy, keep_lin, r, keep_out)
        y = ops.linear_fwd(xg, wg, bg, False, klg, rg, kog)
        close(y, y_ref, what="y")
        dwg, dbg = torch.zeros_like(wg), torch.zeros(N, device=DEV)
        dx = ops.linear_bwd(dyg, xg, wg, bg, y, False, dwg, dbg, True, klg, kog)
        close(dx, dx_ref, what="dx")
        close(dwg, dw_ref, tol=5e-5, what="dw")
        close(dbg, db_ref, tol=5e-5, what="db")


def test_linear_tensor_core_path(ops):
    """>= 8192 rows and 128-wide operands: the tcgen05 kernels (bf16x3 split of both operands on the fly, fp32 accumulation in
    tensor memory; csrc/tb_train_tc.cu): forward with the fused epilogue for N = 128 / 384, dX / dW / db for N = 128 with ReLU,
    row masks and a ragged last tile."""
    torch.manual_seed(31)
    M, K = 20000 + 37, 128
    x, dy = torch.randn(M, K), torch.randn(M, 128)
    kl, ko = (torch.rand(M) < 0.7).to(torch.uint8), (torch.rand(M) < 0.8).to(torch.uint8)
    for N in (128, 384):
        w, b, res = torch.randn(N, K) / K ** 0.5, torch.randn(N) * 0.1, torch.randn(M, N)
        for relu, keep_lin, r, keep_out in ((True, None, None, None), (False, kl, res, ko)):
            y_ref = ORC.linear_fwd(x, w, b, relu, keep_lin, r, keep_out)
            y = ops.linear_fwd(*g(x, w, b), relu, *g(keep_lin, r, keep_out))
            close(y, y_ref, tol=1e-4, what=f"y N={N}")
    # backward: N = 128 (most Linears), N = 256 (K|V projection), N = 384 (GRU), K = 256 (mlp_out of add_goal / add_latent)
    for Kb, Nb, relu, keep_lin, keep_out in ((128, 128, True, None, None), (128, 128, False, kl, ko), (128, 256, False, None, None),
                                             (128, 384, False, None, ko), (256, 128, True, None, None)):
        x = torch.randn(M, Kb)
        dy = torch.randn(M, Nb)
        w, b = torch.randn(Nb, Kb) / Kb ** 0.5, torch.randn(Nb) * 0.1
        y_ref = ORC.linear_fwd(x, w, b, relu, keep_lin, None, keep_out)
        dw_ref, db_ref = torch.zeros_like(w) + 0.5, torch.zeros(Nb) - 0.25
        dx_ref = ORC.linear_bwd(dy, x, w, b, y_ref, relu, dw_ref, db_ref, True, keep_lin, keep_out)
        xg, wg, bg, dyg, klg, kog = g(x, w, b, dy, keep_lin, keep_out)
        y = ops.linear_fwd(xg, wg, bg, relu, klg, None, kog)
        close(y, y_ref, tol=1e-4, what=f"y K={Kb} N={Nb}")
        dwg, dbg = torch.zeros_like(wg) + 0.5, torch.zeros(Nb, device=DEV) - 0.25
        # the ReLU mask of the reference forward (outputs within rounding of 0 may have the other sign in the bf16x3 forward)
        dx = ops.linear_bwd(dyg, xg, wg, bg, y_ref.to(DEV), relu, dwg, dbg, True, klg, kog)
        close(dx, dx_ref, tol=1e-4, what="dx")
        close(dwg, dw_ref, tol=1e-4, what="dw")
        close(dbg, db_ref, tol=1e-4, what="db")


@pytest.mark.parametrize("M,relu", [(5, False), (777, True)])
def test_layernorm(ops, M, relu):
    torch.manual_seed(M)
    x, w, b, dy = torch.randn(M, 128) * 3 + 1, torch.rand(128) + 0.5, torch.randn(128) * 0.2, torch.randn(M, 128)
    y_ref, st_ref = ORC.layernorm_fwd(x, w, b, relu)
    dw_ref, db_ref = torch.ones(128), torch.ones(128)
    dx_ref = ORC.layernorm_bwd(dy, x, w, b, st_ref, y_ref, relu, dw_ref, db_ref)
    xg, wg, bg, dyg = g(x, w, b, dy)
    y, st = ops.layernorm_fwd(xg, wg, bg, relu)
    close(y, y_ref, what="y")
    close(st, st_ref, what="stats")
    dwg, dbg = torch.ones(128, device=DEV), torch.ones(128, device=DEV)
    dx = ops.layernorm_bwd(dyg, xg, wg, bg, st, y, relu, dwg, dbg)
    close(dx, dx_ref, tol=5e-5, what="dx")
    close(dwg, dw_ref, tol=5e-5, what="dw")
    close(dbg, db_ref, tol=5e-5, what="db")


@pytest.mark.parametrize("B,S,T,eye", [(3, 8, 64, False), (2, 20, 20, False), (5, 8, 8, True), (2, 70, 200, False), (1, 64, 1024, False),
                                       (4, 3, 40, False)])
def test_attention(ops, B, S, T, eye):
    torch.manual_seed(B * 1000 + S + T)
    q, kv, do = torch.randn(B, S, 128), torch.randn(B, T, 256), torch.randn(B, S, 128)
    kvalid = (torch.rand(B, T) < 0.7)
    kvalid[0] = False  # a batch entry without any valid key: every row dead
    if B > 1:
        kvalid[1] = False
        kvalid[1, 0] = True  # exactly one valid key (with eye: row 0 is dead)
    kvalid = kvalid.to(torch.uint8)
    o_ref, p_ref, alive_ref = ORC.attention_fwd(q, kv, kvalid, eye)
    dq_ref, dkv_ref = ORC.attention_bwd(do, q, kv, kvalid, eye, p_ref)
    qg, kvg, dog, kvg_valid = g(q, kv, do, kvalid)
    o, p, alive = ops.attention_fwd(qg, kvg, kvg_valid, eye)
    close(o, o_ref, what="o")
    close(p[0], p_ref, what="p")
    assert torch.equal(alive.cpu(), alive_ref) and not bool(alive_ref[0].any())  # batch element 0 has no valid key
    dq, dkv = ops.attention_bwd(dog, qg, kvg, kvg_valid, eye, p)
    close(dq, dq_ref, tol=5e-5, what="dq", atol=1e-5)  # rows with a single admissible key: dq == 0 up to rounding
    close(dkv, dkv_ref, tol=5e-5, what="dkv")


def test_glue_ops(ops):
    torch.manual_seed(1)
    M = 333
    a, b, dy = torch.randn(M, 128), torch.randn(M, 128), torch.randn(M, 128)
    keep = (torch.rand(M) < 0.6).to(torch.uint8)
    ag, bg, dyg, kg = g(a, b, dy, keep)
    close(ops.add_mask_fwd(ag, bg, kg), ORC.add_mask_fwd(a, b, keep))
    close(ops.add_mask_fwd(ag, None, kg), ORC.add_mask_fwd(a, None, keep))
    close(ops.add_mask_fwd(ag, bg, None), ORC.add_mask_fwd(a, b, None))
    close(ops.add_mask_bwd(dyg, kg), ORC.add_mask_bwd(dy, keep))
    keep2 = (torch.rand(M) < 0.5).to(torch.uint8)
    close(ops.add_mask_fwd(ag, bg, kg, keep2.to(DEV)), ORC.add_mask_fwd(a, b, keep, keep2))
    close(ops.add_mask_bwd(dyg, kg, keep2.to(DEV)), ORC.add_mask_bwd(dy, keep, keep2))
    close(ops.select_rows_fwd(kg, ag, bg), ORC.select_rows_fwd(keep, a, b))
    for x, y in zip(ops.select_rows_bwd(dyg, kg), ORC.select_rows_bwd(dy, keep)):
        close(x, y)
    c = torch.randn(M, 96)
    close(ops.cat2_fwd(ag[:, :32].contiguous(), c.to(DEV)), ORC.cat2_fwd(a[:, :32], c))
    for x, y in zip(ops.cat2_bwd(dyg, 32), ORC.cat2_bwd(dy, 32)):
        close(x, y)
    dst = torch.randn(M, 128)
    dstg = dst.to(DEV)
    ops.add_(dstg, ag)
    close(dstg, dst + a)
    big = torch.zeros(M, 256, device=DEV)
    ops.add_(big[:, 64:192], ag)  # strided destination (gradient view of a column slice)
    close(big[:, 64:192], a)
    ops.scale_(dstg, 0.25)
    close(dstg, (dst + a) * 0.25)
    # GRU gates
    gi, gh, h, dh = torch.randn(M, 384), torch.randn(M, 384), torch.randn(M, 128), torch.randn(M, 128)
    gig, ghg, hg, dhg = g(gi, gh, h, dh)
    close(ops.gru_gates_fwd(gig, ghg, hg), ORC.gru_gates_fwd(gi, gh, h))
    for x, y in zip(ops.gru_gates_bwd(dhg, gig, ghg, hg), ORC.gru_gates_bwd(dh, gi, gh, h)):
        close(x, y)
    # gather / scatter-add
    idx = torch.randint(0, 50, (M,))
    src = torch.randn(50, 128)
    close(ops.gather_rows_fwd(src.to(DEV), idx.to(DEV)), ORC.gather_rows_fwd(src, idx))
    close(ops.gather_rows_bwd(dyg, idx.to(DEV), 50), ORC.gather_rows_bwd(dy, idx, 50), tol=5e-5)


@pytest.mark.parametrize("O,R,I,fill", [(40, 20, 1, float("-inf")), (1, 7, 33, -1e3)])
def test_masked_max(ops, O, R, I, fill):
    torch.manual_seed(O)
    x = torch.randn(O, R, I, 128)
    valid = (torch.rand(O, R, I) < 0.5)
    valid[0, :, 0] = False
    valid = valid.to(torch.uint8)
    y_ref, idx_ref = ORC.masked_max_fwd(x, valid, fill)
    y, idx = ops.masked_max_fwd(x.to(DEV), valid.to(DEV), fill)
    close(y, y_ref)
    assert torch.equal(idx.cpu(), idx_ref)
    dy = torch.randn(O, I, 128)
    close(ops.masked_max_bwd(dy.to(DEV), idx, R), ORC.masked_max_bwd(dy, idx_ref, R))


def test_dest_ops(ops):
    torch.manual_seed(3)
    S, P, A = 2, 50, 6
    u, v, dy = torch.randn(S, P, 128), torch.randn(S, A, 128), torch.randn(S, A, P, 128)
    close(ops.pair_add_fwd(u.to(DEV), v.to(DEV)), ORC.pair_add_fwd(u, v))
    for x, y in zip(ops.pair_add_bwd(dy.to(DEV)), ORC.pair_add_bwd(dy)):
        close(x, y, tol=5e-5)
    logits = torch.randn(S, A, P) * 2
    ok = (torch.rand(S, A, P) < 0.4)
    ok[0, 0] = False  # a row that ends up all -inf -> reset to 0
    gt = torch.randint(0, P, (S, A))
    ok[torch.arange(S)[:, None], torch.arange(A)[None], gt] = True
    ok[0, 0] = False
    row_valid = torch.ones(S, A, dtype=torch.bool)
    row_valid[1, 2] = False
    loss_rows = row_valid.clone()
    loss_rows[1, 3] = False
    scale = torch.tensor([0.37])
    u8 = torch.uint8
    nll_ref, dl_ref = ORC.dest_nll(logits, ok.to(u8), row_valid.to(u8), gt, loss_rows.to(u8), scale)
    nll, dl = ops.dest_nll(*g(logits, ok.to(u8), row_valid.to(u8), gt, loss_rows.to(u8), scale))
    close(nll, nll_ref)
    close(dl, dl_ref)


def test_latent_ops(ops):
    torch.manual_seed(4)
    M, E = 50, 16
    mean, ls, eps, dz = torch.randn(M, E), torch.randn(E) * 0.1 - 1, torch.randn(M, E), torch.randn(M, E)
    close(ops.rsample_fwd(*g(mean, ls, eps)), ORC.rsample_fwd(mean, ls, eps))
    dls_ref, dls = torch.ones(E), torch.ones(E, device=DEV)
    ORC.rsample_bwd(dz, eps, ls, dls_ref)
    ops.rsample_bwd(dz.to(DEV), eps.to(DEV), ls.to(DEV), dls)
    close(dls, dls_ref, tol=5e-5)
    mq, mp, lq, lp = torch.randn(M, E) * 0.3, torch.randn(M, E) * 0.3, torch.randn(E) * 0.1 - 1, torch.randn(E) * 0.1 - 1
    mq[:5] = mp[:5]  # rows below the free-nats floor: no gradient
    lq2 = lp.clone()
    valid = (torch.rand(M) < 0.8).to(torch.uint8)
    scale = torch.tensor([0.05])
    for lq_ in (lq, lq2):
        a_ref, b_ref = torch.zeros(E), torch.zeros(E)
        ref = ORC.kl_fwd_bwd(mq, lq_, mp, lp, valid, 0.01, scale, a_ref, b_ref)
        a, b = torch.zeros(E, device=DEV), torch.zeros(E, device=DEV)
        out = ops.kl_fwd_bwd(*g(mq, lq_, mp, lp, valid), 0.01, scale.to(DEV), a, b)
        for x, y in zip(out, ref):
            close(x, y, tol=5e-5)
        close(a, a_ref, tol=5e-5)
        close(b, b_ref, tol=5e-5)
    x = torch.randn(40, 9)
    mask = (torch.rand(40, 9) < 0.5).to(torch.uint8)
    close(ops.masked_sum(x.to(DEV), mask.to(DEV)), ORC.masked_sum(x, mask))
    close(ops.mask_scale(mask[:, 0].contiguous().to(DEV), scale.to(DEV)), ORC.mask_scale(mask[:, 0], scale))


def test_pose_pe_and_simulation_ops(ops):
    from trafficbots_b200 import weights
    torch.manual_seed(5)
    M = 200
    xy, yaw, d = (torch.rand(M, 2) - 0.5) * 200, (torch.rand(M) - 0.5) * 6.28, torch.randn(M, 2)
    fx, fy = weights.pe_freqs_xy(), weights.pe_freqs_yaw()
    close(ops.pose_pe(*g(xy, yaw, fx, fy)), ORC.pose_pe(xy, yaw, fx, fy), tol=1e-5)
    close(ops.dir_to_yaw(d.to(DEV)), ORC.dir_to_yaw(d), tol=1e-6)
    state = torch.cat([xy, yaw[:, None], torch.rand(M, 1) * 10], -1)
    mean, dp = torch.randn(M, 2), torch.randn(M, 4)
    a_type = torch.nn.functional.one_hot(torch.randint(0, 3, (M,)), 3).to(torch.uint8)
    a_type[:3] = 0  # agents without a type
    valid = (torch.rand(M) < 0.8).to(torch.uint8)
    pred_ref = ORC.dynamics_fwd(state, mean, a_type, valid)
    pred = ops.dynamics_fwd(*g(state, mean, a_type, valid))
    close(pred, pred_ref, tol=1e-6)
    for x, y in zip(ops.dynamics_bwd(*g(dp, state, mean, a_type, valid)), ORC.dynamics_bwd(dp, state, mean, a_type, valid)):
        close(x, y)
    gt = state + torch.randn(M, 4) * torch.tensor([2.0, 2.0, 0.5, 1.0])
    rv = (torch.rand(M) < 0.7).to(torch.uint8)
    close(ops.reward_fwd(*g(pred_ref, gt, rv)), ORC.reward_fwd(pred_ref, gt, rv))
    dr = torch.randn(M)
    close(ops.reward_bwd(*g(dr, pred_ref, gt, rv)), ORC.reward_bwd(dr, pred_ref, gt, rv))
    # bookkeeping
    B, A = 4, 50
    st = state.view(B, A, 4).contiguous()
    boundary = torch.tensor([[-80.0, 80.0, -80.0, 80.0]]).expand(B, 4).contiguous()
    dest_pos = st[..., None, :2] + torch.randn(B, A, 20, 2) * 40
    dest_dir = torch.randn(B, A, 20, 2)
    dest_dir[0, 0, 0] = 0.0  # zero-length direction (NaN after normalisation, masked: traffic_rule_checker.py:93,400)
    u8 = torch.uint8
    dest_valid = (torch.rand(B, A, 20) < 0.7).to(u8)
    lane, edge = (torch.rand(B, A) < 0.6).to(u8), (torch.rand(B, A) < 0.3).to(u8)
    flags = [(torch.rand(B, A) < p).to(u8) for p in (0.1, 0.1, 0.8)]
    vb, gtv = valid.view(B, A).contiguous(), (torch.rand(B, A) < 0.5).to(u8)
    for gt_valid_t in (gtv, None):
        ref = ORC.sim_flags(st, vb, gt_valid_t, boundary, dest_pos, dest_dir, dest_valid, lane, edge, *flags)
        out = ops.sim_flags(*g(st, vb, gt_valid_t, boundary, dest_pos, dest_dir, dest_valid, lane, edge, *flags))
        for x, y in zip(out, ref):
            assert torch.equal(x.cpu(), y)


@pytest.mark.parametrize("case", TRAIN_CASES)
def test_training_step_matches_reference(case):
    from trafficbots_b200.train import trainer
    c = load_train_case(case)
    ts = trainer.TrainState(c["sd"], device=DEV)
    batch = {k: v.to(DEV) for k, v in c["batch"].items()}
    n0 = ts.ops.L.tb_launch_count()
    out = ts.forward_backward(batch, c["eps"], c["use_prior"])
    torch.cuda.synchronize()
    assert ts.ops.L.tb_launch_count() - n0 > 10000  # the CUDA primitives did the work
    for k, ref in c["terms"].items():
        assert abs(float(out[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(out[k]), ref)
    # 2e-3 of each gradient's scale: fp32 kernels with other summation orders than torch's, amplified by 80 closed-loop steps
    # (the CPU restatement of the same composition stays within 1e-3: tests/test_train_cpu.py; measured here 0.5-1.2e-3)
    worst = compare_grads({k: v.cpu() for k, v in ts.grads().items()}, c["grads"], rel=2e-3)
    print(f"{case}: loss {float(out['loss']):.6f} (reference {c['terms']['loss']:.6f}), worst gradient deviation {worst:.2e}")


@pytest.mark.parametrize("case", TRAIN_CASES)
def test_training_step_in_concurrent_sub_batches_matches_reference(case):
    """n_split = 2: the scenes of a step run as two concurrent chains on side streams (1 + 1 and 1 + 2 scenes here); the loss
    normalisers stay batch-wide, parameter gradients of both chains accumulate atomically: same loss / gradients."""
    from trafficbots_b200.train import trainer
    c = load_train_case(case)
    ts = trainer.TrainState(c["sd"], device=DEV, n_split=2)
    batch = {k: v.to(DEV) for k, v in c["batch"].items()}
    out = ts.forward_backward(batch, c["eps"], c["use_prior"])
    torch.cuda.synchronize()
    for k, ref in c["terms"].items():
        assert abs(float(out[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(out[k]), ref)
    compare_grads({k: v.cpu() for k, v in ts.grads().items()}, c["grads"], rel=2e-3)


def test_optimizer_steps_match_cpu_restatement():
    from trafficbots_b200.train import trainer
    c = load_train_case(TRAIN_CASES[0])
    ts = trainer.TrainState(c["sd"], device=DEV, lr=1e-3)
    ref = trainer.TrainState(c["sd"], device="cpu", ops=ORC, lr=1e-3)
    batch = {k: v.to(DEV) for k, v in c["batch"].items()}
    losses = []
    for i in range(3):  # clipped Adam steps (lr 1e-3): the loss trajectory must be the CPU restatement's
        out = ts.training_step(batch, c["eps"], c["use_prior"])
        out_ref = ref.training_step(c["batch"], c["eps"], c["use_prior"])
        losses.append(float(out["loss"]))
        assert abs(losses[-1] - float(out_ref["loss"])) <= 2e-4 * abs(float(out_ref["loss"])), (i, losses[-1], float(out_ref["loss"]))
    # Adam normalises every element's update to ~lr: entries whose gradient is zero up to rounding may move in either direction,
    # everything else must follow the CPU restatement
    diff = (ts.flat_p.cpu() - ref.flat_p).abs()
    assert float((diff > 2e-4).float().mean()) < 5e-3, float((diff > 2e-4).float().mean())
    assert float(diff.max()) <= 3.5 * 1e-3

def test_public_training_step_and_inference_sees_updated_weights():
    """`WaymoMotion.training_step` (the reference's entry point, waymo_motion.py:356): gradients land in `p.grad` like after
    Lightning's backward; with automatic optimization the Adam step runs too and the inference engine re-packs the weights."""
    from trafficbots_b200 import config
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    c = load_train_case(TRAIN_CASES[0])
    m = WaymoMotion(**config.default_config(n_joint_future=1))
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    assert m.train_dropout_p == 0.1  # the reference's default; the parity fixtures were produced with dropout 0
    m.train_dropout_p = 0.0
    batch = {k: v.to(DEV) for k, v in c["batch"].items()}
    m.automatic_optimization = False
    import ref_train
    torch.manual_seed(5)  # the fixture's noise seed: torch.rand(1) for prior/posterior, then the rsample noise
    loss = m.training_step(batch, 0)
    assert abs(float(loss) - c["terms"]["loss"]) <= 1e-4 * abs(c["terms"]["loss"])
    grads = {k: p.grad.cpu() for k, p in m.named_parameters(remove_duplicate=False) if p.grad is not None and k in c["grads"]}
    assert len(grads) == len(c["grads"]) == 382  # every parameter but the three action_head.log_std (deterministic actions)
    compare_grads(grads, c["grads"], rel=2e-3)
    assert set(m.state_dict().keys()) == set(c["sd"].keys())
    # one optimizer step: parameters move, the engine used by the inference path picks them up
    feat0 = m.model.encode_input_features(batch)["map_feature"].clone()
    m.automatic_optimization = True
    torch.manual_seed(5)
    m.training_step(batch, 1)
    w = dict(m.named_parameters(remove_duplicate=False))["model.map_encoder.transformer_self_attn.layers.0.linear2.weight"]
    assert float((w.detach().cpu() - c["sd"]["model.map_encoder.transformer_self_attn.layers.0.linear2.weight"]).abs().max()) > 1e-5
    feat1 = m.model.encode_input_features(batch)["map_feature"]
    assert float((feat1 - feat0).abs().max()) > 1e-6
    opt, sch = m.configure_optimizers()
    assert len(opt) == 1 and len(opt[0].param_groups) == 2 and sch[0]["interval"] == "epoch"


# ---------------------------------------------------------------------------------------------------------------------------
# dropout: counter-based hash masks (seed on the device, site id, element index), regenerated in the backward kernels
# ---------------------------------------------------------------------------------------------------------------------------
def _drop(site, p=0.1, seed=12345):
    return (torch.tensor([seed], dtype=torch.int32), site, p)


def _dropg(d):
    return (d[0].to(DEV), d[1], d[2])


def test_dropout_mask_statistics_and_elementwise_op(ops):
    x = torch.ones(1000, 128)
    for site, p in ((1, 0.1), (2, 0.1), (77, 0.5)):
        d = _drop(site, p)
        y = ops.dropout(x.to(DEV), _dropg(d)).cpu()
        close(y, ORC.dropout(x, d), tol=1e-7)
        keep = float((y > 0).float().mean())
        assert abs(keep - (1 - p)) < 0.01, (site, p, keep)
        assert abs(float(y.mean()) - 1.0) < 0.02
    a = ops.dropout(x.to(DEV), _dropg(_drop(1))).cpu()
    b = ops.dropout(x.to(DEV), _dropg(_drop(2))).cpu()
    c = ops.dropout(x.to(DEV), _dropg(_drop(1, seed=999))).cpu()
    assert 0.7 < float(((a > 0) == (b > 0)).float().mean()) < 0.9  # different sites / seeds: independent masks
    assert 0.7 < float(((a > 0) == (c > 0)).float().mean()) < 0.9


def test_dropout_in_linear_layernorm_attention(ops):
    torch.manual_seed(21)
    M, K, N = 300, 128, 128
    x, w, b, res, dy = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N) * 0.1, torch.randn(M, N), torch.randn(M, N)
    kl, ko = (torch.rand(M) < 0.7).to(torch.uint8), (torch.rand(M) < 0.8).to(torch.uint8)
    for relu, keep_lin, r, keep_out, site in ((True, None, None, None, 3), (False, kl, res, None, 4), (False, None, res, ko, 5)):
        d = _drop(site)
        y_ref = ORC.linear_fwd(x, w, b, relu, keep_lin, r, keep_out, d)
        dw_ref, db_ref = torch.zeros_like(w), torch.zeros(N)
        dx_ref = ORC.linear_bwd(dy, x, w, b, y_ref, relu, dw_ref, db_ref, True, keep_lin, keep_out, d)
        xg, wg, bg, dyg, klg, rg, kog = g(x, w, b, dy, keep_lin, r, keep_out)
        y = ops.linear_fwd(xg, wg, bg, relu, klg, rg, kog, _dropg(d))
        close(y, y_ref, what="y")
        dwg, dbg = torch.zeros_like(wg), torch.zeros(N, device=DEV)
        dx = ops.linear_bwd(dyg, xg, wg, bg, y, relu, dwg, dbg, True, klg, kog, _dropg(d))
        close(dx, dx_ref, what="dx")
        close(dwg, dw_ref, tol=5e-5, what="dw")
        close(dbg, db_ref, tol=5e-5, what="db")
    # big-M path (64-row tiles) with a weight-gradient row split
    M2 = 9000
    x2, dy2 = torch.randn(M2, K), torch.randn(M2, N)
    d = _drop(6)
    y_ref = ORC.linear_fwd(x2, w, b, True, drop=d)
    dw_ref, db_ref = torch.zeros_like(w), torch.zeros(N)
    dx_ref = ORC.linear_bwd(dy2, x2, w, b, y_ref, True, dw_ref, db_ref, True, drop=d)
    y = ops.linear_fwd(x2.to(DEV), w.to(DEV), b.to(DEV), True, drop=_dropg(d))
    close(y, y_ref, tol=1e-4)  # tensor-core path (bf16x3)
    dwg, dbg = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    # ReLU mask of the reference forward: outputs within rounding of 0 may have the other sign in the bf16x3 forward
    close(ops.linear_bwd(dy2.to(DEV), x2.to(DEV), w.to(DEV), b.to(DEV), y_ref.to(DEV), True, dwg, dbg, True, drop=_dropg(d)), dx_ref,
          tol=1e-4)
    close(dwg, dw_ref, tol=1e-4)
    # LayerNorm + ReLU + dropout (add_goal.mlp_in)
    lw, lb = torch.rand(128) + 0.5, torch.randn(128) * 0.2
    d = _drop(7)
    y_ref, st_ref = ORC.layernorm_fwd(x, lw, lb, True, d)
    dw_ref, db_ref = torch.zeros(128), torch.zeros(128)
    dx_ref = ORC.layernorm_bwd(dy, x, lw, lb, st_ref, y_ref, True, dw_ref, db_ref, d)
    y, st = ops.layernorm_fwd(x.to(DEV), lw.to(DEV), lb.to(DEV), True, _dropg(d))
    close(y, y_ref)
    dwg, dbg = torch.zeros(128, device=DEV), torch.zeros(128, device=DEV)
    close(ops.layernorm_bwd(dy.to(DEV), x.to(DEV), lw.to(DEV), lb.to(DEV), st, y, True, dwg, dbg, _dropg(d)), dx_ref, tol=5e-5)
    close(dwg, dw_ref, tol=5e-5)
    # attention probabilities: general kernel (T = 200, two key chunks) and the warp-per-head kernel (T = 20)
    for B, S, T, eye, site in ((2, 70, 200, False, 8), (3, 20, 20, False, 9), (4, 8, 8, True, 10)):
        q, kv, do = torch.randn(B, S, 128), torch.randn(B, T, 256), torch.randn(B, S, 128)
        kvalid = (torch.rand(B, T) < 0.8).to(torch.uint8)
        d = _drop(site)
        o_ref, p_ref, _alive_ref = ORC.attention_fwd(q, kv, kvalid, eye, d)
        dq_ref, dkv_ref = ORC.attention_bwd(do, q, kv, kvalid, eye, p_ref, d)
        o, p, _alive = ops.attention_fwd(q.to(DEV), kv.to(DEV), kvalid.to(DEV), eye, _dropg(d))
        close(o, o_ref, what="o")
        close(p[0], p_ref, what="p")
        dq, dkv = ops.attention_bwd(do.to(DEV), q.to(DEV), kv.to(DEV), kvalid.to(DEV), eye, p, _dropg(d))
        close(dq, dq_ref, tol=5e-5, what="dq", atol=1e-5)
        close(dkv, dkv_ref, tol=5e-5, what="dkv")


def test_training_step_with_dropout_matches_cpu_restatement():
    """the whole step with dropout 0.1: the hash masks are a pure function of (seed, site, element), so the torch restatement on
    the CPU draws the same masks; loss and gradients must agree (closed-loop amplification as in the dropout-free test)."""
    from trafficbots_b200.train import trainer
    c = load_train_case(TRAIN_CASES[0])
    ts = trainer.TrainState(c["sd"], device=DEV, dropout_p=0.1)
    ref = trainer.TrainState(c["sd"], device="cpu", ops=ORC, dropout_p=0.1)
    for t in (ts, ref):
        t.new_dropout_seed(4242)
    batch = {k: v.to(DEV) for k, v in c["batch"].items()}
    out = ts.forward_backward(batch, c["eps"], c["use_prior"], new_seed=False)
    out_ref = ref.forward_backward(c["batch"], c["eps"], c["use_prior"], new_seed=False)
    assert abs(float(out["loss"]) - float(out_ref["loss"])) <= 2e-4 * abs(float(out_ref["loss"]))
    assert abs(float(out["loss"]) - c["terms"]["loss"]) > 1e-3  # dropout really changed the step
    total = float(ref.flat_g.norm())
    worst = 0.0
    for k, gref in ref.grads().items():
        n = gref.numel()
        err = float((ts.grads()[k].cpu() - gref).abs().max())
        scale = float(gref.norm()) / n ** 0.5 * 8.0 + 1e-6 * total
        worst = max(worst, err / scale)
    assert worst <= 5e-3, worst
    # a second step draws new masks
    out2 = ts.forward_backward(batch, c["eps"], c["use_prior"])
    assert abs(float(out2["loss"]) - float(out["loss"])) > 1e-6


@pytest.mark.reference
def test_full_size_step_vs_reference_on_the_gpu():
    """BASELINE.json configs[3] scene shape (64 agents, 1024 polylines, 90 steps; 2 scenes): the UNMODIFIED reference's
    `training_step` + backward executed on the same GPU (fp32, TF32 off, dropout 0) against the CUDA step: loss terms and the
    gradients of all 382 parameters.  Needs the reference sources on the box (tools/install_reference.sh -> baseline/_ref)."""
    import torch._dynamo  # noqa: F401  (before ref_loader's module stubs)
    import ref_loader
    import ref_train
    from trafficbots_b200 import synthetic, weights
    from trafficbots_b200.train import trainer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    S, A, P = 2, 64, 1024
    sd = weights.init_state_dict(2023)
    batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=4321)
    model = ref_loader.build_reference(n_agent=A, n_pl=P, n_joint_future=1)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV)
    # the reference draws its latent noise on the device: replay the host-side protocol with a generator-independent patch
    use_prior, eps = ref_train.draw_training_noise(11, S, A)
    import models.modules.distributions as dist_mod
    orig = dist_mod.MyDist.sample

    def sample(self, deterministic):
        if deterministic is False:
            return self.distribution.mean + eps.to(DEV) * self.distribution.stddev
        return orig(self, deterministic)
    dist_mod.MyDist.sample = sample
    try:
        terms, grads, _ = ref_train.run_reference_training(model, {k: v.to(DEV) for k, v in batch.items()}, seed=11,
                                                           p_prior=1.0 if use_prior else 0.0)
    finally:
        dist_mod.MyDist.sample = orig
    ts = trainer.TrainState(sd, device=DEV)
    out = ts.forward_backward({k: v.to(DEV) for k, v in batch.items()}, eps, use_prior)
    for k in ("loss", "vae_kl", "diffbar_reward", "goal_loss"):
        assert abs(float(out[k]) - float(terms[k])) <= 2e-4 * max(1.0, abs(float(terms[k]))), (k, float(out[k]), float(terms[k]))
    total = sum(float((g_.double() ** 2).sum()) for g_ in grads.values() if g_ is not None) ** 0.5
    worst = 0.0
    for k, g_ in ts.grads().items():
        if grads[k] is None:
            assert float(g_.abs().max()) == 0.0, k
            continue
        err = float((g_ - grads[k]).abs().max())
        scale = float(grads[k].norm()) / g_.numel() ** 0.5 * 8.0 + 1e-6 * total
        worst = max(worst, err / scale)
    print(f"full-size step vs reference on the GPU: loss {float(out['loss']):.6f} vs {float(terms['loss']):.6f}, worst gradient deviation {worst:.2e}")
    assert worst <= 5e-3, worst

"""GPU: Waymo post-processing and WOMD packing (`tb_post_process`, `tb_womd_pack`) through the mirrors of the
reference's classes, against golden vectors produced by the reference's own `WaymoPostProcessing.forward` and
`WOMDMetrics.update` (`oracle/make_golden.py`, POST_CASES) and against the oracle on a larger shape."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ("default", "mtr_nms", "mpa_nms", "mtr_mpa_fde", "a64_k6")
STATES = ("prediction_trajectory", "prediction_score", "ground_truth_trajectory", "ground_truth_is_valid",
          "prediction_ground_truth_indices_mask", "object_type")


def _load(name):
    import os
    from golden_util import GOLDEN_DIR, checksum
    from trafficbots_b200 import synthetic
    z = np.load(os.path.join(GOLDEN_DIR, "post_cases.npz"))
    S, A, n, seed, tseed, ade = [int(x) for x in z[f"{name}__meta"]]
    batch = synthetic.make_batch(S, n_agent=A, n_pl=16, seed=seed)
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, n, seed=tseed)
    want = float(z[f"{name}__checksum"])
    assert abs(checksum({"s": scores, "t": trajs, "v": valid.float()}) - want) <= 1e-6 * abs(want)
    gold = {k[len(name) + 2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "__")}
    return gold, batch, valid, scores, trajs, dict(mtr=list(z[f"{name}__mtr"]), mpa=list(z[f"{name}__mpa"]), ade=bool(ade))


@pytest.mark.parametrize("name", CASES)
def test_post_processing_and_womd_pack_match_reference(name):
    from trafficbots_b200.data_modules.waymo_post_processing import WaymoPostProcessing
    from trafficbots_b200.models.metrics.womd import WOMDMetrics
    gold, batch, valid, scores, trajs, cfg = _load(name)
    pp = WaymoPostProcessing(k_pred=6, score_temperature=1e2, mpa_nms_thresh=cfg["mpa"], mtr_nms_thresh=cfg["mtr"], aggr_thresh=[],
                             n_iter_em=3, use_ade=cfg["ade"])
    cb = {k: v.cuda() for k, v in batch.items()}
    d = pp(valid=valid.cuda(), scores=scores.cuda(), trajs=trajs.cuda(), agent_type=cb["agent/type"])
    for k in ("waymo_trajs", "waymo_yaw_bbox", "waymo_spd"):
        assert torch.equal(d[k].cpu(), gold[k]), k  # pure selection / re-layout: bit exact
    assert float((d["waymo_scores"].cpu() - gold["waymo_scores"]).abs().max()) <= 1e-6
    assert d["waymo_valid"].shape == (valid.shape[0], 80, valid.shape[1])
    m = WOMDMetrics("val", step_gt=90, step_current=10, interactive_challenge=False)
    m.update(cb, d["waymo_trajs"], d["waymo_scores"])
    out = m.compute()
    for k in STATES:
        got = out[k][0].cpu()
        assert got.shape == gold[k].shape, k
        if k == "prediction_score":
            assert float((got - gold[k]).abs().max()) <= 1e-6
        else:
            assert torch.equal(got, gold[k]), k
    assert int(m.overflow.item()) == 0


def test_post_processing_reads_the_rollout_buffer_in_place():
    """`trajs = rollout_buffer.preds[:, :, :, step_future_start:]` (waymo_motion.py:714) is a strided view of the kernels'
    [S*K, A, T, 4] output: same result as a dense copy, no copy made."""
    from trafficbots_b200.data_modules.waymo_post_processing import WaymoPostProcessing
    from trafficbots_b200 import synthetic
    S, A, K, T, t0 = 3, 10, 6, 90, 10
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, K, seed=21, n_step=T)
    raw = trajs.transpose(1, 2).reshape(S * K, A, T, 4).contiguous().cuda()  # the rollout kernels' layout
    view = raw.view(S, K, A, T, 4).transpose(1, 2)[:, :, :, t0:]  # flatten_repeat + future slice
    assert not view.is_contiguous()
    at = torch.nn.functional.one_hot(torch.arange(S * A) % 3, 3).bool().view(S, A, 3).cuda()
    pp = WaymoPostProcessing(k_pred=6, mpa_nms_thresh=[2.5, 1.0, 1.5])
    d1 = pp(valid.cuda(), scores.cuda(), view, at)
    d2 = pp(valid.cuda(), scores.cuda(), view.contiguous(), at)
    for k in ("waymo_trajs", "waymo_yaw_bbox", "waymo_spd", "waymo_scores"):
        assert torch.equal(d1[k], d2[k]), k
    assert torch.equal(d1["waymo_trajs"].cpu(), trajs[:, :, :, t0:, :2].movedim(3, 1))


def test_post_vs_oracle_larger_shape():
    import post_oracle as po
    from trafficbots_b200 import synthetic
    from trafficbots_b200.data_modules.waymo_post_processing import WaymoPostProcessing
    from trafficbots_b200.models.metrics.womd import WOMDMetrics
    S, A, n = 4, 64, 24
    batch = synthetic.make_batch(S, n_agent=A, n_pl=16, seed=5)
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, n, seed=99)
    ref = po.post_process(valid, scores, trajs, batch["agent/type"], 6, 1e2, [2.0, 1.0, 1.5], [2.5, 1.0, 1.5], True)
    pp = WaymoPostProcessing(k_pred=6, mpa_nms_thresh=[2.0, 1.0, 1.5], mtr_nms_thresh=[2.5, 1.0, 1.5])
    cb = {k: v.cuda() for k, v in batch.items()}
    d = pp(valid.cuda(), scores.cuda(), trajs.cuda(), cb["agent/type"])
    same = (d["mode_idx"].cpu().long() == ref["mode_idx"]).all(-1)  # an ADE within an ulp of a threshold may pick another mode
    assert same.float().mean() >= 0.99
    assert torch.equal(d["waymo_trajs"].cpu().movedim(1, 3)[same], ref["waymo_trajs"].movedim(1, 3)[same])
    assert float((d["waymo_scores"].cpu() - ref["waymo_scores"])[same].abs().max()) <= 1e-6
    want = po.womd_pack(batch, d["waymo_trajs"].cpu(), d["waymo_scores"].cpu())
    m = WOMDMetrics()
    m.update(cb, d["waymo_trajs"], d["waymo_scores"])
    got = m.compute()
    for k in STATES:
        assert torch.equal(got[k][0].cpu(), want[k]), k
    m.update(cb, d["waymo_trajs"], None)  # uniform scores (womd.py:105-106)
    assert torch.equal(m.compute()["prediction_score"][1].cpu(), po.womd_pack(batch, d["waymo_trajs"].cpu(), None)["prediction_score"])

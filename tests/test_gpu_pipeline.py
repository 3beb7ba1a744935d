"""GPU: `pipeline.ScenePipeline` -- several batches in flight on separate streams give bit-identical results to the same
batches run one after the other, and the 1-CTA-per-scene-mode decode kernel agrees with the cluster-split one."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(K=1):
    from trafficbots_b200 import config, weights
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    sd = weights.init_state_dict(2023)
    m = WaymoMotion(**config.default_config(n_joint_future=K))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _batches(n, S, A, P):
    import bench
    from trafficbots_b200 import host, synthetic
    out = []
    for i in range(n):
        b = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=500 + i)
        out.append(host.pin_batch({k: b[k] for k in bench.USED_KEYS}))
    return out


def _run(pipe, batches):
    res, tickets = [], []
    for i, hb in enumerate(batches):
        if i >= pipe.depth:
            r = pipe.result(tickets[i - pipe.depth])
            res.append({k: v.clone() for k, v in r.items()})
        tickets.append(pipe.submit(hb))
    for t in tickets[len(res):]:
        r = pipe.result(t)
        res.append({k: v.clone() for k, v in r.items()})
    return res


@pytest.mark.parametrize("shape", [(3, 16, 128), (2, 64, 256), (2, 96, 192)])  # the last: two-CTA agent-halves decode mode
def test_in_flight_equals_sequential_bit_for_bit(shape):
    from trafficbots_b200.pipeline import ScenePipeline
    S, A, P = shape
    m = _module()
    batches = _batches(7, S, A, P)
    seq = _run(ScenePipeline(m, depth=1, rollout_cluster=1), batches)
    par = _run(ScenePipeline(m, depth=4), batches)  # depth > 1 defaults to 1-CTA clusters
    assert len(seq) == len(par) == 7
    for a, b in zip(seq, par):
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k
    # different batches do give different results (the comparison above is not vacuous)
    assert not torch.equal(seq[0]["preds"], seq[1]["preds"])


def test_cluster_sizes_agree_through_the_pipeline():
    """cluster-split (4 CTAs per scene-mode, library default for small batches) vs 1 CTA per scene-mode: the same
    arithmetic up to the merge order of the online-softmax partials."""
    from trafficbots_b200.pipeline import ScenePipeline
    m = _module()
    batches = _batches(2, 2, 64, 256)
    one = _run(ScenePipeline(m, depth=1, rollout_cluster=1), batches)
    four = _run(ScenePipeline(m, depth=1, rollout_cluster=4), batches)
    for a, b in zip(one, four):
        assert torch.equal(a["valid"], b["valid"])
        assert float((a["preds"] - b["preds"]).abs().max()) <= 2e-3  # closed-loop tolerance (tests/test_gpu_parity.py)


def test_slot_reuse_is_detected():
    from trafficbots_b200.pipeline import ScenePipeline
    m = _module()
    batches = _batches(3, 1, 8, 64)
    pipe = ScenePipeline(m, depth=2)
    t0 = pipe.submit(batches[0])
    pipe.submit(batches[1])
    pipe.submit(batches[2])  # reuses slot 0 before ticket 0 was collected
    with pytest.raises(RuntimeError):
        pipe.result(t0)
    pipe.drain()


def test_forward_rejects_overrides():
    from trafficbots_b200 import config
    m = _module()
    with pytest.raises(Exception):
        m.forward()  # no rollout open
    m._step_ctx = {"dummy": True}
    try:
        with pytest.raises(config.UnsupportedConfig):
            m.forward(action_override=torch.zeros(1))
        with pytest.raises(config.UnsupportedConfig):
            m.forward(mask_state_override=torch.zeros(1, dtype=torch.bool))
        with pytest.raises(config.UnsupportedConfig):
            m.forward(require_vis_dict=True)
    finally:
        m._step_ctx = None

"""world_size-2 gloo test of the scene-sharding host logic: every scene is owned by exactly one rank, and the packed
metrics all-gather reproduces the single-process table in scene order (including uneven shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_rollout(n_scene, seed=0, A=6, K=3, T=20):
    g = torch.Generator().manual_seed(seed)
    preds = torch.randn(n_scene, A, K, T, 4, generator=g)
    valid = torch.rand(n_scene, A, K, T, generator=g) < 0.8
    viol = {k: torch.rand(n_scene, A, K, T, generator=g) < 0.2 for k in ("outside_map", "goal_reached", "dest_reached")}
    rew = -torch.rand(n_scene, A, K, T, generator=g)
    gt_pos = torch.randn(n_scene, T + 1, A, 2, generator=g)
    gt_valid = torch.rand(n_scene, T + 1, A, generator=g) < 0.9
    return preds, valid, viol, rew, gt_pos, gt_valid


def _worker(rank, world, port, n_scene, q):
    from trafficbots_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    preds, valid, viol, rew, gt_pos, gt_valid = _fake_rollout(n_scene)
    b, e = parallel.scene_shard(n_scene, rank, world)
    local = parallel.pack_scene_metrics(preds[b:e], valid[b:e], {k: v[b:e] for k, v in viol.items()}, rew[b:e], gt_pos[b:e],
                                        gt_valid[b:e], step_current=5)
    table = parallel.all_gather_scenes(local, n_scene)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_scene", [4, 5])
def test_two_rank_gather_matches_single_process(n_scene):
    from trafficbots_b200 import parallel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_scene, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    preds, valid, viol, rew, gt_pos, gt_valid = _fake_rollout(n_scene)
    ref = parallel.pack_scene_metrics(preds, valid, viol, rew, gt_pos, gt_valid, step_current=5)
    assert ref.shape == (n_scene, len(parallel.METRIC_FIELDS))
    for r in range(2):
        assert torch.allclose(got[r], ref, atol=1e-6), r


def test_scene_shard_partitions():
    from trafficbots_b200 import parallel
    for n in (1, 7, 32, 256):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                b, e = parallel.scene_shard(n, r, w)
                cover += list(range(b, e))
            assert cover == list(range(n))


def test_mirror_state_dict_and_config_surface():
    from trafficbots_b200 import config, weights
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    m = WaymoMotion(**config.default_config())
    assert set(m.state_dict()) == set(weights.state_dict_spec())
    m.load_state_dict(weights.init_state_dict(1), strict=True)
    # aliases stay tied after loading (shared modules, latent_encoder.py:39-41)
    a = m.model.latent_encoder.transformer_as2pl.layers[0].norm1.weight if hasattr(m.model.latent_encoder.transformer_as2pl.layers, "__getitem__") \
        else getattr(m.model.latent_encoder.transformer_as2pl.layers, "0").norm1.weight
    b = getattr(m.model.transformer_as2pl.layers, "0").norm1.weight
    assert a.data_ptr() == b.data_ptr()
    for bad in ({"model": {"tf_cfg": {"n_head": 8}}}, {"model": {"agent_temporal": {"_target_": "x.MultiAgentGRUCell"}}},
                {"traffic_rule_checker": {"enable_check_collided": True}}, {"dynamics": {"veh": {"max_acc": 4}}},
                {"model": {"goal_manager": {"goal_attr_mode": "goal_xy"}}}):
        with pytest.raises(config.UnsupportedConfig):
            WaymoMotion(**{**config.default_config(), **bad})
    if not torch.cuda.is_available():
        from trafficbots_b200 import _native as nt
        with pytest.raises(nt.TbError):
            m.engine()

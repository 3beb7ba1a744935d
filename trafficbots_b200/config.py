"""Hyper-parameter surface of the fast path.

The CUDA kernels implement the DEFAULT values of the reference's `configs/model/traffic_bots.yaml` for every key
that changes the arithmetic of the hot path.  `check_supported(cfg)` walks a user config
(nested dict / DictConfig-like) and raises `UnsupportedConfig` for any such key with a different value -- the
path never silently diverges from what was asked for.  Keys that do not touch the hot path (optimizer, loggers,
submission writer, ...) are accepted and ignored.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping


class UnsupportedConfig(NotImplementedError):
    pass


_PE = {"map": "pe_xy_yaw", "tl": "pe_xy_yaw", "agent": "pe_xy_yaw"}
_MLP = {"use_layernorm": False, "activation": "relu"}

# value the kernels implement, by config path (reference yaml line numbers: configs/model/traffic_bots.yaml)
REQUIRED: Dict[str, Any] = {
    "hidden_dim": 128,  # :9
    "time_step_sim_start": 1,  # :8
    "pre_processing.input.pe_dim": 96,  # :20
    "pre_processing.input.pose_pe": _PE,  # :21-24
    "pre_processing.latent.perturb_input_to_latent": False,  # :29
    "model.add_goal_latent_first": False,  # :37
    "model.resample_latent": False,  # :38
    "model.n_layer_tf_as2pl": 3,
    "model.n_layer_tf_as2tl": 3,
    "model.tf_cfg.n_head": 4,  # :43
    "model.tf_cfg.norm_first": True,
    "model.tf_cfg.bias": True,
    "model.tf_cfg.activation": "relu",
    "model.tf_cfg.d_feedforward": 128,
    "model.tf_cfg.out_layernorm": False,
    "model.input_pe_encoder.pe_mode": "cat",
    "model.input_pe_encoder.n_layer": 2,
    "model.input_pe_encoder.mlp_use_layernorm": False,
    "model.map_encoder.pool_mode": "max",
    "model.map_encoder.densetnt_vectornet": True,
    "model.map_encoder.n_layer": 3,
    "model.goal_manager.disable_if_reached": True,
    "model.goal_manager.goal_attr_mode": "dest",
    "model.latent_encoder.latent_dim": 16,
    "model.agent_temporal.num_layers": 3,
    "model.agent_interaction.n_layer": 3,
    "model.agent_interaction.mask_self_agent": True,
    "model.agent_interaction.attn_to_map_aware_feature": True,
    "model.add_latent.mode": "cat",
    "model.add_latent.res_cat": False,
    "model.add_latent.res_add": True,
    "model.add_latent.n_layer_mlp_in": 2,
    "model.add_latent.n_layer_mlp_out": 2,
    "model.add_goal.mode": "cat",
    "model.add_goal.res_cat": False,
    "model.add_goal.res_add": True,
    "model.add_goal.n_layer_mlp_in": 3,
    "model.add_goal.n_layer_mlp_out": 2,
    "model.add_goal.mlp_in_cfg.use_layernorm": True,
    "model.interaction_first": True,
    "model.n_layer_final_mlp": -1,
    "action_head.branch_type": True,
    "action_head.use_layernorm": False,
    "dynamics.use_veh_dynamics_for_all": False,
    "dynamics.veh.max_acc": 5,
    "dynamics.veh.max_yaw_rate": 1.5,
    "dynamics.cyc.max_acc": 6,
    "dynamics.cyc.max_yaw_rate": 3,
    "dynamics.ped.max_acc": 7,
    "dynamics.ped.max_yaw_rate": 7,
    "differentiable_reward.use_il_loss": True,
    "differentiable_reward.l_pos.weight": 1e-1,
    "differentiable_reward.l_rot.weight": 1e1,
    "differentiable_reward.l_rot.angular_type": "cosine",
    "differentiable_reward.l_spd.weight": 1e-1,
}
# free keys of the hot path (every value is implemented): traffic_rule_checker.enable_check_{collided, run_road_edge,
# run_red_light, passive} and differentiable_reward.{w_collision, reduce_collsion_with_max} -> `tb_rule_checks`;
# enable_check_passive without enable_check_run_red_light is rejected like the reference does (NameError there).
REQUIRED_SUFFIX = {  # `_target_` class names (module prefix may be the reference's or ours)
    "model.agent_temporal._target_": "MultiAgentGRULoop",
    "dynamics.veh._target_": "MultiPathPP",
    "dynamics.cyc._target_": "MultiPathPP",
    "dynamics.ped._target_": "MultiPathPP",
}


def _lookup(cfg: Mapping, path: str):
    cur: Any = cfg
    for part in path.split("."):
        if not isinstance(cur, Mapping) or part not in cur:
            return _MISSING
        cur = cur[part]
    return cur


_MISSING = object()


def _same(a, b) -> bool:
    if isinstance(b, Mapping):
        return isinstance(a, Mapping) and all(k in a and _same(a[k], v) for k, v in b.items())
    if isinstance(b, bool) or isinstance(a, bool):
        return bool(a) == bool(b)
    if isinstance(b, (int, float)) and isinstance(a, (int, float)):
        return abs(float(a) - float(b)) <= 1e-12
    return a == b


def check_supported(cfg: Mapping) -> None:
    """raises UnsupportedConfig naming every key whose value the CUDA path does not implement (missing keys = default)."""
    bad = []
    for path, want in REQUIRED.items():
        got = _lookup(cfg, path)
        if got is not _MISSING and not _same(got, want):
            bad.append(f"{path}={got!r} (supported: {want!r})")
    for path, want in REQUIRED_SUFFIX.items():
        got = _lookup(cfg, path)
        if got is not _MISSING and not str(got).endswith(want):
            bad.append(f"{path}={got!r} (supported: *.{want})")
    trc = _lookup(cfg, "traffic_rule_checker")
    if isinstance(trc, Mapping) and trc.get("enable_check_passive") and not trc.get("enable_check_run_red_light"):
        bad.append("traffic_rule_checker.enable_check_passive=True needs enable_check_run_red_light=True (the reference raises a "
                   "NameError for this combination: utils/traffic_rule_checker.py:441-442,457-464)")
    if bad:
        raise UnsupportedConfig("trafficbots_b200 implements the default TrafficBots configuration only; unsupported: "
                                + "; ".join(bad))


def default_config(time_step_end: int = 90, n_joint_future: int = 6, rule_checks: bool = False, w_collision: float = 0.0,
                   reduce_collision_with_max: bool = True) -> Dict[str, Any]:
    """constructor kwargs of `WaymoMotion` for the default model (everything the hot path reads); `rule_checks` switches the
    four optional traffic-rule checks on, `w_collision` the collision reward."""
    cur = 10
    cfg = dict(
        time_step_current=cur, time_step_gt=90, time_step_end=time_step_end, time_step_sim_start=1, hidden_dim=128,
        n_joint_future=n_joint_future,
        pre_processing={"input": {"pe_dim": 96, "pose_pe": dict(_PE), "dropout_p_history": -1},
                        "latent": {"pe_dim": 96, "pose_pe": dict(_PE), "perturb_input_to_latent": False, "dropout_p_history": -1}},
        model={"hidden_dim": 128, "tf_cfg": {"n_head": 4, "d_feedforward": 128, "norm_first": True, "dropout_p": 0.1},
               "goal_manager": {"goal_attr_mode": "dest", "disable_if_reached": True},
               "latent_encoder": {"latent_dim": 16, "temporal_down_sample_rate": 5}},
        teacher_forcing_training={"step_spawn_agent": cur, "step_warm_start": cur},
        teacher_forcing_reactive_replay={"step_spawn_agent": 90, "step_warm_start": cur},
        teacher_forcing_joint_future_pred={"step_spawn_agent": cur, "step_warm_start": cur},
        traffic_rule_checker={"enable_check_collided": rule_checks, "enable_check_run_road_edge": rule_checks,
                              "enable_check_run_red_light": rule_checks, "enable_check_passive": rule_checks},
        differentiable_reward={"w_collision": w_collision, "reduce_collsion_with_max": reduce_collision_with_max, "use_il_loss": True},
    )
    return cfg

"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (zhejz/TrafficBots).

Imports the reference's `pl_modules.waymo_motion.WaymoMotion` straight from `/root/reference/src`
(never copied into this repo) by stubbing the third-party packages that are absent in this image
(hydra, omegaconf, pytorch_lightning, torchmetrics, tensorflow, waymo_open_dataset, gym, transforms3d).
None of the stubbed packages performs arithmetic on the hot path: every number the reference produces
comes from `torch`.

Used by
  * `oracle/make_golden.py`        -- generates `tests/golden/*.npz` (reference outputs on seeded inputs)
  * `tests/test_oracle_vs_reference.py` -- pins `oracle/trafficbots_oracle.py` against the reference
Both only work where `/root/reference` exists (the build container); on the GPU box the committed golden
vectors stand in for the reference.  Nothing under `trafficbots_b200/` may import this file.
"""
from __future__ import annotations

import importlib
import inspect
import os
import sys
import types
from typing import Any, Dict

import torch
from torch import nn

def _find_reference() -> str:
    """`TRAFFICBOTS_REF`, else the read-only checkout of the build container, else the git-ignored copy that travels to the
    GPU box with the repo snapshot (`tools/install_reference.sh` -> baseline/_ref/src; used for the GPU-vs-GPU baseline)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("TRAFFICBOTS_REF"), "/root/reference/src", os.path.join(here, "baseline", "_ref", "src")):
        if cand and os.path.isfile(os.path.join(cand, "pl_modules", "waymo_motion.py")):
            return cand
    return os.environ.get("TRAFFICBOTS_REF", "/root/reference/src")


REF_SRC = _find_reference()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "pl_modules", "waymo_motion.py"))


# ----------------------------------------------------------------------------------------------------------
# stubs
# ----------------------------------------------------------------------------------------------------------
class AttrDict(dict):
    """dict with attribute access, recursively (stands in for omegaconf.DictConfig)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            self[k] = _wrap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = _wrap(v)


class AttrList(list):
    def __class_getitem__(cls, item):
        return cls


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, AttrDict):
        return AttrDict(v)
    if isinstance(v, (list, tuple)) and not isinstance(v, AttrList):
        return AttrList(_wrap(x) for x in v)
    return v


def _get_class(path: str):
    mod, name = path.rsplit(".", 1)
    return getattr(importlib.import_module(mod), name)


def _instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    cfg.pop("_convert_", None)
    kwargs.pop("_recursive_", None)
    kwargs.pop("_convert_", None)
    cfg.update(kwargs)
    return _get_class(target)(*args, **{k: _wrap(v) for k, v in cfg.items()})


class _LightningModule(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        self.current_epoch = 0
        self.global_rank = 0
        self.logger = None
        self._hparams = AttrDict()

    @property
    def hparams(self):
        return self._hparams

    def save_hyperparameters(self, *a, **kw):
        frame = inspect.currentframe().f_back
        loc = frame.f_locals
        sig = inspect.signature(type(loc["self"]).__init__)
        self._hparams = AttrDict({k: loc[k] for k in sig.parameters if k != "self" and k in loc})

    def log(self, *a, **kw):
        pass

    def log_dict(self, *a, **kw):
        pass


class _Metric(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = default
        if torch.is_tensor(default):  # like torchmetrics: the state follows the module's device, stays out of the state_dict
            self.register_buffer(name, default.clone(), persistent=False)
        else:
            setattr(self, name, list(default))

    def reset(self):
        for k, v in self._defaults.items():
            setattr(self, k, v.clone().to(getattr(self, k).device) if torch.is_tensor(v) else list(v))

    def forward(self, *a, **kw):
        self.update(*a, **kw)
        return self.compute()


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


class _Anything:
    """permissive placeholder for import-time-only symbols."""

    def __init__(self, *a, **kw):
        pass

    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **kw):
        return _Anything()

    def SerializeToString(self):
        return b""


_INSTALLED = False


def install_stubs() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    if "omegaconf" not in sys.modules:
        _mod("omegaconf", DictConfig=AttrDict, ListConfig=AttrList, OmegaConf=_Anything())
    if "hydra" not in sys.modules:
        utils = _mod("hydra.utils", instantiate=_instantiate, get_class=_get_class)
        _mod("hydra", utils=utils, main=lambda *a, **k: (lambda f: f))
    if "transforms3d" not in sys.modules:
        _mod("transforms3d", euler=_Anything())
    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning.loggers", WandbLogger=_Anything)
        _mod("pytorch_lightning.callbacks", ModelCheckpoint=_Anything, Callback=_Anything)
        _mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
        _mod("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=object,
             Trainer=_Anything, seed_everything=lambda s, **k: torch.manual_seed(s))
    if "torchmetrics" not in sys.modules:
        metric = _mod("torchmetrics.metric", Metric=_Metric)
        _mod("torchmetrics", Metric=_Metric, metric=metric)
    if "tensorflow" not in sys.modules:
        _mod("tensorflow")
    if "waymo_open_dataset" not in sys.modules:
        _mod("waymo_open_dataset")
        _mod("waymo_open_dataset.protos", motion_metrics_pb2=_Anything(), motion_submission_pb2=_Anything())
        _mod("waymo_open_dataset.metrics")
        _mod("waymo_open_dataset.metrics.python")
        _mod("waymo_open_dataset.metrics.python.config_util_py",
             get_breakdown_names_from_motion_config=lambda *a, **k: [])
        _mod("waymo_open_dataset.metrics.ops", py_metrics_ops=_Anything())
    if "gym" not in sys.modules:
        _mod("gym")
        _mod("gym.wrappers")
        _mod("gym.wrappers.monitoring")
        _mod("gym.wrappers.monitoring.video_recorder", ImageEncoder=_Anything)
    # WOMDMetrics.__init__ parses a text proto into a (stubbed) message: make that a no-op
    from google.protobuf import text_format

    text_format.Parse = lambda *a, **k: None
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


# ----------------------------------------------------------------------------------------------------------
# resolved default config == configs/model/traffic_bots.yaml with the ${...} interpolations done by hand
# ----------------------------------------------------------------------------------------------------------
def data_size(n_agent: int = 64, n_pl: int = 1024, n_step: int = 91, n_step_hist: int = 11, n_tl_stop: int = 40,
              n_pl_node: int = 20) -> Dict[str, Any]:
    """The subset of `DataH5womd.tensor_size_*` (src/data_modules/data_h5_womd.py:85-173) the model reads."""
    return {
        "agent/valid": (n_step, n_agent), "agent/pos": (n_step, n_agent, 2), "agent/vel": (n_step, n_agent, 2),
        "agent/spd": (n_step, n_agent, 1), "agent/acc": (n_step, n_agent, 1), "agent/yaw_bbox": (n_step, n_agent, 1),
        "agent/yaw_rate": (n_step, n_agent, 1), "agent/type": (n_agent, 3), "agent/role": (n_agent, 3),
        "agent/size": (n_agent, 3), "map/valid": (n_pl, n_pl_node), "map/type": (n_pl, 11),
        "map/pos": (n_pl, n_pl_node, 2), "map/dir": (n_pl, n_pl_node, 2), "tl_stop/valid": (n_step, n_tl_stop),
        "tl_stop/state": (n_step, n_tl_stop, 5), "tl_stop/pos": (n_step, n_tl_stop, 2),
        "tl_stop/dir": (n_step, n_tl_stop, 2),
    }


def default_config(time_step_end: int = 90, n_joint_future: int = 6) -> Dict[str, Any]:
    hidden_dim = 128
    time_step_current = 10
    pose_pe = {"map": "pe_xy_yaw", "tl": "pe_xy_yaw", "agent": "pe_xy_yaw"}
    mlp_cfg_plain = {"use_layernorm": False, "activation": "relu", "dropout_p": 0.1}
    latent_prior = {"dist_type": "diag_gaus", "n_cat": 8, "log_std": -1, "use_layernorm": False}
    sub = dict(activate=False, interactive_challenge=False, authors=["NAME1", "NAME2"], affiliation="AFFILIATION",
               description="scr_womd", method_link="METHOD_LINK")
    return dict(
        time_step_current=time_step_current, time_step_gt=90, time_step_end=time_step_end, time_step_sim_start=1,
        hidden_dim=hidden_dim, n_video_batch=3, n_joint_future=n_joint_future, interactive_challenge=False,
        pre_processing={
            "scene_centric": {"_target_": "data_modules.scene_centric.SceneCentricPreProcessing"},
            "input": {"_target_": "data_modules.sc_input.SceneCentricInput", "dropout_p_history": -1, "pe_dim": 96,
                      "pose_pe": pose_pe},
            "latent": {"_target_": "data_modules.sc_latent.SceneCentricLatent", "pe_dim": 96, "pose_pe": pose_pe,
                       "perturb_input_to_latent": False, "dropout_p_history": -1, "max_meter": 50.0, "max_rad": 3.14},
        },
        model={
            "_target_": "models.traffic_bots.TrafficBots", "hidden_dim": hidden_dim, "add_goal_latent_first": False,
            "resample_latent": False, "n_layer_tf_as2pl": 3, "n_layer_tf_as2tl": 3,
            "tf_cfg": {"d_model": hidden_dim, "n_head": 4, "dropout_p": 0.1, "norm_first": True, "bias": True,
                       "activation": "relu", "d_feedforward": 128, "out_layernorm": False},
            "input_pe_encoder": {"pe_mode": "cat", "n_layer": 2, "mlp_dropout_p": 0.1, "mlp_use_layernorm": False},
            "map_encoder": {"pool_mode": "max", "densetnt_vectornet": True, "n_layer": 3, "mlp_dropout_p": 0.1,
                            "mlp_use_layernorm": False},
            "goal_manager": {"disable_if_reached": True,
                             "goal_predictor": {"mode": "mlp", "n_layer_gru": 3, "use_layernorm": True,
                                                "res_add_gru": True, "detach_features": True},
                             "goal_attr_mode": "dest", "goal_in_local": True, "dest_detach_map_feature": False},
            "latent_encoder": {"latent_dim": 16, "temporal_down_sample_rate": 5, "shared_post_prior_net": False,
                               "shared_transformer_as": True, "latent_prior": dict(latent_prior),
                               "latent_post": dict(latent_prior)},
            "temporal_aggregate": {"mode": "max_valid"},
            "agent_temporal": {"_target_": "models.modules.agent_temporal.MultiAgentGRULoop", "num_layers": 3,
                               "dropout": 0.1},
            "agent_interaction": {"n_layer": 3, "mask_self_agent": True, "detach_tgt": False,
                                  "attn_to_map_aware_feature": True},
            "add_latent": {"mode": "cat", "res_cat": False, "res_add": True, "n_layer_mlp_in": 2,
                           "n_layer_mlp_out": 2, "mlp_in_cfg": dict(mlp_cfg_plain), "mlp_out_cfg": dict(mlp_cfg_plain)},
            "add_goal": {"mode": "cat", "res_cat": False, "res_add": True, "n_layer_mlp_in": 3, "n_layer_mlp_out": 2,
                         "mlp_in_cfg": {"use_layernorm": True, "activation": "relu", "dropout_p": 0.1},
                         "mlp_out_cfg": dict(mlp_cfg_plain)},
            "interaction_first": True, "n_layer_final_mlp": -1, "final_mlp": dict(mlp_cfg_plain),
        },
        teacher_forcing_training={"step_spawn_agent": time_step_current, "step_warm_start": time_step_current,
                                  "step_horizon": 0, "step_horizon_decrease_per_epoch": 0, "prob_forcing_agent": 0,
                                  "prob_forcing_agent_decrease_per_epoch": 0},
        action_head={"log_std": -2, "branch_type": True, "use_layernorm": False},
        dynamics={"use_veh_dynamics_for_all": False,
                  "veh": {"_target_": "utils.dynamics.MultiPathPP", "max_acc": 5, "max_yaw_rate": 1.5,
                          "disable_neg_spd": False},
                  "cyc": {"_target_": "utils.dynamics.MultiPathPP", "max_acc": 6, "max_yaw_rate": 3,
                          "disable_neg_spd": False},
                  "ped": {"_target_": "utils.dynamics.MultiPathPP", "max_acc": 7, "max_yaw_rate": 7}},
        differentiable_reward={"w_collision": 0, "reduce_collsion_with_max": True, "use_il_loss": True,
                               "l_pos": {"weight": 1e-1, "criterion": "SmoothL1Loss"},
                               "l_rot": {"weight": 1e1, "criterion": "SmoothL1Loss", "angular_type": "cosine"},
                               "l_spd": {"weight": 1e-1, "criterion": "SmoothL1Loss"}},
        step_detach_hidden=-1, p_drop_hidden=-1.0, p_training_rollout_prior=0.1, detach_state_policy=True,
        training_deterministic_action=True,
        waymo_post_processing={"k_pred": 6, "use_ade": True, "score_temperature": 1e2, "mpa_nms_thresh": [],
                               "mtr_nms_thresh": [], "aggr_thresh": [], "n_iter_em": 3},
        sub_womd_reactive_replay=dict(sub, k_futures=1, method_name="reactive_replay"),
        sub_womd_joint_future_pred=dict(sub, k_futures=6, method_name="joint_future_pred"),
        training_metrics={"w_vae_kl": 1e-1, "kl_balance_scale": -1, "kl_free_nats": 1e-2, "kl_for_unseen_agent": True,
                          "w_diffbar_reward": 1.0, "w_goal": 1.0, "w_relevant_agent": 0, "p_loss_for_irrelevant": -1.0,
                          "loss_for_teacher_forcing": True, "step_training_start": 10},
        optimizer={"_target_": "torch.optim.Adam", "lr": 3e-4}, lr_goal=3e-4,
        lr_scheduler={"_target_": "torch.optim.lr_scheduler.StepLR", "gamma": 0.5, "step_size": 7},
        teacher_forcing_reactive_replay={"step_spawn_agent": 90, "step_warm_start": time_step_current},
        teacher_forcing_joint_future_pred={"step_spawn_agent": time_step_current,
                                           "step_warm_start": time_step_current},
        traffic_rule_checker={"enable_check_collided": False, "enable_check_run_road_edge": False,
                              "enable_check_run_red_light": False, "enable_check_passive": False},
    )


def _deep_update(dst: Dict[str, Any], src: Dict[str, Any]) -> None:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        else:
            dst[k] = v


def build_reference(n_agent: int = 64, n_pl: int = 1024, time_step_end: int = 90, n_joint_future: int = 6,
                    seed: int = 2023, overrides: Dict[str, Any] = None):
    """Instantiate the unmodified reference `WaymoMotion` with default init under `torch.manual_seed(seed)`
    (`configs/run.yaml:17`), in eval mode."""
    install_stubs()
    from pl_modules.waymo_motion import WaymoMotion  # noqa: the reference's own class

    torch.manual_seed(seed)
    plain = default_config(time_step_end=time_step_end, n_joint_future=n_joint_future)
    if overrides:  # e.g. {"traffic_rule_checker": {"enable_check_collided": True}, "differentiable_reward": {"w_collision": 1.0}}
        _deep_update(plain, overrides)
    cfg = _wrap(plain)
    model = WaymoMotion(data_size=_wrap(data_size(n_agent=n_agent, n_pl=n_pl)), **cfg)
    model.eval()
    return model

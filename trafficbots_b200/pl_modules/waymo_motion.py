"""`WaymoMotion` -- drop-in shell with the constructor kwargs, method names and return types of the reference's
LightningModule (`src/pl_modules/waymo_motion.py:27-572`) for the hot path: `forward` (one decode step), `rollout`,
`reactive_replay`, `joint_future_pred`, plus `encode` helpers.  Select it with
`model._target_=trafficbots_b200.pl_modules.waymo_motion.WaymoMotion` (`configs/model/traffic_bots_b200.yaml`).

All model arithmetic runs in `libtrafficbots_b200.so` (see `include/trafficbots_b200.h`); this file is tensor plumbing.
Differences to the reference, all deliberate and documented in DESIGN.md:
  * scene-level tensors are NOT `repeat_interleave`d per joint future; the kernels index scene = scene_mode // K.
  * `rollout()` runs all steps inside the library; the traffic-rule checks, kill, goal disabling and the imitation
    reward are part of the fused step, so the `rule_checker` argument only carries the raw tensors they need.
  * `require_vis_dict=True` / `need_weights` (attention maps for videos) are not provided by the fused path and raise.
If pytorch_lightning is importable the class derives from `LightningModule`, otherwise from `nn.Module`.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from .. import _native as nt
from .. import config as tb_config
from .. import host, weights
from ..data_modules.scene_centric import SceneCentricInput, SceneCentricLatent, SceneCentricPreProcessing
from ..data_modules.waymo_post_processing import WaymoPostProcessing
from ..engine import Engine, gt_from_batch, raw_map_from_batch
from ..models.metrics.womd import WOMDMetrics
from ..models.distributions import DestCategorical, DiagGaussian
from ..models.traffic_bots import TrafficBots, register_param_tree
from ..utils.buffer import RolloutBuffer

try:  # pragma: no cover - Lightning is optional in this image
    from pytorch_lightning import LightningModule as _Base
except Exception:  # noqa: BLE001
    _Base = nn.Module


class _Recorder:
    """default stand-in for the reference's torchmetrics / submission objects (`err_metrics_*`, `rule_metrics_*`,
    `train_metrics_*`, `sub_womd_*`): remembers the keyword tensors of the last call.  Replace the attribute with the
    reference's own object to get its behaviour -- the call sites pass exactly the reference's keywords."""

    def __init__(self, name: str) -> None:
        self.name, self.last, self.n_call = name, None, 0

    def __call__(self, *args, **kw):
        self.last, self.n_call = (args, kw), self.n_call + 1
        return {}

    add_to_submissions = __call__

    def reset(self) -> None:
        pass


class _HParams(dict):
    __getattr__ = dict.__getitem__


class TeacherForcing:
    """utils/teacher_forcing.py:9-74 without the (default-off) schedules."""

    def __init__(self, step_spawn_agent: int = 10, step_warm_start: int = 10, step_horizon: int = 0,
                 step_horizon_decrease_per_epoch: int = 0, prob_forcing_agent: float = 0,
                 prob_forcing_agent_decrease_per_epoch: float = 0) -> None:
        if step_horizon or prob_forcing_agent:
            raise tb_config.UnsupportedConfig("teacher-forcing schedules (step_horizon / prob_forcing_agent) are not supported")
        self.step_spawn_agent, self.step_warm_start = step_spawn_agent, step_warm_start

    def get(self, as_valid: Tensor, current_epoch: int = 0) -> Tensor:
        return host.teacher_forcing_mask(as_valid, self.step_spawn_agent, self.step_warm_start)


class TrafficRuleChecker:
    """Carrier of the raw tensors the rule checks read, with the constructor signature of
    utils/traffic_rule_checker.py:16-44.  The always-on checks (outside_map, goal / dest reached) run inside the fused decode
    step; the four optional ones (`enable_check_*`) are evaluated right after the rollout by `tb_rule_checks`, which gives
    the same result because none of them feeds back into the simulation."""

    def __init__(self, map_boundary, map_valid, map_type, map_pos, map_dir, tl_stop_valid=None, tl_stop_pos=None,
                 tl_stop_state=None, agent_type=None, agent_size=None, agent_goal=None, agent_dest=None,
                 enable_check_collided=False, enable_check_run_road_edge=False, enable_check_run_red_light=False,
                 enable_check_passive=False, collision_size_scale: float = 1.1) -> None:
        if enable_check_passive and not enable_check_run_red_light:
            raise tb_config.UnsupportedConfig("enable_check_passive without enable_check_run_red_light raises a NameError in the "
                                              "reference (traffic_rule_checker.py:441-442,457-464)")
        self.raw_map = {"boundary": map_boundary, "valid": map_valid, "type": map_type, "pos": map_pos, "dir": map_dir}
        self.agent_goal, self.agent_dest = agent_goal, agent_dest
        self.enable = {"collided": bool(enable_check_collided), "run_road_edge": bool(enable_check_run_road_edge),
                       "run_red_light": bool(enable_check_run_red_light), "passive": bool(enable_check_passive)}
        self.tl = None
        if tl_stop_valid is not None:
            self.tl = {"valid": tl_stop_valid, "pos": tl_stop_pos, "state": tl_stop_state}
        self.collision_size_scale = collision_size_scale

    @property
    def any_optional(self) -> bool:
        return any(self.enable.values())


class WaymoMotion(_Base):
    def __init__(self, time_step_current: int = 10, time_step_gt: int = 90, time_step_end: int = 90,
                 time_step_sim_start: int = 1, hidden_dim: int = 128, data_size: Optional[Mapping] = None,
                 pre_processing: Optional[Mapping] = None, step_detach_hidden: int = -1, model: Optional[Mapping] = None,
                 p_training_rollout_prior: float = 0.1, detach_state_policy: bool = True,
                 training_deterministic_action: bool = True, differentiable_reward: Optional[Mapping] = None,
                 p_drop_hidden: float = -1.0, n_video_batch: int = 3, n_joint_future: int = 6,
                 waymo_post_processing: Optional[Mapping] = None, dynamics: Optional[Mapping] = None,
                 action_head: Optional[Mapping] = None, teacher_forcing_training: Optional[Mapping] = None,
                 teacher_forcing_reactive_replay: Optional[Mapping] = None,
                 teacher_forcing_joint_future_pred: Optional[Mapping] = None, training_metrics: Optional[Mapping] = None,
                 traffic_rule_checker: Optional[Mapping] = None, optimizer: Optional[Mapping] = None,
                 lr_scheduler: Optional[Mapping] = None, lr_goal: float = 3e-4,
                 sub_womd_reactive_replay: Optional[Mapping] = None, sub_womd_joint_future_pred: Optional[Mapping] = None,
                 interactive_challenge: bool = False, wb_artifact: Optional[str] = None) -> None:
        super().__init__()
        cfg = dict(hidden_dim=hidden_dim, time_step_sim_start=time_step_sim_start, pre_processing=pre_processing or {},
                   model=model or {}, differentiable_reward=differentiable_reward or {}, dynamics=dynamics or {},
                   action_head=action_head or {}, traffic_rule_checker=traffic_rule_checker or {})
        tb_config.check_supported(cfg)
        if not detach_state_policy:
            raise tb_config.UnsupportedConfig("detach_state_policy=False")
        self.tb_hparams = dict(time_step_current=time_step_current, time_step_gt=time_step_gt, time_step_end=time_step_end,
                               time_step_sim_start=time_step_sim_start, n_joint_future=n_joint_future,
                               traffic_rule_checker=dict(traffic_rule_checker or {}),
                               w_collision=float((differentiable_reward or {}).get("w_collision", 0) or 0),
                               reduce_collision_with_max=bool((differentiable_reward or {}).get("reduce_collsion_with_max", True)))
        spec = weights.state_dict_spec()
        # pre_processing: the reference's Sequential of three modules, under the same names (state_dict keys
        # `pre_processing.{input,latent}.*`), waymo_motion.py:66-72
        pp = dict(pre_processing or {})
        drop = ("_target_", "_recursive_", "_convert_")
        ppk = lambda name: {k: v for k, v in dict(pp.get(name) or {}).items() if k not in drop}  # noqa: E731
        self.pre_processing = nn.Sequential(OrderedDict([
            ("scene_centric", SceneCentricPreProcessing(time_step_current=time_step_current, data_size=data_size)),
            ("input", SceneCentricInput(time_step_current=time_step_current, data_size=data_size, **ppk("input"))),
            ("latent", SceneCentricLatent(time_step_current=time_step_current, data_size=data_size, **ppk("latent")))]))
        mcfg = {k: v for k, v in dict(model or {}).items() if k not in ("_target_", "hidden_dim")}
        self.model = TrafficBots(hidden_dim=hidden_dim, **mcfg)
        self.model.set_owner(self)
        self.action_head = nn.Module()
        register_param_tree(self.action_head, spec, "action_head.")
        self.teacher_forcing_training = TeacherForcing(**(teacher_forcing_training or {}))
        self.teacher_forcing_reactive_replay = TeacherForcing(**(teacher_forcing_reactive_replay or {"step_spawn_agent": 90}))
        self.teacher_forcing_joint_future_pred = TeacherForcing(**(teacher_forcing_joint_future_pred or {}))
        wpp = {k: v for k, v in dict(waymo_post_processing or {}).items() if k not in ("_target_",)}
        self.waymo_post_processing = WaymoPostProcessing(**wpp)
        self.womd_metrics_reactive_replay = WOMDMetrics("reactive_replay", time_step_end, time_step_current, interactive_challenge)
        self.womd_metrics_joint_future_pred = WOMDMetrics("joint_future_pred", time_step_end, time_step_current, interactive_challenge)
        # optional consumers a caller may attach (the reference's torchmetrics / submission objects are outside the hot path):
        # callables `sink(name, **tensors)` invoked by validation_step / test_step with what the reference passes to
        # err_metrics / rule_metrics / train_metrics / sub_womd (waymo_motion.py:613-668,691-733,936-944)
        self.metric_sinks = []
        for leg in ("reactive_replay", "joint_future_pred"):
            setattr(self, f"err_metrics_{leg}", _Recorder(f"err_metrics_{leg}"))
            setattr(self, f"rule_metrics_{leg}", _Recorder(f"rule_metrics_{leg}"))
            setattr(self, f"sub_womd_{leg}", _Recorder(f"sub_womd_{leg}"))
        self.train_metrics_reactive_replay = _Recorder("train_metrics_reactive_replay")
        if _Base is nn.Module:  # what LightningModule would provide; the reference's step methods read them
            self.__dict__["hparams"] = _HParams(time_step_current=time_step_current, time_step_gt=time_step_gt, time_step_end=time_step_end,
                                                time_step_sim_start=time_step_sim_start, n_video_batch=n_video_batch,
                                                n_joint_future=n_joint_future, interactive_challenge=interactive_challenge)
            self.__dict__["current_epoch"] = 0
            self.__dict__["global_rank"] = 0
            self.__dict__["logger"] = None
        self._eng: Optional[Engine] = None
        self._eng_slot: Optional[Engine] = None
        self._packed_version = None
        self._params_dirty = True
        self._param_list = None
        self._param_ptrs = None
        self._step_ctx = None
        self._train_state = None
        self.__dict__["automatic_optimization"] = True  # False: training_step leaves the gradients in p.grad and stops there
        tm = dict(training_metrics or {})
        defaults = dict(w_vae_kl=0.1, kl_balance_scale=-1, kl_free_nats=0.01, kl_for_unseen_agent=True, w_diffbar_reward=1.0, w_goal=1.0,
                        w_relevant_agent=0, p_loss_for_irrelevant=-1.0, loss_for_teacher_forcing=True, step_training_start=10)
        for k, v in defaults.items():  # the loss of train/graph.py implements the default TrainingMetrics configuration
            if k in tm and tm[k] != v and not (isinstance(v, float) and abs(float(tm[k]) - v) < 1e-12):
                raise tb_config.UnsupportedConfig(f"training_metrics.{k}={tm[k]!r} (supported: {v!r})")
        if step_detach_hidden > 0 or p_drop_hidden > 0 or not training_deterministic_action:
            raise tb_config.UnsupportedConfig("step_detach_hidden / p_drop_hidden / stochastic training actions are not supported")
        opt, sch = dict(optimizer or {}), dict(lr_scheduler or {})
        if opt.get("_target_", "torch.optim.Adam") != "torch.optim.Adam" or sch.get("_target_", "torch.optim.lr_scheduler.StepLR") != \
                "torch.optim.lr_scheduler.StepLR":
            raise tb_config.UnsupportedConfig("optimizer / lr_scheduler other than Adam / StepLR")
        # dropout of the training step: the reference has ONE probability at all its sites (tf_cfg.dropout_p, *.mlp_dropout_p,
        # mlp_*_cfg.dropout_p, agent_temporal.dropout: 0.1 in traffic_bots.yaml); set `train_dropout_p = 0` for the parity runs
        mc = dict(model or {})
        ps = {float((mc.get("tf_cfg") or {}).get("dropout_p", 0.1)), float((mc.get("input_pe_encoder") or {}).get("mlp_dropout_p", 0.1)),
              float((mc.get("map_encoder") or {}).get("mlp_dropout_p", 0.1)), float((mc.get("agent_temporal") or {}).get("dropout", 0.1)),
              float(((mc.get("add_latent") or {}).get("mlp_in_cfg") or {}).get("dropout_p", 0.1)),
              float(((mc.get("add_goal") or {}).get("mlp_in_cfg") or {}).get("dropout_p", 0.1))}
        if len(ps) != 1:
            raise tb_config.UnsupportedConfig(f"different dropout probabilities at different sites are not supported: {sorted(ps)}")
        self.__dict__["train_dropout_p"] = ps.pop()
        self._train_hparams = dict(lr=float(opt.get("lr", 3e-4)), lr_goal=float(lr_goal), max_grad_norm=5.0,
                                   p_training_rollout_prior=float(p_training_rollout_prior), lr_gamma=float(sch.get("gamma", 0.5)),
                                   lr_step_size=int(sch.get("step_size", 7)))

    # ------------------------------------------------------------------------------------------------ engine / parameters
    def _param_version(self):
        """(tensor identities, version counters) of the 483 state_dict entries.  The tensor list is cached (rebuilt whenever
        `_apply` / `load_state_dict` / `mark_params_dirty` flag a change), so the per-call cost is one attribute read per
        tensor, not a `state_dict()` rebuild."""
        if self._param_list is None or self._params_dirty:
            self._param_list = list(self.state_dict(keep_vars=True).values())
            self._param_ptrs = tuple(p.data_ptr() for p in self._param_list)
        return self._param_ptrs, tuple(p._version for p in self._param_list)

    def mark_params_dirty(self) -> None:
        """Call after writing parameters in a way autograd's version counters do not see (`p.data.copy_()`, EMA updates
        through `.data`, raw pointer writes): the next `engine()` call re-packs the kernel weight blob.  `load_state_dict`,
        `.to()` / `.cuda()` and in-place updates (optimizer steps, `p.add_()`, `p.copy_()`) are detected automatically."""
        self._params_dirty = True

    def _apply(self, fn, *a, **kw):
        self._params_dirty = True
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._params_dirty = True
        return super()._load_from_state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        self._params_dirty = True
        return super().load_state_dict(*a, **kw)

    def engine(self) -> Engine:
        """the CUDA engine with the CURRENT parameters packed: autograd's version counters of the (cached) parameter list are
        compared on every call -- except inside an open stepwise rollout -- and the blob is re-packed when they moved (see
        `mark_params_dirty` for writes the counters do not see).  Inside `pipeline.ScenePipeline` slots the slot's forked
        engine (same packed blob, own workspaces) is returned."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise nt.TbError("trafficbots_b200.WaymoMotion must live on a CUDA device (no CPU implementation of the hot path)")
        if self._eng is None or self._eng.device != torch.device(dev.type, dev.index if dev.index is not None else torch.cuda.current_device()):
            self._eng = Engine(self.state_dict(), dev)
            self._packed_version = self._param_version()
            self._params_dirty = False
            self._eng_slot = None
        elif self._step_ctx is None:
            ver = self._param_version()
            if ver != self._packed_version or self._params_dirty:
                self._eng.load_state_dict(self.state_dict())
                self._packed_version = ver
            self._params_dirty = False
        return self._eng_slot if self._eng_slot is not None else self._eng

    def use_engine(self, eng: Optional[Engine]) -> None:
        """route the next calls through a forked engine (`Engine.fork`) -- used by `pipeline.ScenePipeline`; None restores
        the module's own engine."""
        self._eng_slot = eng

    # ------------------------------------------------------------------------------------------------ encoding
    def encode(self, batch: Mapping[str, Tensor]):
        """`pre_processing` + `model.encode_input_features` of validation_step / test_step (waymo_motion.py:576-583)."""
        return self.model.encode_input_features(batch)

    # ------------------------------------------------------------------------------------------------ rollout
    def rollout(self, features: Mapping[str, Tensor], latent: Union[DiagGaussian, Tensor], goal: Tensor, goal_valid: Tensor,
                mask_teacher_forcing: Tensor, rule_checker: TrafficRuleChecker,
                deterministic_latent: Union[bool, Tensor], deterministic_action: bool, step_end: int, step_start: int,
                require_vis_dict: bool = False, gt_sdc=None) -> RolloutBuffer:
        """waymo_motion.py:205-354.  `features`: the reference's keys (`map_valid`, `map_feature`, `tl_valid`, `tl_feature`,
        `agent_valid`, `agent_state`, `agent_type`, `agent_size`, `vel`, `acc`, `yaw_rate`) with leading dim n_scene, plus
        the K|V caches of `encode` (`_kv_map`, `_kv_tl`) and optionally `_n_mode` (joint futures per scene);
        `goal` [n_scene*n_mode, n_agent] int64 destination index; latent / goal_valid per scene-mode."""
        if require_vis_dict or gt_sdc is not None or not deterministic_action:
            raise tb_config.UnsupportedConfig("require_vis_dict / gt_sdc / stochastic actions are not supported by the fused rollout")
        if step_start != 1:
            raise tb_config.UnsupportedConfig("step_start must be 1")
        n_mode = int(features.get("_n_mode", 1))
        self.model.init(latent, deterministic_latent)
        eng = self.engine()
        if "_gt" in features:  # un-concatenated GT tensors straight from the batch (no copies)
            gt = features["_gt"]
        else:
            st = features["agent_state"]
            gt = {"valid": features["agent_valid"], "pos": st[..., :2].contiguous(), "yaw_bbox": st[..., 2:3].contiguous(),
                  "spd": st[..., 3:4].contiguous(), "vel": features["vel"], "acc": features["acc"], "yaw_rate": features["yaw_rate"]}
        feat = {"map_feature": features["map_feature"], "map_feature_valid": features["map_valid"],
                "tl_feature_valid": features["tl_valid"], "_kv_map": features["_kv_map"], "_kv_tl": features["_kv_tl"]}
        for k in ("_kv_map_tc", "_kv_tl_tc", "_n_key_map", "_n_key_tl"):
            if k in features:
                feat[k] = features[k]
        args = (feat, gt, mask_teacher_forcing, features["agent_type"], features["agent_size"], rule_checker.raw_map,
                self.model.latent_sample, self.model.latent_logp, goal.contiguous(), goal_valid.contiguous(), rule_checker.agent_goal)
        post = dict(gt=gt, agent_type=features["agent_type"], agent_size=features["agent_size"], rc=rule_checker, n_mode=n_mode)
        if features.get("_stepwise", False):  # caller drives the steps through forward()
            self._step_ctx = eng.begin_rollout(*args, n_mode=n_mode, n_step=step_end)
            self._step_ctx["post"] = post
            return None
        out = self._optional_checks(eng, eng.rollout(*args, n_mode=n_mode, n_step=step_end), post)
        return RolloutBuffer(step_start, step_end, self.tb_hparams["time_step_current"], out)

    def _optional_checks(self, eng: Engine, out: Dict[str, Tensor], post: Mapping) -> Dict[str, Tensor]:
        """optional traffic-rule checks + collision reward on the finished rollout (`tb_rule_checks`)."""
        rc: TrafficRuleChecker = post["rc"]
        w = self.tb_hparams["w_collision"]
        if not rc.any_optional and w <= 0:
            return out
        return eng.rule_checks(out, post["gt"], post["agent_type"], post["agent_size"], rc.raw_map, rc.tl, rc.enable,
                               n_mode=post["n_mode"], w_collision=w,
                               reduce_collision_with_max=self.tb_hparams["reduce_collision_with_max"],
                               collision_size_scale=rc.collision_size_scale)

    def forward(self, map_feature: Optional[Tensor] = None, map_valid: Optional[Tensor] = None,
                tl_feature: Optional[Tensor] = None, tl_valid: Optional[Tensor] = None, goal_feature: Optional[Tensor] = None,
                goal_valid: Optional[Tensor] = None, action_override: Optional[Tensor] = None,
                mask_action_override: Optional[Tensor] = None, state_override: Optional[Mapping] = None,
                mask_state_override: Optional[Tensor] = None, deterministic_action: bool = True,
                require_train_dict: bool = True, require_vis_dict: bool = False):
        """One decode step (signature of waymo_motion.py:108-123) of the rollout opened with
        `rollout(features | {"_stepwise": True}, ...)`.  The reference passes the step's map / traffic-light features and the
        goal feature as arguments; in the fused path they are BOUND when the rollout is opened (K|V caches, goal features,
        `goal_valid` bookkeeping live in the engine state), so the first six arguments are accepted and not re-read.  The
        state override of step t is `GT[t]` under `mask_teacher_forcing[:, t]` -- what the reference's `rollout()` passes
        (:300-309) -- also bound at opening.  A caller that passes overrides of its own (`action_override`,
        `state_override`, `mask_state_override`), stochastic actions or asks for attention maps gets `UnsupportedConfig`
        instead of a silently different result.  Returns (state [B,A,4], valid [B,A], train_dict, vis_dict); `train_dict`
        holds this step's pre-override prediction under the reference's keys (:181-188)."""
        if self._step_ctx is None:
            raise nt.TbError("forward(): open a rollout first (features['_stepwise'] = True)")
        if action_override is not None or mask_action_override is not None or state_override is not None or \
                mask_state_override is not None:
            raise tb_config.UnsupportedConfig("forward(): per-call action / state overrides are not supported by the fused step; the "
                                              "overrides are the GT tensors and `mask_teacher_forcing` bound by rollout()")
        if require_vis_dict or not deterministic_action:
            raise tb_config.UnsupportedConfig("forward(): require_vis_dict / stochastic actions are not supported by the fused step")
        eng = self.engine()
        t = eng.step(self._step_ctx)
        o = self._step_ctx["out"]
        train = {} if not require_train_dict else {
            "latent_log_prob": o["latent_log_probs"][:, :, t - 1], "action_log_prob": o["action_log_probs"][:, :, t - 1],
            "pred_valid": o["valid"][:, :, t - 1], "pred_state": o["preds"][:, :, t - 1]}
        return eng.state_field(nt.STATE_AGENT_STATE), eng.state_field(nt.STATE_VALID), train, {}

    def finish_rollout(self) -> RolloutBuffer:
        ctx, self._step_ctx = self._step_ctx, None
        eng = self.engine()
        out = self._optional_checks(eng, eng._finish(ctx["out"]), ctx["post"])
        return RolloutBuffer(1, ctx["n_step"], self.tb_hparams["time_step_current"], out)

    def _features(self, batch: Mapping[str, Tensor], f: Mapping[str, Tensor], n_mode: int, n_gt: Optional[int] = None) -> Dict:
        gt = gt_from_batch(batch, n_gt)
        return {"map_valid": f["map_feature_valid"], "map_feature": f["map_feature"], "tl_valid": f["tl_feature_valid"],
                "tl_feature": f["tl_feature"], "agent_type": batch["history/agent/type"], "agent_size": batch["history/agent/size"],
                "agent_valid": gt["valid"], "vel": gt["vel"], "acc": gt["acc"], "yaw_rate": gt["yaw_rate"], "agent_state": None,
                "_gt": gt, "_kv_map": f["_kv_map"], "_kv_tl": f["_kv_tl"], "_n_mode": n_mode,
                **{k: f[k] for k in ("_kv_map_tc", "_kv_tl_tc", "_n_key_map", "_n_key_tl") if k in f}}

    def reactive_replay(self, batch: Mapping[str, Tensor], input_feature_dict: Mapping[str, Tensor], mask_teacher_forcing: Tensor,
                        latent, goal: Optional[Tensor], goal_valid: Optional[Tensor], deterministic_latent: bool,
                        deterministic_action: bool, require_vis_dict: bool = False) -> RolloutBuffer:
        """waymo_motion.py:420-476: one rollout per scene with GT destination, spawning over the whole episode."""
        rc = TrafficRuleChecker(batch["map/boundary"], batch["map/valid"], batch["map/type"], batch["map/pos"], batch["map/dir"],
                                batch.get("tl_stop/valid"), batch.get("tl_stop/pos"), batch.get("tl_stop/state"),
                                agent_goal=batch.get("agent/goal"), agent_dest=batch.get("agent/dest"),
                                **self.tb_hparams["traffic_rule_checker"])  # full-episode traffic lights (:439-441)
        feats = self._features(batch, input_feature_dict, 1)
        return self.rollout(feats, latent=latent, goal=goal, goal_valid=goal_valid, mask_teacher_forcing=mask_teacher_forcing,
                            rule_checker=rc, step_start=self.tb_hparams["time_step_sim_start"],
                            step_end=self.tb_hparams["time_step_end"], deterministic_latent=deterministic_latent,
                            deterministic_action=deterministic_action, require_vis_dict=require_vis_dict)

    def joint_future_pred(self, batch: Mapping[str, Tensor], input_feature_dict: Mapping[str, Tensor], latent: DiagGaussian,
                          goal: DestCategorical, goal_valid: Tensor, require_vis_dict: bool = False
                          ) -> Tuple[RolloutBuffer, Tensor, Tensor]:
        """waymo_motion.py:478-572: K joint futures per scene; mode 0 is deterministic (prior mean, arg-max destination).
        Test mode (only the 11 history frames as GT, no goal check) is selected like in the reference's `test_step` by
        a batch whose `agent/valid` has 11 frames (waymo_motion.py:923-924)."""
        K = self.tb_hparams["n_joint_future"]
        S, A = batch["history/agent/valid"].shape[0], batch["history/agent/valid"].shape[2]
        det = torch.zeros(S * K, A, dtype=torch.bool, device=goal_valid.device)
        det[::K] = True
        latent.repeat_interleave_(K, 0)
        goal.repeat_interleave_(K, 0)
        goal_sample = goal.sample(det)
        goal_log_probs = goal.log_prob(goal_sample)
        gvalid = goal_valid.repeat_interleave(K, 0)
        rc = TrafficRuleChecker(batch["map/boundary"], batch["map/valid"], batch["map/type"], batch["map/pos"], batch["map/dir"],
                                batch.get("history/tl_stop/valid"), batch.get("history/tl_stop/pos"),
                                batch.get("history/tl_stop/state"),  # history traffic lights, frozen after frame 10 (:523-525)
                                agent_goal=batch.get("agent/goal"), agent_dest=goal_sample, **self.tb_hparams["traffic_rule_checker"])
        feats = self._features(batch, input_feature_dict, K)
        tf = self.teacher_forcing_joint_future_pred.get(feats["agent_valid"], 0)
        buf = self.rollout(feats, latent=latent, goal=goal_sample, goal_valid=gvalid, mask_teacher_forcing=tf, rule_checker=rc,
                           step_start=self.tb_hparams["time_step_sim_start"], step_end=self.tb_hparams["time_step_end"],
                           deterministic_latent=det, deterministic_action=True, require_vis_dict=require_vis_dict)
        buf.flatten_repeat(K)
        return buf, goal_sample.view(S, K, A).transpose(1, 2), goal_log_probs.view(S, K, A).transpose(1, 2)

    # ------------------------------------------------------------------------------------------------ eval loop
    def _split(self, batch: Mapping, group: str) -> Dict:
        return {k.split(group + "/")[-1]: v for k, v in batch.items() if (group + "/") in k}

    def _emit(self, name: str, **tensors) -> None:
        for sink in self.metric_sinks:
            sink(name, **tensors)

    if _Base is nn.Module:  # LightningModule.log stand-in: the last value of every logged key (device scalars are not synchronised)
        def log(self, name: str, value, **kw) -> None:
            self.__dict__.setdefault("logged", {})[name] = value

    @torch.no_grad()
    def validation_step(self, batch: Dict[str, Tensor], batch_idx: int = 0) -> Dict:
        """waymo_motion.py:574-733 without the video logging: pre_processing, the three `encode_input_features` calls (the
        aliased ones are served from the first), GT / predicted destination, posterior / prior latent, `reactive_replay`,
        `joint_future_pred`, Waymo post-processing and the WOMD packing of both legs.  Returns what the reference feeds its
        metric objects; the same tensors are pushed to `self.metric_sinks`."""
        batch = self.pre_processing(batch)
        input_dict, post_dict, prior_dict = self._split(batch, "input"), self._split(batch, "latent_post"), self._split(batch, "latent_prior")
        input_feature_dict = self.model.encode_input_features(**input_dict)
        latent_post_feature_dict = self.model.encode_input_features(**post_dict)
        latent_prior_feature_dict = self.model.encode_input_features(**prior_dict)
        goal_gt, goal_valid = self.model.goal_manager.get_gt_goal(agent_valid=input_dict["agent_valid"], gt_dest=batch["gt/dest"],
                                                                  gt_goal=batch["gt/goal"])
        goal_pred = self.model.goal_manager.pred_goal(agent_type=batch["ref/agent_type"], map_type=batch["ref/map_type"],
                                                      agent_state=batch["ref/agent_state"], **input_feature_dict)
        latent_post = self.model.latent_encoder(posterior=True, **latent_post_feature_dict)
        latent_prior = self.model.latent_encoder(**latent_prior_feature_dict)
        t0 = self.tb_hparams["time_step_sim_start"]
        gt_valid = batch["gt/valid"][:, t0:].transpose(1, 2)
        gt_states = batch["gt/state"][:, t0:].transpose(1, 2)
        out: Dict = {"goal_pred": goal_pred, "goal_gt": goal_gt, "goal_valid": goal_valid, "latent_post": latent_post,
                     "latent_prior": latent_prior}

        # ! reactive_replay: scene reconstruction given the complete episode (:597-668)
        buf = self.reactive_replay(batch=batch, input_feature_dict=input_feature_dict,
                                   mask_teacher_forcing=self.teacher_forcing_reactive_replay.get(batch["gt/valid"], 0),
                                   latent=latent_post, goal=goal_gt, goal_valid=goal_valid, deterministic_latent=True,
                                   deterministic_action=True, require_vis_dict=False)
        buf.flatten_repeat(1)
        self._emit("err_metrics_reactive_replay", pred_valid=buf.valid, pred_states=buf.preds, gt_valid=gt_valid, gt_states=gt_states,
                   override_masks=buf.override_masks, agent_role=batch["ref/agent_role"])
        self._emit("rule_metrics_reactive_replay", valid=buf.valid, override_masks=buf.override_masks, agent_type=batch["ref/agent_type"],
                   **{k: buf.violations[k] for k in ("outside_map", "collided", "run_road_edge", "run_red_light", "passive",
                                                     "goal_reached", "dest_reached")})
        self._emit("train_metrics_reactive_replay", pred_valid=buf.valid.squeeze(2), diffbar_rewards_valid=buf.diffbar_rewards_valid.squeeze(2),
                   diffbar_rewards=buf.diffbar_rewards.squeeze(2), override_masks=buf.override_masks.squeeze(2),
                   agent_role=batch["ref/agent_role"], goal_valid=goal_valid, goal_pred=goal_pred, goal_gt=goal_gt,
                   latent_post=latent_post, latent_prior=latent_prior)
        pred_dict = self.waymo_post_processing(valid=buf.valid[:, :, 0].any(-1), scores=torch.ones_like(buf.preds[:, :, :, 0, 0]),
                                               trajs=buf.preds[:, :, :, buf.step_future_start:], agent_type=batch["ref/agent_type"])
        out["womd_records_reactive_replay"] = self.womd_metrics_reactive_replay.update(batch, pred_dict["waymo_trajs"], pred_dict["waymo_scores"])
        self.womd_metrics_reactive_replay.aggregate_on_cpu(out["womd_records_reactive_replay"])  # like :655-660
        self.womd_metrics_reactive_replay.reset()
        out["reactive_replay"], out["pred_dict_reactive_replay"] = buf, pred_dict

        # ! joint_future_pred (:683-722)
        buf, goal_sample, goal_log_probs = self.joint_future_pred(batch=batch, input_feature_dict=input_feature_dict, latent=latent_prior,
                                                                  goal=goal_pred, goal_valid=goal_valid, require_vis_dict=False)
        self._emit("err_metrics_joint_future_pred", pred_valid=buf.valid, pred_states=buf.preds, gt_valid=gt_valid, gt_states=gt_states,
                   override_masks=buf.override_masks, agent_role=batch["ref/agent_role"])
        self._emit("rule_metrics_joint_future_pred", valid=buf.valid, override_masks=buf.override_masks, agent_type=batch["ref/agent_type"],
                   **{k: buf.violations[k] for k in ("outside_map", "collided", "run_road_edge", "run_red_light", "passive",
                                                     "goal_reached", "dest_reached")})
        pred_dict = self.waymo_post_processing(valid=buf.valid[:, :, 0].any(-1),
                                               scores=torch.exp(buf.latent_log_probs[..., 0] + goal_log_probs),
                                               trajs=buf.preds[:, :, :, buf.step_future_start:], agent_type=batch["ref/agent_type"])
        out["womd_records_joint_future_pred"] = self.womd_metrics_joint_future_pred.update(batch, pred_dict["waymo_trajs"],
                                                                                           pred_dict["waymo_scores"])
        self.womd_metrics_joint_future_pred.aggregate_on_cpu(out["womd_records_joint_future_pred"])
        self.womd_metrics_joint_future_pred.reset()
        self._emit("sub_womd_joint_future_pred", waymo_trajs=pred_dict["waymo_trajs"], waymo_scores=pred_dict["waymo_scores"],
                   mask_pred=batch["history/agent/role"][..., 2] if "history/agent/role" in batch else None)
        out["joint_future_pred"], out["pred_dict_joint_future_pred"] = buf, pred_dict
        out["goal_sample"], out["goal_log_probs"] = goal_sample, goal_log_probs
        return out

    @torch.no_grad()
    def test_step(self, batch: Dict[str, Tensor], batch_idx: int = 0) -> Dict:
        """waymo_motion.py:902-944: only the history is available; K joint futures from the prior latent and the predicted
        destination, post-processed for the submission writer (attach one through `metric_sinks`)."""
        batch = self.pre_processing(batch)
        input_dict, prior_dict = self._split(batch, "input"), self._split(batch, "latent_prior")
        input_feature_dict = self.model.encode_input_features(**input_dict)
        latent_prior_feature_dict = self.model.encode_input_features(**prior_dict)
        goal_valid = input_dict["agent_valid"].any(1)
        goal_pred = self.model.goal_manager.pred_goal(agent_type=batch["ref/agent_type"], map_type=batch["ref/map_type"],
                                                      agent_state=batch["ref/agent_state"], **input_feature_dict)
        latent_prior = self.model.latent_encoder(**latent_prior_feature_dict)
        for k in ("valid", "vel", "acc", "yaw_rate", "pos", "yaw_bbox", "spd", "size"):
            batch[f"agent/{k}"] = batch[f"history/agent/{k}"]
        buf, goal_sample, goal_log_probs = self.joint_future_pred(batch=batch, input_feature_dict=input_feature_dict, latent=latent_prior,
                                                                  goal=goal_pred, goal_valid=goal_valid, require_vis_dict=False)
        pred_dict = self.waymo_post_processing(valid=buf.valid[:, :, 0].any(-1),
                                               scores=torch.exp(buf.latent_log_probs[..., 0] + goal_log_probs),
                                               trajs=buf.preds[:, :, :, buf.step_future_start:], agent_type=batch["ref/agent_type"])
        self._emit("sub_womd_joint_future_pred", waymo_trajs=pred_dict["waymo_trajs"], waymo_scores=pred_dict["waymo_scores"],
                   mask_pred=batch["history/agent/role"][..., 2] if "history/agent/role" in batch else None,
                   object_id=batch.get("history/agent/object_id"), scenario_center=batch.get("scenario_center"),
                   scenario_yaw=batch.get("scenario_yaw"), scenario_id=batch.get("scenario_id"))
        return {"joint_future_pred": buf, "pred_dict": pred_dict, "goal_sample": goal_sample, "goal_log_probs": goal_log_probs}

    # ------------------------------------------------------------------------------------------------ training
    def train_state(self):
        """flat parameter / gradient / Adam buffers of the training path (`train.trainer.TrainState`), created on first use
        from the current parameters.  From then on every `nn.Parameter` of this module is a VIEW into the flat parameter buffer
        and its `.grad` a view into the flat gradient buffer, so `state_dict()`, checkpoints, torch optimizers and the
        inference engine all see the trained values."""
        if self._train_state is None:
            from ..train.trainer import TrainState
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise nt.TbError("trafficbots_b200.WaymoMotion must live on a CUDA device (no CPU implementation of the training path)")
            hp = self._train_hparams
            ts = TrainState(self.state_dict(), device=dev, lr=hp["lr"], lr_goal=hp["lr_goal"], max_grad_norm=hp["max_grad_norm"],
                            p_rollout_prior=hp["p_training_rollout_prior"], dropout_p=self.train_dropout_p)
            named = dict(self.named_parameters(remove_duplicate=False))  # shared blocks appear under both of their names
            for k, view in ts.params.t.items():
                named[k].data = view
                named[k].requires_grad_(True)
                named[k].grad = ts.params.g[k]
            self._train_state = ts
            self.mark_params_dirty()
        return self._train_state

    def training_step(self, batch: Dict[str, Tensor], batch_idx: int = 0):
        """waymo_motion.py:356-418 on the raw episode batch (`agent/*`, `tl_stop/*`, `map/*`, `agent/dest`, ...: the keys the
        reference's data module delivers; the re-keying of `pre_processing` is folded into `train/graph.py`).  Runs the forward
        AND the backward of the step (the gradients are in `p.grad` of every parameter when it returns, like after Lightning's
        `loss.backward()`), and -- with `self.automatic_optimization` left True -- also what the Lightning trainer does next:
        gradient averaging over the data-parallel ranks (one NCCL all-reduce of the flat buffer), `clip_grad_norm_` at
        `gradient_clip_val` (configs/trainer/default.yaml:12) and the Adam step (:955-973).  Dropout is not implemented: the
        step is the reference's with every dropout probability at 0.  Returns the loss (device scalar); the terms of
        `TrainingMetrics.compute` go to `self.log` as `training/*`."""
        ts = self.train_state()
        if ts._static_batch is not None and all(tuple(batch[k].shape) == tuple(v.shape) for k, v in ts._static_batch.items()):
            out = ts.replay(batch)  # `train_state().capture(batch)` was called for this batch shape: whole-step CUDA graph
        else:
            out = ts.forward_backward({k: v.to(ts.device, non_blocking=True) for k, v in batch.items()})
        if self.automatic_optimization:
            ts.all_reduce_grads()
            out["grad_sq_norm"] = ts.optimizer_step()
            self.mark_params_dirty()  # the inference engine re-packs its weight blob on its next use
        for k in ("loss", "vae_kl", "diffbar_reward", "goal_loss"):
            self.log(f"training/{k}", out[k], on_step=True)
        return out["loss"]

    def configure_optimizers(self):
        """waymo_motion.py:955-973: Adam with a second parameter group (`lr_goal`) for the goal predictor and StepLR per epoch.
        Returns handles on the fused flat-buffer optimizer (the step itself runs inside `training_step`)."""
        ts = self.train_state()
        hp = self._train_hparams

        class _FlatAdam:
            param_groups = [{"lr": hp["lr"], "name": "model"}, {"lr": hp["lr_goal"], "name": "goal_predictor"}]

            def step(self_inner):
                ts.optimizer_step()
                self.mark_params_dirty()

            def zero_grad(self_inner, set_to_none: bool = False):
                ts.flat_g.zero_()

        class _StepLR:
            def __init__(self_inner):
                self_inner.epoch = 0

            def step(self_inner):
                self_inner.epoch += 1
                if self_inner.epoch % hp["lr_step_size"] == 0:
                    ts.ops.scale_(ts.lr, hp["lr_gamma"])

        return [_FlatAdam()], [{"scheduler": _StepLR(), "monitor": "val/loss", "interval": "epoch", "frequency": 1, "strict": True}]

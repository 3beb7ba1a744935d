"""`TrafficBots` -- mirror of the reference world model's interface (`src/models/traffic_bots.py:19-247`) on top of the
CUDA library.  Parameters are registered under the reference's `state_dict` keys (`model.*`), so a reference
checkpoint loads with `load_state_dict` unchanged; they are re-laid into the kernel layout lazily
(`tb_pack_weights`) whenever they change.  The per-step `forward` of the reference is fused into the rollout kernels
(`tb_step_front` / `tb_step_back`); what remains here is what the outer loop calls directly:
`encode_input_features`, `init`, the `hidden` / `latent_sample` attributes and `goal_manager` / `latent_encoder` handles.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Mapping, Optional, Union

import torch
from torch import Tensor, nn

from .. import _native as nt
from .. import config as tb_config
from .. import weights
from ..engine import Engine, SceneFeatures
from .distributions import DestCategorical, DiagGaussian


def register_param_tree(root: nn.Module, spec: Mapping[str, tuple], prefix: str, buffers: bool = False) -> None:
    """creates nested container modules so that `root.state_dict()` has exactly the keys of `spec` below `prefix`."""
    for key, shape in spec.items():
        if not key.startswith(prefix):
            continue
        parts = key[len(prefix):].split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        t = torch.zeros(shape)
        if buffers:
            mod.register_buffer(parts[-1], t)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))


class _Head(nn.Module):
    """parameter container of a pre-rollout head that calls back into the owning `TrafficBots` (not a submodule link)."""

    def _model(self) -> "TrafficBots":
        return self.__dict__["_tb"]()


class LatentEncoder(_Head):
    def forward(self, agent_feature: Tensor = None, agent_feature_valid: Tensor = None, map_feature: Tensor = None,
                map_feature_valid: Tensor = None, tl_feature: Tensor = None, tl_feature_valid: Tensor = None,
                posterior: bool = False, **private) -> DiagGaussian:
        """`LatentEncoder.forward` (src/models/latent_encoder.py:70-147): called like the reference with the dict of
        `encode_input_features` (`model.latent_encoder(posterior=True, **feat)`), whose private `_kv_*` entries carry the
        projected K|V caches.  Returns the diagonal Gaussian with the learned constant log_std (:195-199)."""
        if "_kv_map" not in private or "_kv_tl" not in private:
            raise nt.TbError("latent_encoder expects the feature dict returned by encode_input_features (with its K|V caches)")
        feat = dict(private, agent_feature=agent_feature, agent_feature_valid=agent_feature_valid, map_feature=map_feature,
                    map_feature_valid=map_feature_valid, tl_feature=tl_feature, tl_feature_valid=tl_feature_valid)
        mean, valid = self._model()._engine().latent_encoder(feat, posterior=posterior)
        dist = self.latent_post_dist if posterior else self.latent_prior_dist
        return DiagGaussian(mean, dist.log_std.detach(), valid=valid)


class GoalManager(_Head):
    @torch.no_grad()
    def get_gt_goal(self, agent_valid: Tensor, gt_goal: Tensor, gt_dest: Tensor):
        """goal_manager.py:49-74, goal_attr_mode "dest": the GT destination index, valid for agents seen in the history."""
        return gt_dest, agent_valid.any(1)

    def pred_goal(self, agent_type: Tensor, map_type: Tensor, agent_state: Tensor = None, **feat) -> DestCategorical:
        """`GoalManager.pred_goal` -> `DestPredictor.forward`, mode mlp (src/models/goal_manager.py:77-81,202-333)."""
        probs, _logp, valid = self._model()._engine().dest_predictor(feat, agent_type, map_type)
        return DestCategorical(probs=probs, valid=valid)


class TrafficBots(nn.Module):
    def __init__(self, hidden_dim: int = 128, **cfg) -> None:
        super().__init__()
        tb_config.check_supported({"hidden_dim": hidden_dim, "model": dict(cfg, hidden_dim=hidden_dim)})
        self.hidden_dim = hidden_dim
        import weakref
        self.latent_encoder = LatentEncoder()
        self.goal_manager = GoalManager()
        for head in (self.latent_encoder, self.goal_manager):
            head.__dict__["_tb"] = weakref.ref(self)
        spec = weights.state_dict_spec()
        register_param_tree(self, {k: v for k, v in spec.items() if weights._alias_of(k) is None}, "model.")
        # shared modules: the latent encoder re-exports the policy's cross-attention blocks (latent_encoder.py:39-41)
        self.latent_encoder.add_module("transformer_as2pl", self.transformer_as2pl)
        self.latent_encoder.add_module("transformer_as2tl", self.transformer_as2tl)
        gm = cfg.get("goal_manager", {}) or {}
        self.goal_manager.dummy = False
        self.goal_manager.update_goal = False
        self.goal_manager.goal_attr_mode = gm.get("goal_attr_mode", "dest")
        self.latent_sample: Optional[Tensor] = None
        self.latent_logp: Optional[Tensor] = None
        self.__dict__["_owner"] = None  # weak handle on the WaymoMotion shell that holds the Engine (not a submodule)

    # the Engine (packed parameters + device buffers) lives on the LightningModule shell
    def _engine(self) -> Engine:
        owner = self.__dict__["_owner"]() if self.__dict__["_owner"] is not None else None
        if owner is None:
            raise nt.TbError("TrafficBots must be owned by a trafficbots_b200 WaymoMotion module")
        return owner.engine()

    def set_owner(self, owner) -> None:
        import weakref
        self.__dict__["_owner"] = weakref.ref(owner)

    def encode_input_features(self, batch: Mapping[str, Tensor] = None, prefix: str = "history/", raw=None, **kw) -> SceneFeatures:
        """SceneCentricInput + encode_input_features (sc_input.py:98-140, traffic_bots.py:109-151) fused: consumes the
        RAW scene tensors; positional encodings are computed inside the kernels instead of being materialised
        (`input/map_pe` alone is 252 MB at 32 scenes in the reference).  Two call forms:
          * the reference's, `encode_input_features(**input_dict)` with the dict built from the `input/` / `latent_prior/` /
            `latent_post/` keys of `pre_processing(batch)` (waymo_motion.py:577-583): the `raw` entry is the `RawScene` handle
            of that group.  A handle that was already encoded returns its cached features (`latent_prior` aliases `input` in
            eval mode), and a handle with `map_from` re-uses that group's map features (the posterior pass);
          * `encode_input_features(batch, prefix=...)` on a raw batch dict (`map/*`, `{prefix}agent/*`, `{prefix}tl_stop/*`).
        Returns the reference's feature dict (+ private K|V caches)."""
        if raw is not None:
            if raw.features is None:
                share = None
                if raw.map_from is not None:
                    if raw.map_from.features is None:
                        raw.map_from.features = self._engine().encode_scene(raw.map_from.batch, raw.map_from.prefix)
                    share = raw.map_from.features
                raw.features = self._engine().encode_scene(raw.batch, raw.prefix, share_map=share)
            return raw.features
        if batch is None:
            if "agent_attr" in kw or "map_pe" in kw:
                raise nt.TbError("encode_input_features: got materialised attr / pe tensors; use trafficbots_b200's pre_processing "
                                 "modules (data_modules/scene_centric.py), which pass the raw tensors through a `raw` handle")
            batch = kw
        return self._engine().encode_scene(batch, prefix)

    def init(self, latent: Union[DiagGaussian, Tensor], deterministic: Union[bool, Tensor]) -> None:
        """traffic_bots.py:153-161,196-199 -- the latent is sampled once per rollout; GRU hidden restarts at zero."""
        if isinstance(latent, Tensor):
            self.latent_sample = latent
            self.latent_logp = torch.zeros(latent.shape[:-1], device=latent.device)
        else:
            self.latent_sample = latent.sample(deterministic).contiguous()
            self.latent_logp = latent.log_prob(self.latent_sample).contiguous()

    @property
    def hidden(self) -> Optional[Tensor]:
        eng = self._engine()
        return None if eng._state is None else eng.state_field(nt.STATE_HIDDEN)

    def forward(self, *a, **kw):
        raise nt.TbError("TrafficBots.forward is fused into the rollout kernels: call WaymoMotion.forward / rollout")

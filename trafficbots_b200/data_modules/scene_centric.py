"""Mirrors of the reference's batch-assembly modules with the same class names, constructor keywords and
batch keys, so that `batch = self.pre_processing(batch)` followed by the reference's three
`self.model.encode_input_features(**{input, latent_post, latent_prior}_dict)` calls (waymo_motion.py:576-583, :904-908)
works unchanged on the fused path:

  SceneCentricPreProcessing  data_modules/scene_centric.py:8-135   re-keying `sc/*`, `gt/*`, `ref/*` (views, no copies)
  SceneCentricInput          data_modules/sc_input.py:8-140        `input/*`
  SceneCentricLatent         data_modules/sc_latent.py:10-241      `latent_prior/*`, `latent_post/*`

What is deliberately different: the reference materialises `input/{agent,map,tl}_attr` and `_pe` (333 MB of fp32 at 32 scenes,
of which `map_pe` [S,P,20,96] alone is 252 MB) and its encoders read them back; here the encoders compute attributes and
positional encodings in-kernel from the raw tensors, so these modules emit the cheap keys (`*_valid`, `*_pos`) and ONE extra
entry per group, `<group>/raw`: a `RawScene` handle naming the raw tensors of that group.  `encode_input_features(**dict)`
takes the handle.  In eval mode `latent_prior/*` aliases `input/*` (sc_latent.py:96-97,120-121,166-167), so both carry the
SAME handle and the second call returns the first call's features instead of encoding the map again; `latent_post/raw` is
the full-episode view and shares the map features of `input/raw` (the reference encodes the identical map three times).
Training-time augmentations (`dropout_p_history`, `perturb_input_to_latent`) are not implemented and are rejected.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch
from torch import Tensor, nn

from ..config import UnsupportedConfig


class RawScene:
    """handle on the raw tensors of one encoder input group: `batch` + key prefix of the agent / traffic-light tensors
    ("sc/" = the re-keyed history tensors `sc/agent_*`, `sc/tl_*`; "" = the whole episode `agent/*`, `tl_stop/*`); `features` caches the encoder output; `map_from` names the
    group whose map features are re-used."""

    def __init__(self, batch: Mapping[str, Tensor], prefix: str, map_from: Optional["RawScene"] = None) -> None:
        self.batch, self.prefix, self.map_from = batch, prefix, map_from
        self.features = None

    def __repr__(self) -> str:
        return f"RawScene(prefix={self.prefix!r}, encoded={self.features is not None})"


class SceneCentricPreProcessing(nn.Module):
    def __init__(self, time_step_current: int = 10, data_size: Optional[Mapping] = None) -> None:
        super().__init__()
        self.n_step_hist = time_step_current + 1
        self.model_kwargs: Dict = {}

    def forward(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        prefix = "" if self.training else "history/"
        h = self.n_step_hist
        for k in ("valid", "pos", "z", "vel", "spd", "acc", "yaw_bbox", "yaw_rate"):
            if f"{prefix}agent/{k}" in batch:
                batch[f"sc/agent_{k}"] = batch[f"{prefix}agent/{k}"][:, :h].contiguous()
        for k in ("type", "role", "size"):
            if f"{prefix}agent/{k}" in batch:
                batch[f"sc/agent_{k}"] = batch[f"{prefix}agent/{k}"]
        if "agent/valid" in batch:  # training / validation: ground truth for losses and metrics
            for k in ("cmd", "goal", "dest"):
                if f"agent/{k}" in batch:
                    batch[f"gt/{k}"] = batch[f"agent/{k}"]
            for k in ("valid", "spd", "pos", "vel", "yaw_bbox"):
                batch[f"gt/{k}"] = batch[f"agent/{k}"]
            batch["gt/state"] = torch.cat([batch["gt/pos"], batch["gt/yaw_bbox"], batch["gt/spd"]], dim=-1)
        for k in ("valid", "type", "pos", "dir"):
            batch[f"sc/map_{k}"] = batch[f"map/{k}"]
        for k in ("valid", "state", "pos", "dir"):
            batch[f"sc/tl_{k}"] = batch[f"{prefix}tl_stop/{k}"][:, :h].contiguous()
        if not self.training:
            for k in ("valid", "pos", "z", "vel", "spd", "yaw_bbox"):
                if f"history/agent_no_sim/{k}" in batch:
                    batch[f"sc/agent_no_sim_{k}"] = batch[f"history/agent_no_sim/{k}"][:, :h].contiguous()
            for k in ("type", "size"):
                if f"history/agent_no_sim/{k}" in batch:
                    batch[f"sc/agent_no_sim_{k}"] = batch[f"history/agent_no_sim/{k}"]
        batch["ref/agent_type"] = batch[prefix + "agent/type"]
        batch["ref/agent_role"] = batch[prefix + "agent/role"] if prefix + "agent/role" in batch else None
        batch["ref/map_type"] = batch["map/type"]
        batch["ref/agent_state"] = torch.cat([batch["sc/agent_pos"], batch["sc/agent_yaw_bbox"], batch["sc/agent_spd"]], dim=-1)
        return batch


class _WithPeBuffers(nn.Module):
    """registers the reference's buffers under the reference's names (`pl_node_ohe`, `pose_pe_{agent,map,tl}.pe_{xy,yaw}.freqs`)
    so that `state_dict` keys match; the kernels read the packed copies."""

    def __init__(self, which: str) -> None:
        super().__init__()
        from .. import weights
        from ..models.traffic_bots import register_param_tree
        register_param_tree(self, weights.state_dict_spec(), f"pre_processing.{which}.", buffers=True)
        self.pl_node_ohe.copy_(torch.eye(weights.N_PL_NODE))
        for who in ("agent", "map", "tl"):
            pe = getattr(self, f"pose_pe_{who}")
            pe.pe_xy.freqs.copy_(weights.pe_freqs_xy())
            pe.pe_yaw.freqs.copy_(weights.pe_freqs_yaw())


class SceneCentricInput(_WithPeBuffers):
    def __init__(self, time_step_current: int = 10, data_size: Optional[Mapping] = None, dropout_p_history: float = -1,
                 pe_dim: int = 96, pose_pe: Optional[Mapping] = None) -> None:
        super().__init__("input")
        if 0 < dropout_p_history <= 1.0:
            raise UnsupportedConfig("pre_processing.input.dropout_p_history (training-time history dropout) is not implemented")
        self.n_step_hist = time_step_current + 1
        self.model_kwargs = {"n_step_hist": self.n_step_hist, "n_pl_node": 20}

    def forward(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        batch["input/agent_valid"] = batch["sc/agent_valid"]
        batch["input/tl_valid"] = batch["sc/tl_valid"]
        batch["input/map_valid"] = batch["sc/map_valid"]
        batch["input/agent_pos"] = batch["sc/agent_pos"]
        batch["input/map_pos"] = batch["sc/map_pos"][:, :, 0]
        batch["input/tl_pos"] = batch["sc/tl_pos"]
        batch["input/raw"] = RawScene(batch, "sc/")
        return batch


class SceneCentricLatent(_WithPeBuffers):
    def __init__(self, time_step_current: int = 10, data_size: Optional[Mapping] = None, dropout_p_history: float = -1,
                 pe_dim: int = 96, pose_pe: Optional[Mapping] = None, perturb_input_to_latent: bool = False,
                 max_meter: float = 50.0, max_rad: float = 3.14) -> None:
        super().__init__("latent")
        if perturb_input_to_latent or 0 < dropout_p_history <= 1.0:
            raise UnsupportedConfig("pre_processing.latent: perturb_input_to_latent / dropout_p_history are not implemented")
        self.model_kwargs: Dict = {}

    def forward(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        for who in ("agent", "map", "tl"):
            for k in ("valid", "pos"):
                batch[f"latent_prior/{who}_{k}"] = batch[f"input/{who}_{k}"]
        batch["latent_prior/raw"] = batch["input/raw"]  # the same inputs: encoded once
        if "agent/valid" in batch:  # training / validation: posterior over the whole episode
            for k in ("valid", "pos"):
                batch[f"latent_post/map_{k}"] = batch[f"input/map_{k}"]
            batch["latent_post/tl_valid"] = batch["tl_stop/valid"]
            batch["latent_post/tl_pos"] = batch["tl_stop/pos"]
            batch["latent_post/agent_valid"] = batch["agent/valid"]
            batch["latent_post/agent_pos"] = batch["agent/pos"]
            batch["latent_post/raw"] = RawScene(batch, "", map_from=batch["input/raw"])
        return batch

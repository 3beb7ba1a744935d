"""Parameter schema of the hot path = the reference `WaymoMotion.state_dict()` (483 tensors).

`state_dict_spec()` lists every key with its shape exactly as the reference registers them
(`src/pl_modules/waymo_motion.py:66-77`, `src/models/traffic_bots.py:63-107`, `src/models/modules/*.py`), so a
reference checkpoint loads unchanged.  `init_state_dict(seed)` fills the schema deterministically (default-init
scales of torch's Linear / GRU / xavier attention, with LayerNorm affine and biases perturbed so every term of the
arithmetic is exercised) -- this is what tests, golden vectors and `bench.py` use, since no trained checkpoint is
distributable (reference `README.md:38`).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

D = 128  # hidden_dim   (configs/model/traffic_bots.yaml:9)
N_HEAD = 4  # tf_cfg.n_head (:43)
D_FF = 128  # tf_cfg.d_feedforward (:48)
PE_DIM = 96  # pre_processing.input.pe_dim (:20)
N_GRU_LAYER = 3  # agent_temporal.num_layers (:91)
LATENT_DIM = 16  # latent_encoder.latent_dim (:73)
N_PL_NODE = 20
N_PL_TYPE = 11
N_TL_STATE = 5
AGENT_ATTR_DIM = 11
MAP_ATTR_DIM = N_PL_TYPE + N_PL_NODE  # 31


def pe_freqs_xy(dim: int = PE_DIM // 4, theta: float = 1e3) -> torch.Tensor:
    """`PositionalEmbedding.freqs` (src/utils/pos_emb.py:11-13): 1/theta^(2i/dim), each repeated twice."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    return freqs.repeat_interleave(2, 0)


def pe_freqs_yaw(dim: int = PE_DIM // 2) -> torch.Tensor:
    """`PositionalEmbeddingRad.freqs` (src/utils/pos_emb.py:42-43): 1,1,2,2,...,dim/2,dim/2."""
    freqs = torch.arange(0, dim // 2) + 1.0
    return freqs.repeat_interleave(2, 0)


def _tf_layer(prefix: str, spec: "OrderedDict[str, Tuple[int, ...]]") -> None:
    # registration order of TransformerCrossAttention.__init__ (src/models/modules/transformer.py:113-133)
    spec[f"{prefix}.norm1.weight"] = (D,)
    spec[f"{prefix}.norm1.bias"] = (D,)
    spec[f"{prefix}.norm_tgt.weight"] = (D,)
    spec[f"{prefix}.norm_tgt.bias"] = (D,)
    spec[f"{prefix}.attn.in_proj_weight"] = (3 * D, D)
    spec[f"{prefix}.attn.out_proj_weight"] = (D, D)
    spec[f"{prefix}.attn.in_proj_bias"] = (3 * D,)
    spec[f"{prefix}.attn.out_proj_bias"] = (D,)
    spec[f"{prefix}.linear1.weight"] = (D_FF, D)
    spec[f"{prefix}.linear1.bias"] = (D_FF,)
    spec[f"{prefix}.linear2.weight"] = (D, D_FF)
    spec[f"{prefix}.linear2.bias"] = (D,)
    spec[f"{prefix}.norm2.weight"] = (D,)
    spec[f"{prefix}.norm2.bias"] = (D,)


def _tf_block(prefix: str, n_layer: int, spec) -> None:
    for i in range(n_layer):
        _tf_layer(f"{prefix}.layers.{i}", spec)


def _gru(prefix: str, spec) -> None:
    for i in range(N_GRU_LAYER):
        spec[f"{prefix}.rnn.weight_ih_l{i}"] = (3 * D, D)
        spec[f"{prefix}.rnn.weight_hh_l{i}"] = (3 * D, D)
        spec[f"{prefix}.rnn.bias_ih_l{i}"] = (3 * D,)
        spec[f"{prefix}.rnn.bias_hh_l{i}"] = (3 * D,)


def _mlp(prefix: str, dims, spec, use_layernorm: bool, dropout: bool, end_layer_activation: bool) -> None:
    """index bookkeeping of `MLP.__init__` (src/models/modules/mlp.py:36-64)."""
    idx = 0
    n = len(dims) - 1
    for i in range(n):
        spec[f"{prefix}.fc_layers.{idx}.weight"] = (dims[i + 1], dims[i])
        spec[f"{prefix}.fc_layers.{idx}.bias"] = (dims[i + 1],)
        idx += 1
        last = i == n - 1
        if (not last) or end_layer_activation:
            if use_layernorm:
                spec[f"{prefix}.fc_layers.{idx}.weight"] = (dims[i + 1],)
                spec[f"{prefix}.fc_layers.{idx}.bias"] = (dims[i + 1],)
                idx += 1
            if dropout:
                idx += 1
        if not last:
            idx += 1  # activation module


def state_dict_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for pp in ("input", "latent"):
        spec[f"pre_processing.{pp}.pl_node_ohe"] = (N_PL_NODE, N_PL_NODE)
        for who in ("agent", "map", "tl"):
            spec[f"pre_processing.{pp}.pose_pe_{who}.pe_xy.freqs"] = (PE_DIM // 4,)
            spec[f"pre_processing.{pp}.pose_pe_{who}.pe_yaw.freqs"] = (PE_DIM // 2,)
    m = "model"
    _mlp(f"{m}.map_encoder.input_pe_encoder.mlp", [MAP_ATTR_DIM, D - PE_DIM, D - PE_DIM], spec, False, True, False)
    _tf_block(f"{m}.map_encoder.transformer_densetnt", 3, spec)
    _tf_block(f"{m}.map_encoder.transformer_self_attn", 1, spec)
    _mlp(f"{m}.tl_encoder.mlp", [N_TL_STATE, D - PE_DIM, D - PE_DIM], spec, False, True, False)
    _mlp(f"{m}.agent_encoder.mlp", [AGENT_ATTR_DIM, D - PE_DIM, D - PE_DIM], spec, False, True, False)
    _tf_block(f"{m}.transformer_as2pl", 3, spec)
    _tf_block(f"{m}.transformer_as2tl", 3, spec)
    _gru(f"{m}.goal_manager.goal_predictor.gru_as", spec)
    _mlp(f"{m}.goal_manager.goal_predictor.mlp", [2 * D, D, D, 1], spec, True, False, False)
    # latent encoder: the two shared transformers appear again under its prefix (aliases, latent_encoder.py:39-41)
    _tf_block(f"{m}.latent_encoder.transformer_as2pl", 3, spec)
    _tf_block(f"{m}.latent_encoder.transformer_as2tl", 3, spec)
    for pp in ("prior", "post"):
        spec[f"{m}.latent_encoder.latent_{pp}_dist.log_std"] = (LATENT_DIM,)
        _mlp(f"{m}.latent_encoder.latent_{pp}_dist.mlp_mean", [D, D, LATENT_DIM], spec, False, False, False)
    _gru(f"{m}.latent_encoder.agent_temporal_post", spec)
    _tf_block(f"{m}.latent_encoder.agent_interaction_post.transformer", 3, spec)
    _gru(f"{m}.latent_encoder.agent_temporal_prior", spec)
    _tf_block(f"{m}.latent_encoder.agent_interaction_prior.transformer", 3, spec)
    _gru(f"{m}.agent_temporal", spec)
    _tf_block(f"{m}.agent_interaction.transformer", 3, spec)
    _mlp(f"{m}.add_goal.mlp_in", [D, D, D, D], spec, True, True, True)
    _mlp(f"{m}.add_goal.mlp_out", [2 * D, D, D], spec, False, True, True)
    _mlp(f"{m}.add_latent.mlp_in", [LATENT_DIM, D, D], spec, False, True, True)
    _mlp(f"{m}.add_latent.mlp_out", [2 * D, D, D], spec, False, True, True)
    for c in range(3):
        _mlp(f"action_head.mlp_mean.{c}", [D, D, 2], spec, False, False, False)
    for c in range(3):
        spec[f"action_head.log_std.{c}"] = (2,)
    return spec


ALIASES = {  # shared modules: the latent encoder re-exports the policy's cross-attention blocks
    "model.latent_encoder.transformer_as2pl": "model.transformer_as2pl",
    "model.latent_encoder.transformer_as2tl": "model.transformer_as2tl",
}


def _alias_of(key: str):
    for a, b in ALIASES.items():
        if key.startswith(a + "."):
            return b + key[len(a):]
    return None


def init_state_dict(seed: int = 2023, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic parameters for every key of `state_dict_spec()` (CPU tensors).

    Same seed -> bit-identical tensors on any box with this torch build (CPU philox-free mt19937 generator),
    which is what lets golden vectors generated in the build container be replayed on the GPU box.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    spec = state_dict_spec()
    for k, shape in spec.items():
        src = _alias_of(k)
        if src is not None:
            sd[k] = sd[src]
            continue
        if k.endswith("pl_node_ohe"):
            v = torch.eye(N_PL_NODE)
        elif k.endswith("pe_xy.freqs"):
            v = pe_freqs_xy()
        elif k.endswith("pe_yaw.freqs"):
            v = pe_freqs_yaw()
        elif ".log_std" in k:
            v = torch.full(shape, -2.0 if k.startswith("action_head") else -1.0) + uni(shape, 0.05)
        elif "norm" in k or (len(shape) == 1 and k.endswith(".weight")):  # LayerNorm affine
            v = 1.0 + uni(shape, 0.1) if k.endswith("weight") else uni(shape, 0.1)
        elif k.endswith("in_proj_weight") or k.endswith("out_proj_weight"):
            v = uni(shape, math.sqrt(6.0 / (shape[0] + shape[1])))
        elif k.endswith("in_proj_bias") or k.endswith("out_proj_bias"):
            v = uni(shape, 0.02)
        elif ".rnn." in k:
            v = uni(shape, 1.0 / math.sqrt(D))
        elif k.endswith(".weight"):
            v = uni(shape, 1.0 / math.sqrt(shape[1]))
        elif k.endswith(".bias"):
            fan_in = spec[k[: -len("bias")] + "weight"][-1]
            v = uni(shape, 1.0 / math.sqrt(fan_in))
        else:
            raise KeyError(k)
        assert tuple(v.shape) == tuple(shape), (k, v.shape, shape)
        sd[k] = v.to(dtype).contiguous()
    return sd


def count_parameters(sd: Dict[str, torch.Tensor]) -> int:
    seen = set()
    n = 0
    for k, v in sd.items():
        if k.startswith("pre_processing."):
            continue
        if v.data_ptr() in seen:
            continue
        seen.add(v.data_ptr())
        n += v.numel()
    return n

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TrafficBots hot path (fp32, torch CPU ops).

A functional restatement of the reference's algorithm for the path named in BASELINE.json `north_star`:
scene encoding (map / agent / traffic-light encoders), prior latent, destination prediction and the
closed-loop multi-agent rollout.  Every function cites the reference `file:line` it follows (paths relative
to the reference's `src/`).  It is written "as implemented" by the reference (K/V are re-projected every step,
loop-invariant MLPs are re-evaluated every step) so that, timed on the host cores, it stands in for the
reference's own CPU path (`bench.py --impl reference`, `cpu_baseline.kind == "port"`).

PARITY PINNING: pinned against the UNMODIFIED reference executed in the build container
(`tests/test_oracle_vs_reference.py`, via `oracle/ref_loader.py`) and against the committed golden vectors
`tests/golden/*.npz` that `oracle/make_golden.py` produced from the reference (`tests/test_oracle_golden.py`).
The reference itself ships no tests / golden vectors.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference` legs may import
this module.  Nothing under `trafficbots_b200/` does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]
D = 128
N_HEAD = 4
DT = 0.1
MAX_ACC = (5.0, 7.0, 6.0)  # veh, ped, cyc  (configs/model/traffic_bots.yaml:142-155; order utils/dynamics.py:23-27)
MAX_YAW_RATE = (1.5, 7.0, 3.0)


# ----------------------------------------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------------------------------------
def pose_pe(xy: Tensor, yaw: Tensor, f_xy: Tensor, f_yaw: Tensor) -> Tensor:
    """`PosePE.forward` mode pe_xy_yaw (utils/pose_pe.py:57-62) with `PositionalEmbedding(.Rad).forward`
    (utils/pos_emb.py:23-26, 53-56): [cos(x f_even), sin(x f_odd), cos(y ..), sin(y ..), cos(k yaw), sin(k yaw)]."""

    def emb(v: Tensor, freqs: Tensor) -> Tensor:
        e = v.unsqueeze(-1) * freqs.view([1] * v.dim() + [-1])
        return torch.cat([torch.cos(e[..., ::2]), torch.sin(e[..., 1::2])], dim=-1)

    return torch.cat([emb(xy[..., 0], f_xy), emb(xy[..., 1], f_xy), emb(yaw, f_yaw)], dim=-1)


def layer_norm(x: Tensor, sd: SD, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def linear(x: Tensor, sd: SD, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def input_pe_encoder(sd: SD, prefix: str, valid: Tensor, attr: Tensor, pe: Tensor) -> Tensor:
    """`InputPeEncoder.forward`, pe_mode "cat" (models/modules/input_pe_encoder.py:52-59); its MLP is
    Linear-Dropout-ReLU-Linear without final activation (models/modules/mlp.py:36-64, eval: dropout = id)."""
    x = linear(torch.relu(linear(attr, sd, prefix + ".mlp.fc_layers.0")), sd, prefix + ".mlp.fc_layers.3")
    x = torch.cat([x, pe], dim=-1)
    return x.masked_fill(~valid.unsqueeze(-1), 0.0)


def attention(sd: SD, prefix: str, src: Tensor, tgt: Tensor, tgt_invalid: Tensor,
              attn_mask: Optional[Tensor]) -> Tensor:
    """`Attention.forward` (models/modules/attention.py:79-146), cross-attention branch, no dropout.
    src [B,S,D], tgt [B,T,D], tgt_invalid [B,T] bool, attn_mask [S,T] or [B,S,T] bool (True = disabled)."""
    B, S, _ = src.shape
    T = tgt.shape[1]
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(src, w[:D], b[:D])  # :85
    kv = F.linear(tgt, w[D:], b[D:])  # :86  (recomputed at every call by the reference)
    k, v = kv.chunk(2, dim=-1)
    invalid = tgt_invalid.unsqueeze(1).expand(-1, S, -1)  # :91-94
    if attn_mask is not None:
        invalid = invalid | attn_mask  # :95-99
    dead = invalid.all(-1)  # :103  rows without any valid key
    invalid = invalid & ~dead.unsqueeze(-1)  # :105  (un-mask them so softmax stays finite)
    dh = D // N_HEAD
    q = q.view(B, S, N_HEAD, dh).transpose(1, 2)
    k = k.view(B, T, N_HEAD, dh).transpose(1, 2)
    v = v.view(B, T, N_HEAD, dh).transpose(1, 2)
    logits = torch.matmul(q, k.transpose(-2, -1))  # :115
    logits = logits.masked_fill(invalid.unsqueeze(1), float("-inf"))  # :128
    p = torch.softmax(logits / math.sqrt(dh), dim=-1)  # :130
    o = torch.matmul(p, v).transpose(1, 2).flatten(2, 3)  # :136-141
    o = F.linear(o, sd[prefix + ".out_proj_weight"], sd[prefix + ".out_proj_bias"])  # :142
    return o.masked_fill(dead.unsqueeze(-1), 0.0)  # :144-146


def xlayer(sd: SD, prefix: str, src: Tensor, src_invalid: Tensor, tgt: Tensor, tgt_invalid: Tensor,
           attn_mask: Optional[Tensor] = None) -> Tensor:
    """`TransformerCrossAttention.forward`, norm_first, d_feedforward>0, no decoder self-attention
    (models/modules/transformer.py:186-237)."""
    s2 = layer_norm(src, sd, prefix + ".norm1")  # :190
    t2 = layer_norm(tgt, sd, prefix + ".norm_tgt")  # :192
    s2 = attention(sd, prefix + ".attn", s2, t2, tgt_invalid, attn_mask)  # :197
    src = src + s2  # :203
    s2 = layer_norm(src, sd, prefix + ".norm2")  # :208
    s2 = linear(torch.relu(linear(s2, sd, prefix + ".linear1")), sd, prefix + ".linear2")  # :213-217
    src = src + s2  # :220
    return src.masked_fill(src_invalid.unsqueeze(-1), 0.0)  # :236-237


def tf_block(sd: SD, prefix: str, n_layer: int, src: Tensor, src_invalid: Tensor, tgt: Tensor, tgt_invalid: Tensor,
             attn_mask: Optional[Tensor] = None) -> Tensor:
    """`TransformerBlock.forward` (models/modules/transformer.py:82-95): every layer sees the SAME tgt."""
    for i in range(n_layer):
        src = xlayer(sd, f"{prefix}.layers.{i}", src, src_invalid, tgt, tgt_invalid, attn_mask)
    return src


def interaction(sd: SD, prefix: str, x: Tensor, valid: Tensor) -> Tensor:
    """`MultiAgentTF.forward` (models/modules/agent_interaction.py:51-93): tgt = the block input, eye mask,
    scenes with exactly one valid agent are passed through untouched.  x [B,A,D], valid [B,A]."""
    A = valid.shape[-1]
    eye = torch.eye(A, dtype=torch.bool)
    y = tf_block(sd, prefix + ".transformer", 3, x, ~valid, x, ~valid, eye)
    single = valid.sum(-1) == 1  # :61
    return torch.where(single.view(-1, 1, 1), x, y)


def gru_cell(sd: SD, prefix: str, layer: int, x: Tensor, h: Tensor) -> Tensor:
    """one `nn.GRU` layer, one time step (torch GRU equations; gate order r,z,n in the stacked weights)."""
    gi = F.linear(x, sd[f"{prefix}.rnn.weight_ih_l{layer}"], sd[f"{prefix}.rnn.bias_ih_l{layer}"])
    gh = F.linear(h, sd[f"{prefix}.rnn.weight_hh_l{layer}"], sd[f"{prefix}.rnn.bias_hh_l{layer}"])
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def gru_step(sd: SD, prefix: str, x: Tensor, valid: Tensor, h: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """`MultiAgentGRULoop.forward`, 2-D valid branch (models/modules/agent_temporal.py:147-153).
    x [B,A,D], valid [B,A], h [3,B*A,D] or None -> (x_1 [B,A,D], h_1 [3,B*A,D])."""
    B, A, _ = x.shape
    if h is None:
        h = torch.zeros(3, B * A, D)
    inp = x.flatten(0, 1)
    hs = []
    for layer in range(3):
        inp = gru_cell(sd, prefix, layer, inp, h[layer])
        hs.append(inp)
    invalid = ~valid.flatten(0, 1).unsqueeze(-1)
    h1 = torch.stack(hs, 0).masked_fill(invalid.unsqueeze(0), 0.0)
    return inp.masked_fill(invalid, 0.0).view(B, A, D), h1


def gru_sequence(sd: SD, prefix: str, x: Tensor, valid: Tensor) -> Tensor:
    """`MultiAgentGRULoop.forward`, 3-D valid branch (models/modules/agent_temporal.py:133-146).
    x [B,T,A,D], valid [B,T,A] -> [B,T,A,D]; the hidden state of invalid agents is zeroed after every step,
    the outputs are zeroed where invalid."""
    B, T, A, _ = x.shape
    h = torch.zeros(3, B * A, D)
    outs = []
    for t in range(T):
        inp = x[:, t].flatten(0, 1)
        hs = []
        for layer in range(3):
            inp = gru_cell(sd, prefix, layer, inp, h[layer])
            hs.append(inp)
        invalid = ~valid[:, t].flatten(0, 1).unsqueeze(-1)
        h = torch.stack(hs, 0).masked_fill(invalid.unsqueeze(0), 0.0)
        outs.append(inp.masked_fill(invalid, 0.0).view(B, A, D))
    return torch.stack(outs, dim=1)


def goal_mlp_in(sd: SD, z: Tensor) -> Tensor:
    """`add_goal.mlp_in` before its mask/end-activation: 3 x [Linear, LayerNorm, (Dropout)] with ReLU between
    (models/modules/mlp.py:36-64 with use_layernorm, end_layer_activation; indices 0,1,4,5,8,9)."""
    p = "model.add_goal.mlp_in.fc_layers"
    z = torch.relu(layer_norm(linear(z, sd, f"{p}.0"), sd, f"{p}.1"))
    z = torch.relu(layer_norm(linear(z, sd, f"{p}.4"), sd, f"{p}.5"))
    return layer_norm(linear(z, sd, f"{p}.8"), sd, f"{p}.9")


def latent_mlp_in(sd: SD, z: Tensor) -> Tensor:
    """`add_latent.mlp_in` before its mask/end-activation: Linear(16,128)-ReLU-Linear(128,128)."""
    p = "model.add_latent.mlp_in.fc_layers"
    return linear(torch.relu(linear(z, sd, f"{p}.0")), sd, f"{p}.3")


def add_latent_goal(sd: SD, prefix: str, x: Tensor, x_valid: Tensor, z_in: Tensor, z_valid: Tensor) -> Tensor:
    """`AddLatentGoal.forward`, mode cat, res_add (models/modules/add_latent_goal.py:57-77).  `z_in` is the
    output of the last Linear/LayerNorm of `mlp_in`; `MLP.forward` then masks and applies ReLU (mlp.py:80-85)."""
    z = torch.relu(z_in.masked_fill(~z_valid.unsqueeze(-1), 0.0))
    p = prefix + ".mlp_out.fc_layers"
    h = torch.relu(linear(torch.relu(linear(torch.cat([x, z], -1), sd, f"{p}.0")), sd, f"{p}.3"))
    h = h.masked_fill(~z_valid.unsqueeze(-1), 0.0) + x
    return h.masked_fill(~x_valid.unsqueeze(-1), 0.0)


def action_head(sd: SD, x: Tensor, valid: Tensor, agent_type: Tensor) -> Tuple[Tensor, Tensor]:
    """`ActionHead.forward`, branch_type, fixed log_std (models/modules/action_head.py:70-87)."""
    mask_type = agent_type & valid.unsqueeze(-1)
    mean = 0
    log_std = 0
    for c in range(3):
        p = f"action_head.mlp_mean.{c}.fc_layers"
        m = linear(torch.relu(linear(x, sd, f"{p}.0")), sd, f"{p}.2")
        mean = mean + m.masked_fill(~mask_type[..., c].unsqueeze(-1), 0.0)
        ls = sd[f"action_head.log_std.{c}"][None, None, :].expand(*valid.shape, -1)
        log_std = log_std + ls.masked_fill(~mask_type[..., [c]], 0.0)
    return mean, log_std


def diag_gauss_log_prob(value: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """`Independent(Normal(mean, std), 1).log_prob` (torch.distributions.Normal.log_prob), summed over the last dim."""
    var = std ** 2
    return (-((value - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)


def cast_rad(a: Tensor) -> Tensor:
    """utils/transform_utils.py:10-12"""
    return (a + math.pi) % (2 * math.pi) - math.pi


# ----------------------------------------------------------------------------------------------------------
# scene encoding (once per scene)
# ----------------------------------------------------------------------------------------------------------
def map_encoder(sd: SD, map_valid: Tensor, map_type: Tensor, map_pos: Tensor, map_dir: Tensor) -> Tuple[Tensor, Tensor]:
    """`SceneCentricInput.forward` map part (data_modules/sc_input.py:124-134) + `MapEncoder.forward`
    (models/modules/map_encoder.py:72-115, densetnt_vectornet, max pool).
    map_valid [S,P,N] bool, map_type [S,P,11] bool, map_pos/dir [S,P,N,2] -> (map_feature [S,P,D], pl_valid [S,P])."""
    S, P, N = map_valid.shape
    f_xy = sd["pre_processing.input.pose_pe_map.pe_xy.freqs"]
    f_yaw = sd["pre_processing.input.pose_pe_map.pe_yaw.freqs"]
    ohe = sd["pre_processing.input.pl_node_ohe"]
    attr = torch.cat([map_type.unsqueeze(-2).expand(-1, -1, N, -1), ohe[None, None].expand(S, P, -1, -1)], -1)
    pe = pose_pe(map_pos, torch.atan2(map_dir[..., 1], map_dir[..., 0]), f_xy, f_yaw)
    x = input_pe_encoder(sd, "model.map_encoder.input_pe_encoder", map_valid, attr, pe).flatten(0, 1)  # [S*P,N,D]
    v = map_valid.flatten(0, 1)
    x = tf_block(sd, "model.map_encoder.transformer_densetnt", 3, x, ~v, x, ~v)  # :78-84 tgt = initial features
    x = x.view(S, P, N, D).masked_fill(~map_valid.unsqueeze(-1), float("-inf")).amax(dim=2)  # :95-97
    pl_valid = map_valid.any(-1)
    x = x.masked_fill(~pl_valid.unsqueeze(-1), 0.0)  # :105-106
    x = tf_block(sd, "model.map_encoder.transformer_self_attn", 1, x, ~pl_valid, x, ~pl_valid)  # :108-114
    return x, pl_valid


def agent_attr(vel: Tensor, spd: Tensor, yaw_rate: Tensor, acc: Tensor, size: Tensor, a_type: Tensor) -> Tensor:
    """data_modules/sc_input.py:153-163 -- [vel2, spd1, yaw_rate1, acc1, size3, type3]"""
    return torch.cat([vel, spd, yaw_rate, acc, size, a_type.to(vel.dtype)], dim=-1)


def encode_agents(sd: SD, valid: Tensor, pos: Tensor, yaw: Tensor, vel: Tensor, spd: Tensor, yaw_rate: Tensor,
                  acc: Tensor, size: Tensor, a_type: Tensor) -> Tensor:
    """history agent encoder: data_modules/sc_input.py:109-122 + `agent_encoder` (traffic_bots.py:149).
    valid [S,T,A]; pos [S,T,A,2]; yaw/spd/yaw_rate/acc [S,T,A,1]; size/type [S,A,3] -> [S,T,A,D]."""
    T = valid.shape[1]
    attr = agent_attr(vel, spd, yaw_rate, acc, size.unsqueeze(1).expand(-1, T, -1, -1),
                      a_type.unsqueeze(1).expand(-1, T, -1, -1))
    pe = pose_pe(pos, yaw.squeeze(-1), sd["pre_processing.input.pose_pe_agent.pe_xy.freqs"],
                 sd["pre_processing.input.pose_pe_agent.pe_yaw.freqs"])
    return input_pe_encoder(sd, "model.agent_encoder", valid, attr, pe)


def encode_tl(sd: SD, valid: Tensor, state: Tensor, pos: Tensor, dir_: Tensor) -> Tensor:
    """traffic-light encoder: data_modules/sc_input.py:136-139 + `tl_encoder` (traffic_bots.py:150)."""
    pe = pose_pe(pos, torch.atan2(dir_[..., 1], dir_[..., 0]), sd["pre_processing.input.pose_pe_tl.pe_xy.freqs"],
                 sd["pre_processing.input.pose_pe_tl.pe_yaw.freqs"])
    return input_pe_encoder(sd, "model.tl_encoder", valid, state.to(pos.dtype), pe)


def encode_scene(sd: SD, batch: Dict[str, Tensor], n_step_hist: int = 11) -> Dict[str, Tensor]:
    """`SceneCentricPreProcessing` + `SceneCentricInput` + `TrafficBots.encode_input_features`
    (data_modules/scene_centric.py:100-133, sc_input.py:98-140, models/traffic_bots.py:146-151) in eval mode."""
    h = slice(0, n_step_hist)
    f: Dict[str, Tensor] = {}
    f["map_feature"], f["map_feature_valid"] = map_encoder(
        sd, batch["map/valid"], batch["map/type"], batch["map/pos"], batch["map/dir"])
    f["agent_feature_valid"] = batch["agent/valid"][:, h]
    f["agent_feature"] = encode_agents(
        sd, batch["agent/valid"][:, h], batch["agent/pos"][:, h], batch["agent/yaw_bbox"][:, h],
        batch["agent/vel"][:, h], batch["agent/spd"][:, h], batch["agent/yaw_rate"][:, h], batch["agent/acc"][:, h],
        batch["agent/size"], batch["agent/type"])
    f["tl_feature_valid"] = batch["tl_stop/valid"][:, h]
    f["tl_feature"] = encode_tl(sd, batch["tl_stop/valid"][:, h], batch["tl_stop/state"][:, h],
                                batch["tl_stop/pos"][:, h], batch["tl_stop/dir"][:, h])
    return f


def temporal_max_valid(x: Tensor, valid: Tensor) -> Tuple[Tensor, Tensor]:
    """`TemporalAggregate` mode max_valid (models/modules/agent_temporal.py:31-32,43-44)."""
    agg = x.masked_fill(~valid.unsqueeze(-1), -1e3).amax(1)
    v = valid.any(1)
    return agg.masked_fill(~v.unsqueeze(-1), 0.0), v


def temporal_last_valid(x: Tensor, valid: Tensor) -> Tuple[Tensor, Tensor]:
    """`TemporalAggregate` mode last_valid (models/modules/agent_temporal.py:33-36,43-44)."""
    B, T, A = valid.shape
    idx = T - 1 - torch.max(valid.flip(1).to(torch.uint8), dim=1)[1]
    agg = x[torch.arange(B).unsqueeze(1), idx, torch.arange(A).unsqueeze(0)]
    v = valid.any(1)
    return agg.masked_fill(~v.unsqueeze(-1), 0.0), v


def latent_encoder(sd: SD, feat: Dict[str, Tensor], posterior: bool = False, down: int = 5) -> Dict[str, Tensor]:
    """`LatentEncoder.forward` (models/latent_encoder.py:95-147) + `DistEncoder.forward` diag_gaus with a learned
    constant log_std (:195-199).  Returns mean [S,A,16], std [S,A,16] (broadcast), valid [S,A]."""
    which = "post" if posterior else "prior"
    av = feat["agent_feature_valid"][:, ::down]
    af = feat["agent_feature"][:, ::down]
    tv = feat["tl_feature_valid"][:, ::down]
    tf_ = feat["tl_feature"][:, ::down]
    S, T, A, _ = af.shape
    x = tf_block(sd, "model.transformer_as2pl", 3, af.flatten(1, 2), ~av.flatten(1, 2), feat["map_feature"],
                 ~feat["map_feature_valid"]).view(S, T, A, D)  # :108-114
    x = tf_block(sd, "model.transformer_as2tl", 3, x.flatten(0, 1), ~av.flatten(0, 1), tf_.flatten(0, 1),
                 ~tv.flatten(0, 1)).view(S, T, A, D)  # :116-122
    x = interaction(sd, f"model.latent_encoder.agent_interaction_{which}", x.flatten(0, 1), av.flatten(0, 1))
    x = gru_sequence(sd, f"model.latent_encoder.agent_temporal_{which}", x.view(S, T, A, D), av)
    x, v = temporal_max_valid(x, av)
    p = f"model.latent_encoder.latent_{which}_dist"
    mean = linear(torch.relu(linear(x, sd, f"{p}.mlp_mean.fc_layers.0")), sd, f"{p}.mlp_mean.fc_layers.2")
    mean = mean.masked_fill(~v.unsqueeze(-1), 0.0)  # MLP.forward valid_mask (mlp.py:81-82)
    std = sd[f"{p}.log_std"].exp().expand_as(mean)
    return {"mean": mean, "std": std, "valid": v}


def dest_predictor(sd: SD, feat: Dict[str, Tensor], agent_type: Tensor, map_type: Tensor) -> Dict[str, Tensor]:
    """`DestPredictor.forward`, mode mlp (models/goal_manager.py:228-246,294-307,328-333) and the
    `Categorical(logits=...)` normalisation of `DestCategorical` (models/modules/distributions.py:161-165).
    Returns probs [S,A,P], log-probs [S,A,P], valid [S,A]."""
    af, av = feat["agent_feature"], feat["agent_feature_valid"]
    mf, mv = feat["map_feature"], feat["map_feature_valid"]
    S, P, _ = mf.shape
    A = av.shape[2]
    type_mask = ~(mv & map_type[:, :, :5].any(-1))
    m_veh = agent_type[:, :, [0]] & map_type[:, :, 3].unsqueeze(1)
    m_ped = agent_type[:, :, [1]] & map_type[:, :, :4].any(-1).unsqueeze(1)
    m_cyc = agent_type[:, :, [2]] & map_type[:, :, :3].any(-1).unsqueeze(1)
    attn_mask = m_veh | m_ped | m_cyc
    dist_valid = av.any(1)
    tgt = gru_sequence(sd, "model.goal_manager.goal_predictor.gru_as", af, av) + af  # :298-300
    tgt, _ = temporal_last_valid(tgt, av)
    x = torch.cat([mf.unsqueeze(1).expand(-1, A, -1, -1), tgt.unsqueeze(2).expand(-1, -1, P, -1)], dim=-1)
    p = "model.goal_manager.goal_predictor.mlp.fc_layers"
    x = torch.relu(layer_norm(linear(x, sd, f"{p}.0"), sd, f"{p}.1"))
    x = torch.relu(layer_norm(linear(x, sd, f"{p}.3"), sd, f"{p}.4"))
    logits = linear(x, sd, f"{p}.6").squeeze(-1)
    logits = logits.masked_fill(type_mask.unsqueeze(1), float("-inf"))
    logits = logits.masked_fill(attn_mask, float("-inf"))
    logits = logits.masked_fill(~dist_valid.unsqueeze(-1), 0.0)
    logits = logits.masked_fill((logits == float("-inf")).all(-1).unsqueeze(-1), 0.0)
    logp = logits - logits.logsumexp(dim=-1, keepdim=True)
    return {"probs": torch.softmax(logp, dim=-1), "logp": logp, "valid": dist_valid}


def sample_dest(probs: Tensor, deterministic: Tensor) -> Tuple[Tensor, Tensor]:
    """`DestCategorical.repeat_interleave_` + `.sample(deterministic)` + `.log_prob`
    (models/modules/distributions.py:174-201): `probs` is already repeated per mode; `Categorical(probs=)`
    re-normalises, samples with `torch.multinomial`.  Consumes the global torch RNG like the reference."""
    p = probs / probs.sum(-1, keepdim=True)
    det = p.argmax(-1)
    rnd = torch.multinomial(p.reshape(-1, p.shape[-1]), 1, True).T.reshape(p.shape[:-1])
    sample = det.masked_fill(~deterministic, 0) + rnd.masked_fill(deterministic, 0)
    eps = torch.finfo(p.dtype).eps
    logp = torch.log(p.clamp(min=eps, max=1 - eps))  # torch.distributions.utils.probs_to_logits
    return sample, logp.gather(-1, sample.unsqueeze(-1)).squeeze(-1)


def sample_latent(mean: Tensor, std: Tensor, deterministic: Tensor) -> Tuple[Tensor, Tensor]:
    """`MyDist.sample(deterministic tensor)` + `log_prob` for the DiagGaussian latent
    (models/modules/distributions.py:19-37; models/traffic_bots.py:196-199).  Consumes the global torch RNG."""
    eps = torch.empty_like(mean).normal_()
    rnd = mean + eps * std
    sample = mean.masked_fill(~deterministic.unsqueeze(-1), 0) + rnd.masked_fill(deterministic.unsqueeze(-1), 0)
    return sample, diag_gauss_log_prob(sample, mean, std)


def teacher_forcing_mask(valid: Tensor, step_spawn_agent: int, step_warm_start: int) -> Tensor:
    """`TeacherForcing.get` without the (disabled) schedules (utils/teacher_forcing.py:43-56).  valid [B,T,A]."""
    m = torch.zeros_like(valid)
    m[:, 0] |= valid[:, 0]
    if step_spawn_agent > 0:
        spawn = (~valid[:, :-1]) & valid[:, 1:]
        spawn[:, step_spawn_agent:] = False
        m[:, 1:] |= spawn
    if step_warm_start >= 0:
        m[:, : step_warm_start + 1] |= valid[:, : step_warm_start + 1]
    return m


# ----------------------------------------------------------------------------------------------------------
# the closed-loop rollout
# ----------------------------------------------------------------------------------------------------------
def policy_step(sd: SD, agent_feature: Tensor, valid: Tensor, map_feature: Tensor, map_valid: Tensor,
                tl_feature: Tensor, tl_valid: Tensor, goal_feature: Tensor, goal_valid: Tensor,
                latent_sample: Tensor, hidden: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """`TrafficBots.forward` with interaction_first, add_goal_latent_first=False (models/traffic_bots.py:201-241)."""
    x = tf_block(sd, "model.transformer_as2pl", 3, agent_feature, ~valid, map_feature, ~map_valid)
    x = tf_block(sd, "model.transformer_as2tl", 3, x, ~valid, tl_feature, ~tl_valid)
    x = interaction(sd, "model.agent_interaction", x, valid)
    x, hidden = gru_step(sd, "model.agent_temporal", x, valid, hidden)
    x = add_latent_goal(sd, "model.add_goal", x, valid, goal_mlp_in(sd, goal_feature), goal_valid)
    x = add_latent_goal(sd, "model.add_latent", x, valid, latent_mlp_in(sd, latent_sample), valid)
    return x, hidden


@torch.no_grad()
def rollout(sd: SD, *, map_feature: Tensor, map_valid: Tensor, tl_feature: Tensor, tl_valid: Tensor,
            gt_valid: Tensor, gt_state: Tensor, gt_vel: Tensor, gt_acc: Tensor, gt_yaw_rate: Tensor,
            agent_type: Tensor, agent_size: Tensor, tf_mask: Tensor, latent_sample: Tensor, latent_logp: Tensor,
            dest: Tensor, goal_valid: Tensor, goal_gt: Optional[Tensor], map_boundary: Tensor,
            raw_map_valid: Tensor, raw_map_type: Tensor, raw_map_pos: Tensor, raw_map_dir: Tensor,
            step_start: int = 1, step_end: int = 90, return_trace: bool = False,
            rules: Optional[Dict] = None, w_collision: float = 0.0, reduce_collision_with_max: bool = True
            ) -> Dict[str, Tensor]:
    """`WaymoMotion.rollout` + `.forward` in eval mode with the default config (pl_modules/waymo_motion.py:108-354),
    including `Dynamics` (utils/dynamics.py:29-167), `MultiPathPP` (:187-228), the always-on checks of
    `TrafficRuleChecker` (utils/traffic_rule_checker.py:82-119,338-516), `disable_goal_reached`
    (models/goal_manager.py:155-161) and the IL part of `DifferentiableReward.get` (utils/rewards.py:117-131).

    `rules` (optional checks): dict(enable={collided, run_road_edge, run_red_light, passive: bool}, tl_valid,
    tl_pos, tl_state [B,T_tl,TL,..]) -> `rule_checks_oracle.check_optional` every step; `w_collision` > 0 adds the collision
    term of the reward (`rule_checks_oracle.collision_term`).

    Every tensor's leading dim is B = n_scene * K scene-modes (the reference `repeat_interleave`s everything,
    waymo_motion.py:547-548).  gt_* have T_gt frames (91 for train/val, 11 for test); tl_* have T_tl frames.
    """
    B, A = agent_type.shape[:2]
    T_gt = gt_valid.shape[1]
    f_xy = sd["pre_processing.input.pose_pe_agent.pe_xy.freqs"]
    f_yaw = sd["pre_processing.input.pose_pe_agent.pe_yaw.freqs"]
    bidx = torch.arange(B).unsqueeze(1)

    # Dynamics.init with frame 0 (waymo_motion.py:251-259)
    valid = gt_valid[:, 0].clone()
    killed = torch.zeros_like(valid)
    state = gt_state[:, 0].clone()
    vel, acc, yaw_rate = gt_vel[:, 0].clone(), gt_acc[:, 0].clone(), gt_yaw_rate[:, 0].clone()
    hidden = None
    goal_valid = goal_valid.clone()
    goal_feature = map_feature[bidx, dest]  # goal_manager.py:131-138

    # TrafficRuleChecker.__init__ (traffic_rule_checker.py:45-51,74-98)
    outside_map = torch.zeros_like(valid)
    goal_reached = torch.zeros_like(valid)
    dest_reached = torch.zeros_like(valid)
    goal_thresh_pos = agent_size[:, :, 0] * 8
    dest_valid = raw_map_valid[bidx, dest]
    dest_type = raw_map_type[bidx, dest]
    dest_pos = raw_map_pos[bidx, dest]
    dest_dir = raw_map_dir[bidx, dest]
    dest_dir = dest_dir / torch.norm(dest_dir, dim=-1, keepdim=True)
    dest_thresh_pos = torch.ones_like(agent_size[:, :, 0]) * 50
    dest_thresh_pos = dest_thresh_pos * (1 - dest_type[:, :, 4] * 0.8)
    max_acc = torch.tensor(MAX_ACC)
    max_yaw = torch.tensor(MAX_YAW_RATE)
    type_f = agent_type.to(torch.float32)

    keys = ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid", "latent_log_probs",
            "action_log_probs", "outside_map", "outside_map_this_step", "goal_reached", "goal_reached_this_step",
            "dest_reached", "dest_reached_this_step")
    out = {k: [] for k in keys}
    rs = None
    if rules is not None:
        import rule_checks_oracle as rco
        rs = rco.init_rules(agent_type, agent_size, raw_map_valid, raw_map_type, raw_map_pos, raw_map_dir, rules["tl_valid"],
                            rules["tl_pos"], rules["tl_state"], rules["enable"])
        for k in rco.OPTIONAL_KEYS:
            out[k] = []
            out[k + "_this_step"] = []
    trace = {"policy_feature": [], "action_mean": [], "goal_valid": [], "agent_valid_post": []}
    zeros_b = torch.zeros_like(valid)

    for t in range(step_start, step_end + 1):
        ovr = tf_mask[:, t] if t < T_gt else zeros_b  # waymo_motion.py:271-274
        tl_t = min(t - 1, tl_valid.shape[1] - 1)  # :287

        # ---- forward(): state embedding (sc_input.py:142-165; stale vel/acc/yaw_rate) ----
        attr = agent_attr(vel, state[..., 3:4], yaw_rate, acc, agent_size, agent_type)
        pe = pose_pe(state[..., :2], state[..., 2], f_xy, f_yaw)
        feat = input_pe_encoder(sd, "model.agent_encoder", valid, attr, pe)
        x, hidden = policy_step(sd, feat, valid, map_feature, map_valid, tl_feature[:, tl_t], tl_valid[:, tl_t],
                                goal_feature, goal_valid, latent_sample, hidden)
        mean, log_std = action_head(sd, x, valid, agent_type)

        # ---- Dynamics.update (dynamics.py:74-119), deterministic action = mean ----
        a_logp = diag_gauss_log_prob(mean, mean, log_std.exp()).masked_fill(~valid, 0.0)
        tanh_a = torch.tanh(mean)
        action = torch.stack([tanh_a[..., 0] * (type_f * max_acc).sum(-1), tanh_a[..., 1] * (type_f * max_yaw).sum(-1)],
                             dim=-1).masked_fill(~valid.unsqueeze(-1), 0.0)
        a_acc, a_yr = action[..., 0], action[..., 1]
        v_tilde = state[..., 3] + 0.5 * DT * a_acc  # dynamics.py:210-211
        th_tilde = state[..., 2] + 0.5 * DT * a_yr
        delta = torch.stack([v_tilde * torch.cos(th_tilde), v_tilde * torch.sin(th_tilde), a_yr, a_acc], dim=-1)
        has_type = agent_type.any(-1, keepdim=True)  # the per-type masked sum drops agents without a type
        pred_state = ((state + DT * delta) * has_type).masked_fill(~valid.unsqueeze(-1), 0.0)
        pred_valid = valid

        # ---- Dynamics.override_states (dynamics.py:132-149) ----
        m = ovr & ~killed
        valid = valid | m
        if t < T_gt:
            mm = m.unsqueeze(-1)
            state = torch.where(mm, gt_state[:, t], pred_state)
            vel = torch.where(mm, gt_vel[:, t], vel)
            acc = torch.where(mm, gt_acc[:, t], acc)
            yaw_rate = torch.where(mm, gt_yaw_rate[:, t], yaw_rate)
        else:
            state = pred_state

        # ---- TrafficRuleChecker.check, always-on subset (traffic_rule_checker.py:423-424,474-496) ----
        px, py = state[..., 0], state[..., 1]
        out_t = ((px > map_boundary[:, [1]]) | (px < map_boundary[:, [0]]) | (py > map_boundary[:, [3]])
                 | (py < map_boundary[:, [2]])) & valid
        outside_map = outside_map | out_t
        if goal_gt is None:
            goal_t = torch.zeros_like(valid)
        else:
            pos_ok = torch.norm(state[..., :2] - goal_gt[..., :2], dim=-1) < goal_thresh_pos
            rot_ok = torch.abs(cast_rad(state[..., 2] - goal_gt[..., 2])) < math.radians(15)
            goal_t = pos_ok & rot_ok & valid & ~goal_reached
        goal_reached = goal_reached | goal_t
        dist = torch.norm(state[..., :2].unsqueeze(2) - dest_pos, dim=-1).masked_fill(~dest_valid, 1e4)
        pos_reached = (dist < dest_thresh_pos.unsqueeze(-1)).any(-1)
        head = torch.stack([torch.cos(state[..., 2]), torch.sin(state[..., 2])], dim=-1)
        rot = (head.unsqueeze(2) * dest_dir).sum(-1).masked_fill(~dest_valid, 0.0)
        rot_reached = (rot > math.cos(math.radians(30))).any(-1)
        lane = dest_type[:, :, :4].any(-1)
        edge = dest_type[:, :, 4]
        dest_t = ~dest_reached & valid & ((lane & pos_reached & rot_reached) | (edge & pos_reached))
        dest_reached = dest_reached | dest_t
        if rs is not None:  # optional checks see the same post-override state / valid (traffic_rule_checker.py:420-472)
            for k, v in rco.check_optional(rs, t, valid, state).items():
                out[k].append(v)

        # ---- Dynamics.kill (dynamics.py:161-167), disable_goal_reached (goal_manager.py:155-161) ----
        kill = out_t & ~gt_valid[:, t] if t < T_gt else out_t
        killed = killed | kill
        valid = valid & ~kill
        goal_valid = goal_valid & valid & ~dest_reached

        # ---- DifferentiableReward.get: collision term (rewards.py:49-115), IL part (:117-131) ----
        reward0 = torch.zeros_like(px)
        if w_collision > 0:
            import rule_checks_oracle as rco2
            col = rco2.collision_term(pred_valid, pred_state, agent_size, reduce_collision_with_max)
            reward0 = reward0 - w_collision * col.masked_fill(~pred_valid, 0.0)
        if t < T_gt:
            rv = pred_valid & gt_valid[:, t]
            gs = gt_state[:, t].masked_fill(~rv.unsqueeze(-1), 0.0)
            ps = pred_state.masked_fill(~rv.unsqueeze(-1), 0.0)
            e_pos = F.smooth_l1_loss(gs[..., :2], ps[..., :2], reduction="none").sum(-1)
            e_rot = 0.5 * (1 - torch.cos(gs[..., 2] - ps[..., 2]))
            e_spd = F.smooth_l1_loss(gs[..., 3], ps[..., 3], reduction="none")
            reward = (reward0 - (0.1 * e_pos + 10.0 * e_rot + 0.1 * e_spd)).masked_fill(~rv, 0.0)
        else:
            rv = pred_valid
            reward = reward0.masked_fill(~rv, 0.0)

        for k, v in (("preds", pred_state), ("valid", pred_valid), ("override_masks", ovr),
                     ("diffbar_rewards", reward), ("diffbar_rewards_valid", rv), ("latent_log_probs", latent_logp),
                     ("action_log_probs", a_logp), ("outside_map", outside_map), ("outside_map_this_step", out_t),
                     ("goal_reached", goal_reached), ("goal_reached_this_step", goal_t),
                     ("dest_reached", dest_reached), ("dest_reached_this_step", dest_t)):
            out[k].append(v)
        if return_trace:
            trace["policy_feature"].append(x)
            trace["action_mean"].append(mean)
            trace["goal_valid"].append(goal_valid)
            trace["agent_valid_post"].append(valid)

    res = {k: torch.stack(v, dim=2) for k, v in out.items()}  # RolloutBuffer.finish (utils/buffer.py:72-90)
    res["hidden"] = hidden
    res["final_state"] = state
    res["final_valid"] = valid
    if return_trace:
        for k, v in trace.items():
            res["trace/" + k] = torch.stack(v, dim=2)
    return res


# ----------------------------------------------------------------------------------------------------------
# end-to-end drivers mirroring validation_step / test_step
# ----------------------------------------------------------------------------------------------------------
def rollout_inputs(batch: Dict[str, Tensor], feat: Dict[str, Tensor], k: int, test_mode: bool = False) -> Dict:
    """tensor plumbing of `joint_future_pred` / `reactive_replay` (waymo_motion.py:434-466,523-548): picks the
    GT tensors used for overriding (91 frames, or the 11 history frames in test mode) and repeats per mode."""
    T = 11 if test_mode else batch["agent/valid"].shape[1]
    ri = lambda x: x.repeat_interleave(k, 0)  # noqa: E731
    return dict(
        map_feature=ri(feat["map_feature"]), map_valid=ri(feat["map_feature_valid"]),
        tl_feature=ri(feat["tl_feature"]), tl_valid=ri(feat["tl_feature_valid"]),
        gt_valid=ri(batch["agent/valid"][:, :T]),
        gt_state=ri(torch.cat([batch["agent/pos"], batch["agent/yaw_bbox"], batch["agent/spd"]], -1)[:, :T]),
        gt_vel=ri(batch["agent/vel"][:, :T]), gt_acc=ri(batch["agent/acc"][:, :T]),
        gt_yaw_rate=ri(batch["agent/yaw_rate"][:, :T]), agent_type=ri(batch["agent/type"]),
        agent_size=ri(batch["agent/size"]), map_boundary=ri(batch["map/boundary"]),
        raw_map_valid=ri(batch["map/valid"]), raw_map_type=ri(batch["map/type"]), raw_map_pos=ri(batch["map/pos"]),
        raw_map_dir=ri(batch["map/dir"]),
    )


@torch.no_grad()
def joint_future_pred(sd: SD, batch: Dict[str, Tensor], k: int = 6, sample_seed: Optional[int] = 7,
                      step_end: int = 90, test_mode: bool = False, feat: Optional[Dict[str, Tensor]] = None,
                      return_trace: bool = False, rules_enable: Optional[Dict[str, bool]] = None, w_collision: float = 0.0,
                      reduce_collision_with_max: bool = True) -> Dict[str, Tensor]:
    """encode -> prior latent -> dest prediction -> K-mode closed-loop rollout, as `validation_step`'s
    joint_future_pred leg (waymo_motion.py:581-598,683-690) / `test_step` (:905-934).
    Output layout after `flatten_repeat` (utils/buffer.py:92-123): [S, A, K, T, ...]."""
    if feat is None:
        feat = encode_scene(sd, batch)
    S, A = batch["agent/type"].shape[:2]
    prior = latent_encoder(sd, feat, posterior=False)
    dest = dest_predictor(sd, feat, batch["agent/type"], batch["map/type"])
    goal_valid = feat["agent_feature_valid"].any(1)
    det = torch.zeros(S * k, A, dtype=torch.bool)
    det[::k] = True  # mode 0 of every scene is deterministic (waymo_motion.py:489-491)
    if sample_seed is not None:
        torch.manual_seed(sample_seed)
    dest_sample, dest_logp = sample_dest(dest["probs"].repeat_interleave(k, 0), det)  # :497-500
    rin = rollout_inputs(batch, feat, k, test_mode)
    tf_mask = teacher_forcing_mask(rin["gt_valid"], 10, 10)  # teacher_forcing_joint_future_pred
    lat_sample, lat_logp = sample_latent(prior["mean"].repeat_interleave(k, 0), prior["std"].repeat_interleave(k, 0), det)
    goal_gt = None if "agent/goal" not in batch or test_mode else batch["agent/goal"].repeat_interleave(k, 0)
    rules = None
    if rules_enable is not None:  # the history traffic lights (waymo_motion.py:523-525): frozen after frame 10
        rules = dict(enable=rules_enable, tl_valid=batch["history/tl_stop/valid"].repeat_interleave(k, 0),
                     tl_pos=batch["history/tl_stop/pos"].repeat_interleave(k, 0),
                     tl_state=batch["history/tl_stop/state"].repeat_interleave(k, 0))
    res = rollout(sd, **rin, tf_mask=tf_mask, latent_sample=lat_sample, latent_logp=lat_logp, dest=dest_sample,
                  goal_valid=goal_valid.repeat_interleave(k, 0), goal_gt=goal_gt, step_end=step_end,
                  return_trace=return_trace, rules=rules, w_collision=w_collision,
                  reduce_collision_with_max=reduce_collision_with_max)
    out = {}
    for name, v in res.items():
        if name in ("hidden", "final_state", "final_valid"):
            out[name] = v
        else:
            out[name] = v.view(S, k, *v.shape[1:]).transpose(1, 2)
    out["goal_sample"] = dest_sample.view(S, k, A).transpose(1, 2)
    out["goal_log_probs"] = dest_logp.view(S, k, A).transpose(1, 2)
    out["latent_sample"] = lat_sample
    out["dest_probs"] = dest["probs"]
    out["latent_prior_mean"] = prior["mean"]
    return out


@torch.no_grad()
def reactive_replay(sd: SD, batch: Dict[str, Tensor], step_end: int = 90,
                    feat: Optional[Dict[str, Tensor]] = None, return_trace: bool = False,
                    rules_enable: Optional[Dict[str, bool]] = None, w_collision: float = 0.0,
                    reduce_collision_with_max: bool = True) -> Dict[str, Tensor]:
    """`validation_step`'s reactive-replay leg (waymo_motion.py:597-611): posterior latent (deterministic = mean),
    GT destination, agents spawn over the whole episode (teacher_forcing_reactive_replay)."""
    if feat is None:
        feat = encode_scene(sd, batch)
    # posterior inputs = the full 91-frame episode (data_modules/sc_latent.py:166-168,211-236)
    feat_post = dict(feat)
    feat_post["agent_feature_valid"] = batch["agent/valid"]
    feat_post["agent_feature"] = encode_agents(
        sd, batch["agent/valid"], batch["agent/pos"], batch["agent/yaw_bbox"], batch["agent/vel"], batch["agent/spd"],
        batch["agent/yaw_rate"], batch["agent/acc"], batch["agent/size"], batch["agent/type"])
    feat_post["tl_feature_valid"] = batch["tl_stop/valid"]
    feat_post["tl_feature"] = encode_tl(sd, batch["tl_stop/valid"], batch["tl_stop/state"], batch["tl_stop/pos"],
                                        batch["tl_stop/dir"])
    post = latent_encoder(sd, feat_post, posterior=True)
    goal_valid = feat["agent_feature_valid"].any(1)
    rin = rollout_inputs(batch, feat, 1)
    tf_mask = teacher_forcing_mask(rin["gt_valid"], 90, 10)
    lat_logp = diag_gauss_log_prob(post["mean"], post["mean"], post["std"])
    rules = None
    if rules_enable is not None:  # the full-episode traffic lights (waymo_motion.py:439-441)
        rules = dict(enable=rules_enable, tl_valid=batch["tl_stop/valid"], tl_pos=batch["tl_stop/pos"],
                     tl_state=batch["tl_stop/state"])
    res = rollout(sd, **rin, tf_mask=tf_mask, latent_sample=post["mean"], latent_logp=lat_logp,
                  dest=batch["agent/dest"], goal_valid=goal_valid, goal_gt=batch["agent/goal"], step_end=step_end,
                  return_trace=return_trace, rules=rules, w_collision=w_collision,
                  reduce_collision_with_max=reduce_collision_with_max)
    res["latent_post_mean"] = post["mean"]
    return res

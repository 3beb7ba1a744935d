"""The reference's `training_step` (pl_modules/waymo_motion.py:356-418) composed from the primitives of `tape.Fn`.

Forward structure (every function cites the reference code it follows; `src/` relative paths):
  encode_input_features x3 on aliased inputs (:366-368) -> ONE map encoding + history / posterior agent and traffic-light
  encodings; pred_goal (:374-379, features detached: goal_manager.py:222-224); posterior + prior latent encoder (:382-383);
  latent choice (:384-387); rollout with teacher forcing up to t = 10, reparameterised latent sample, deterministic action
  (:390-400, :205-354); TrainingMetrics (models/metrics/training.py:62-158).
Loop-invariant work is shared instead of repeated (the K|V projections of the map / traffic-light keys, `mlp_in` of the goal
and latent features): gradients of all uses accumulate in the shared buffer, which is mathematically what autograd does with
the reference's repeated evaluation.

Dropout is not implemented: the step is the reference's with every dropout probability set to 0 (the parity configuration
of BASELINE.json configs[3]).  Layouts: every buffer is 2-D [rows, 128]; agent rows are scene-major [S, A], episodes are
frame-major [T, S, A] where a recurrent loop walks over frames.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from .tape import Fn, ReplayOps, StepStack, Var

D = 128
U8 = torch.uint8


class Params:
    """name -> Var over the flat parameter / gradient buffers (views; gradients accumulate in place)."""

    def __init__(self, tensors: Dict[str, Tensor], grads: Dict[str, Tensor], buffers: Dict[str, Tensor]) -> None:
        self.t, self.g, self.buffers = tensors, grads, buffers
        self._cache: Dict[tuple, Var] = {}

    def __call__(self, name: str, rows: Optional[Tuple[int, int]] = None, cols: Optional[Tuple[int, int]] = None) -> Var:
        key = (name, rows, cols)
        v = self._cache.get(key)
        if v is None:
            t, g = self.t[name], self.g[name]
            if t.dim() == 1:
                t, g = t.view(1, -1), g.view(1, -1)
                if rows is not None:  # slice of a bias vector
                    t, g = t[:, rows[0]:rows[1]], g[:, rows[0]:rows[1]]
            else:
                if rows is not None:
                    t, g = t[rows[0]:rows[1]], g[rows[0]:rows[1]]
                if cols is not None:
                    t, g = t[:, cols[0]:cols[1]], g[:, cols[0]:cols[1]]
            v = Var(t, True, g)
            self._cache[key] = v
        return v


class Model:
    def __init__(self, fn: Fn, params: Params, drop_seed: Optional[Tensor] = None, drop_p: float = 0.0) -> None:
        """drop_seed: device int32 [1] holding this step's dropout seed; drop_p: the reference's single dropout probability
        (traffic_bots.yaml: tf_cfg.dropout_p = *.mlp_dropout_p = *.dropout_p = agent_temporal.dropout = 0.1).  Every dropout
        site of the forward gets the next site id; the masks are regenerated from (seed, site, element) in the backward."""
        self.f, self.p = fn, params
        self.drop_seed, self.drop_p, self.n_site = drop_seed, float(drop_p), 0

    def dp(self):
        """the next dropout call site, or None when dropout is off."""
        if self.drop_seed is None or self.drop_p <= 0.0:
            return None
        self.n_site += 1
        return (self.drop_seed, self.n_site, self.drop_p)

    # ------------------------------------------------------------------ building blocks
    def lin(self, x: Var, prefix: str, relu: bool = False, drop=None) -> Var:
        return self.f.linear(x, self.p(prefix + ".weight"), self.p(prefix + ".bias"), relu, drop=drop)

    def ln(self, x: Var, prefix: str, relu: bool = False, drop=None) -> Var:
        return self.f.layernorm(x, self.p(prefix + ".weight"), self.p(prefix + ".bias"), relu, drop=drop)

    def input_pe_encoder(self, prefix: str, valid: Tensor, attr: Tensor, pe: Tensor) -> Var:
        """`InputPeEncoder.forward`, pe_mode cat (models/modules/input_pe_encoder.py:52-59); MLP = Linear-Dropout-ReLU-Linear
        (mlp.py:36-64; relu(dropout(x)) == dropout(relu(x)))."""
        x = self.lin(self.lin(Var(attr), prefix + ".mlp.fc_layers.0", relu=True, drop=self.dp()), prefix + ".mlp.fc_layers.3")
        return self.f.add_mask(self.f.cat2(x, Var(pe)), None, valid.reshape(-1))

    def kv_project(self, prefix: str, tgt: Var) -> Var:
        """LN_tgt + the K|V rows of in_proj (models/modules/transformer.py:192, attention.py:86)."""
        t2 = self.ln(tgt, prefix + ".norm_tgt")
        return self.f.linear(t2, self.p(prefix + ".attn.in_proj_weight", rows=(D, 3 * D)),
                             self.p(prefix + ".attn.in_proj_bias", rows=(D, 3 * D)))

    def xlayer(self, prefix: str, src: Var, src_keep: Tensor, kv: Var, key_valid: Tensor, B: int, S: int, T: int,
               eye: bool = False, kv_shared: bool = False) -> Var:
        """`TransformerCrossAttention.forward`, norm_first (models/modules/transformer.py:186-237) with
        `Attention.forward` (models/modules/attention.py:79-146)."""
        f = self.f
        s2 = self.ln(src, prefix + ".norm1")
        q = f.linear(s2, self.p(prefix + ".attn.in_proj_weight", rows=(0, D)), self.p(prefix + ".attn.in_proj_bias", rows=(0, D)))
        o, alive = f.attention(q, kv, key_valid, B, S, T, eye, drop=self.dp(), kv_shared=kv_shared)  # attention.py:131-132
        # out-projection, dead rows forced to 0 (attention.py:144-146), dropout1 + residual (transformer.py:202-205): one Linear
        src = f.linear(o, self.p(prefix + ".attn.out_proj_weight"), self.p(prefix + ".attn.out_proj_bias"),
                       keep_lin=alive, res=src, drop=self.dp())
        s2 = self.ln(src, prefix + ".norm2")
        s2 = self.lin(s2, prefix + ".linear1", relu=True, drop=self.dp())  # linear2(dropout(relu(linear1))) (:214-217)
        # second FFN Linear + dropout2 + residual (:219-222) + zeroing of the invalid source rows (:236-237)
        return f.linear(s2, self.p(prefix + ".linear2.weight"), self.p(prefix + ".linear2.bias"), res=src, keep_out=src_keep,
                        drop=self.dp())

    def tf_block(self, prefix: str, n_layer: int, src: Var, src_keep: Tensor, kvs: List[Var], key_valid: Tensor, B: int, S: int,
                 T: int, eye: bool = False, kv_shared: bool = False) -> Var:
        for i in range(n_layer):
            src = self.xlayer(f"{prefix}.layers.{i}", src, src_keep, kvs[i], key_valid, B, S, T, eye, kv_shared)
        return src

    def decode_front(self, attr: Tensor, pe: Tensor, vflat: Tensor, valid2d: Tensor, kv_map: List[Var], map_valid: Tensor,
                     kv_tl: List[Var], tl_valid: Tensor, B: int, A: int, P: int, TL: int, kv_shared: bool = False) -> Var:
        """the part of a decode step in front of the GRU: state embedding (pl_modules/waymo_motion.py:136-153) and the three
        attention blocks of `TrafficBots.forward` (models/traffic_bots.py:201-226).  Its inputs are gradient-free (the policy
        input is detached), so its backward does not take part in the back-propagation through time."""
        x = self.input_pe_encoder("model.agent_encoder", vflat, attr, pe)
        x = self.tf_block("model.transformer_as2pl", 3, x, vflat, kv_map, map_valid, B, A, P, kv_shared=kv_shared)
        x = self.tf_block("model.transformer_as2tl", 3, x, vflat, kv_tl, tl_valid, B, A, TL)
        return self.interaction("model.agent_interaction", x, valid2d, B, A)

    def interaction(self, prefix: str, x: Var, valid: Tensor, B: int, A: int) -> Var:
        """`MultiAgentTF.forward` (models/modules/agent_interaction.py:51-93).  valid [B, A]."""
        tp = prefix + ".transformer"
        kvs = [self.kv_project(f"{tp}.layers.{i}", x) for i in range(3)]
        y = self.tf_block(tp, 3, x, valid.reshape(-1), kvs, valid, B, A, A, eye=True)
        single = (valid.sum(-1) == 1).to(U8).unsqueeze(-1).expand(-1, A).reshape(-1)  # :61
        return self.f.select_rows(single.contiguous(), x, y)

    def gru_layers(self, prefix: str, x: Var, h: List[Var], keep: Tensor) -> Tuple[Var, List[Var]]:
        """one time step of the 3-layer `nn.GRU` + masks of `MultiAgentGRULoop` (models/modules/agent_temporal.py:133-153)."""
        f = self.f
        inp, hs = x, []
        for layer in range(3):
            gi = f.linear(inp, self.p(f"{prefix}.rnn.weight_ih_l{layer}"), self.p(f"{prefix}.rnn.bias_ih_l{layer}"))
            gh = f.linear(h[layer], self.p(f"{prefix}.rnn.weight_hh_l{layer}"), self.p(f"{prefix}.rnn.bias_hh_l{layer}"))
            inp = f.gru_gates(gi, gh, h[layer])
            hs.append(f.add_mask(inp, None, keep))
            if layer < 2:
                inp = f.dropout(inp, self.dp())  # nn.GRU(dropout=...): on the outputs of every layer but the last
        return hs[2], hs

    def gru_sequence(self, prefix: str, frames: List[Var], valid_tm: Tensor) -> List[Var]:
        """3-D branch (agent_temporal.py:133-146): frames[t] [S*A, D], valid_tm [T, S*A]."""
        zeros = Var(self.f.ops.zeros((frames[0].rows, D), frames[0].data))
        h = [zeros, zeros, zeros]
        outs = []
        for t, x in enumerate(frames):
            y, h = self.gru_layers(prefix, x, h, valid_tm[t].contiguous())
            outs.append(y)
        return outs

    def mlp_in_latent(self, z: Var) -> Var:
        """relu(add_latent.mlp_in(z)) = Linear-Dropout-ReLU-Linear-Dropout, then mask / ReLU in MLP.forward (mlp.py:36-85): the
        row mask and the dropout factor commute with ReLU."""
        lp = "model.add_latent.mlp_in.fc_layers"
        return self.lin(self.lin(z, f"{lp}.0", relu=True, drop=self.dp()), f"{lp}.3", relu=True, drop=self.dp())

    def mlp_in_goal(self, goal_feature: Var) -> Var:
        """relu(add_goal.mlp_in(goal_feature)): 3 x [Linear, LayerNorm, Dropout] with ReLU between / after."""
        gp = "model.add_goal.mlp_in.fc_layers"
        g = self.ln(self.lin(goal_feature, f"{gp}.0"), f"{gp}.1", relu=True, drop=self.dp())
        g = self.ln(self.lin(g, f"{gp}.4"), f"{gp}.5", relu=True, drop=self.dp())
        return self.ln(self.lin(g, f"{gp}.8"), f"{gp}.9", relu=True, drop=self.dp())

    def decode_tail(self, x: Var, goal_in: Var, lat_in: Var, goal_valid: Tensor, vflat: Tensor, a_type: Tensor, hoisted: bool) -> Var:
        """the part of a decode step behind the GRU: add_goal, add_latent (models/modules/add_latent_goal.py:57-77, mode cat,
        res_add) and the action head (models/modules/action_head.py:70-87, branch_type) -> mean of the action distribution.
        goal_in / lat_in: relu(mlp_in(.)) when hoisted (no dropout: loop invariants of the rollout), else the raw goal feature /
        latent sample (the reference draws fresh dropout masks in mlp_in at every step).  Its backward depends on the other steps
        only through the dynamics chain, which consumes `mean`."""
        f = self.f
        zg = goal_in if hoisted else self.mlp_in_goal(goal_in)
        zl = lat_in if hoisted else self.mlp_in_latent(lat_in)
        for name, zr, zv in (("model.add_goal", zg, goal_valid), ("model.add_latent", zl, vflat)):
            zz = f.add_mask(zr, None, zv)
            h = self.lin(self.lin(f.cat2(x, zz), f"{name}.mlp_out.fc_layers.0", relu=True, drop=self.dp()),
                         f"{name}.mlp_out.fc_layers.3", relu=True, drop=self.dp())
            x = f.add_mask(h, x, vflat, keep_a=zv)  # (h * z_valid + x) * x_valid
        mean = None
        for c in range(3):
            ap = f"action_head.mlp_mean.{c}.fc_layers"
            mc = f.add_mask(self.lin(self.lin(x, f"{ap}.0", relu=True), f"{ap}.2"), None, (a_type[:, c] & vflat).contiguous())
            mean = mc if mean is None else f.add_mask(mean, mc, None)
        return mean

    # ------------------------------------------------------------------ encoders
    def map_encoder(self, batch: Dict[str, Tensor]) -> Tuple[Var, Tensor]:
        """data_modules/sc_input.py:124-134 + `MapEncoder.forward` (models/modules/map_encoder.py:72-115)."""
        f, ops = self.f, self.f.ops
        mv = batch["map/valid"]
        S, P, N = mv.shape
        ohe = self.p.buffers["pre_processing.input.pl_node_ohe"]
        attr = torch.cat([batch["map/type"].unsqueeze(-2).expand(-1, -1, N, -1).to(torch.float32),
                          ohe[None, None].expand(S, P, -1, -1)], -1).reshape(S * P * N, -1).contiguous()
        pe = ops.pose_pe(batch["map/pos"].reshape(-1, 2), ops.dir_to_yaw(batch["map/dir"].reshape(-1, 2)),
                         self.p.buffers["pre_processing.input.pose_pe_map.pe_xy.freqs"],
                         self.p.buffers["pre_processing.input.pose_pe_map.pe_yaw.freqs"])
        node_valid = mv.reshape(S * P, N).to(U8).contiguous()
        x0 = self.input_pe_encoder("model.map_encoder.input_pe_encoder", node_valid, attr, pe)
        tp = "model.map_encoder.transformer_densetnt"
        kvs = [self.kv_project(f"{tp}.layers.{i}", x0) for i in range(3)]  # tgt = the initial node features (:78-84)
        x = self.tf_block(tp, 3, x0, node_valid.reshape(-1), kvs, node_valid, S * P, N, N)
        x = f.masked_max(x, node_valid, S * P, N, 1, float("-inf"))  # :95-97, :105-106
        pl_valid = mv.any(-1).to(U8).contiguous()
        tp = "model.map_encoder.transformer_self_attn"
        kv = self.kv_project(f"{tp}.layers.0", x)
        x = self.xlayer(f"{tp}.layers.0", x, pl_valid.reshape(-1), kv, pl_valid, S, P, P)  # :108-114
        return x, pl_valid

    def encode_agents(self, valid: Tensor, pos: Tensor, yaw: Tensor, vel: Tensor, spd: Tensor, yaw_rate: Tensor, acc: Tensor,
                      size: Tensor, a_type: Tensor) -> Var:
        """data_modules/sc_input.py:109-122,153-163 + agent_encoder (traffic_bots.py:149).  Leading dims [..., A] arbitrary."""
        ops = self.f.ops
        attr = torch.cat([vel, spd, yaw_rate, acc, size, a_type.to(torch.float32)], -1).reshape(-1, 11).contiguous()
        pe = ops.pose_pe(pos.reshape(-1, 2).contiguous(), yaw.reshape(-1).contiguous(),
                         self.p.buffers["pre_processing.input.pose_pe_agent.pe_xy.freqs"],
                         self.p.buffers["pre_processing.input.pose_pe_agent.pe_yaw.freqs"])
        return self.input_pe_encoder("model.agent_encoder", valid.reshape(-1).to(U8).contiguous(), attr, pe)

    def encode_tl(self, valid: Tensor, state: Tensor, pos: Tensor, dir_: Tensor) -> Var:
        """data_modules/sc_input.py:136-139 + tl_encoder (traffic_bots.py:150)."""
        ops = self.f.ops
        pe = ops.pose_pe(pos.reshape(-1, 2).contiguous(), ops.dir_to_yaw(dir_.reshape(-1, 2).contiguous()),
                         self.p.buffers["pre_processing.input.pose_pe_tl.pe_xy.freqs"],
                         self.p.buffers["pre_processing.input.pose_pe_tl.pe_yaw.freqs"])
        return self.input_pe_encoder("model.tl_encoder", valid.reshape(-1).to(U8).contiguous(),
                                     state.to(torch.float32).reshape(-1, state.shape[-1]).contiguous(), pe)

    # ------------------------------------------------------------------ heads
    def latent_encoder(self, which: str, af: Var, av: Tensor, kv_map: List[Var], map_valid: Tensor, kv_tl: List[Var],
                       tl_valid: Tensor, S: int, T: int, A: int, P: int, TL: int) -> Tuple[Var, Tensor]:
        """`LatentEncoder.forward` after down-sampling (models/latent_encoder.py:104-147) + `DistEncoder` mean (:195-199).
        af [S*T*A, D] scene-major, av [S,T,A], kv_tl[i] [S*T*TL, 2D], tl_valid [S,T,TL].  Returns (mean [S*A, 16], valid)."""
        f = self.f
        avf = av.reshape(-1).to(U8).contiguous()
        x = self.tf_block("model.transformer_as2pl", 3, af, avf, kv_map, map_valid, S, T * A, P)
        x = self.tf_block("model.transformer_as2tl", 3, x, avf, kv_tl, tl_valid.reshape(S * T, TL).to(U8).contiguous(), S * T, A, TL)
        x = self.interaction(f"model.latent_encoder.agent_interaction_{which}", x, av.reshape(S * T, A).to(U8).contiguous(), S * T, A)
        # frame-major for the recurrent loop
        dev = av.device
        idx = (torch.arange(S, device=dev)[None, :, None] * T + torch.arange(T, device=dev)[:, None, None]) * A \
            + torch.arange(A, device=dev)[None, None, :]
        x_tm = f.gather_rows(x, idx.reshape(-1))
        frames = [f.row_slice(x_tm, t * S * A, (t + 1) * S * A) for t in range(T)]
        av_tm = av.transpose(0, 1).reshape(T, S * A).to(U8).contiguous()
        outs = self.gru_sequence(f"model.latent_encoder.agent_temporal_{which}", frames, av_tm)
        y = f.masked_max(f.cat_rows(outs), av_tm, 1, T, S * A, -1e3)  # TemporalAggregate max_valid (agent_temporal.py:31-44)
        v = av.any(1).reshape(-1).to(U8).contiguous()
        p = f"model.latent_encoder.latent_{which}_dist.mlp_mean.fc_layers"
        mean = self.lin(self.lin(y, f"{p}.0", relu=True), f"{p}.2")
        return f.add_mask(mean, None, v), v  # MLP.forward valid_mask (mlp.py:81-82)

    def dest_logits(self, af_hist: Var, av: Tensor, map_feature: Var, S: int, T: int, A: int, P: int) -> Var:
        """`DestPredictor.forward`, mode mlp, up to the raw logits (models/goal_manager.py:222-224,294-307).
        af_hist [S*T*A, D] scene-major (detached), av [S,T,A]."""
        f = self.f
        dev = av.device
        af_tm = Var(af_hist.data.view(S, T, A, D).transpose(0, 1).reshape(T * S * A, D).contiguous())  # detach_features
        frames = [f.row_slice(af_tm, t * S * A, (t + 1) * S * A) for t in range(T)]
        av_tm = av.transpose(0, 1).reshape(T, S * A).to(U8).contiguous()
        outs = self.gru_sequence("model.goal_manager.goal_predictor.gru_as", frames, av_tm)
        tgt = f.cat_rows([f.add_mask(o, fr, None) for o, fr in zip(outs, frames)])  # res_add_gru (:299-300)
        last = T - 1 - torch.max(av.flip(1).to(U8), dim=1)[1]  # TemporalAggregate last_valid (agent_temporal.py:33-36)
        idx = last.reshape(-1) * (S * A) + torch.arange(S * A, device=dev)
        tgt = f.add_mask(f.gather_rows(tgt, idx), None, av.any(1).reshape(-1).to(U8).contiguous())
        p = "model.goal_manager.goal_predictor.mlp.fc_layers"
        u = f.linear(map_feature.detach(), self.p(f"{p}.0.weight", cols=(0, D)), None)
        v = f.linear(tgt, self.p(f"{p}.0.weight", cols=(D, 2 * D)), self.p(f"{p}.0.bias"))
        x = f.pair_add(u, v, S, P, A)  # == Linear(cat[map_feature, tgt]) (:304-307)
        x = self.ln(x, f"{p}.1", relu=True)
        x = self.ln(self.lin(x, f"{p}.3"), f"{p}.4", relu=True)
        return self.lin(x, f"{p}.6")  # [S*A*P, 1]


def dest_masks(batch: Dict[str, Tensor], pl_valid: Tensor) -> Tensor:
    """pair_ok [S,A,P] = not masked to -inf by `DestPredictor.forward` (models/goal_manager.py:228-246,328-329)."""
    mt, at = batch["map/type"].bool(), batch["agent/type"].bool()
    type_ok = pl_valid.bool() & mt[:, :, :5].any(-1)
    m_veh = at[:, :, 0:1] & mt[:, :, 3].unsqueeze(1)
    m_ped = at[:, :, 1:2] & mt[:, :, :4].any(-1).unsqueeze(1)
    m_cyc = at[:, :, 2:3] & mt[:, :, :3].any(-1).unsqueeze(1)
    return (type_ok.unsqueeze(1) & ~(m_veh | m_ped | m_cyc)).to(U8).contiguous()


def teacher_forcing_mask(valid: Tensor, step_spawn_agent: int, step_warm_start: int) -> Tensor:
    """`TeacherForcing.get` without the (disabled) schedules (utils/teacher_forcing.py:43-56).  valid [B,T,A] bool."""
    m = torch.zeros_like(valid)
    m[:, 0] |= valid[:, 0]
    if step_spawn_agent > 0:
        spawn = (~valid[:, :-1]) & valid[:, 1:]
        spawn[:, step_spawn_agent:] = False
        m[:, 1:] |= spawn
    if step_warm_start >= 0:
        m[:, : step_warm_start + 1] |= valid[:, : step_warm_start + 1]
    return m


LOSS_CFG = dict(w_vae_kl=0.1, kl_free_nats=0.01, w_diffbar_reward=1.0, w_goal=1.0, step_training_start=10)


def training_forward(fn: Fn, params: Params, batch: Dict[str, Tensor], eps: Tensor, use_prior: bool, n_step: int = 90,
                     n_hist: int = 11, down: int = 5, loss_cfg: Dict = LOSS_CFG, return_buffers: bool = False,
                     drop_seed: Optional[Tensor] = None, drop_p: float = 0.0, defer_loss: bool = False,
                     first_drop_site: int = 0, stack_backward: Optional[bool] = None) -> Dict[str, Tensor]:
    """forward of `training_step` + seeding of the loss gradients; call `fn.backward()` afterwards.

    batch: the raw episode (`agent/*` [S,91,A,..], `tl_stop/*` [S,91,TL,..], `map/*`, `agent/dest`, ...) on the device of
    the parameters.  eps [S,A,16]: the standard-normal draw of `Normal.rsample` (distributions.py:30); use_prior: the outcome of
    `torch.rand(1) < p_training_rollout_prior` (:384-387).  drop_seed (device int32 [1]) / drop_p: dropout as in the reference's
    training mode (every nn.Dropout / nn.GRU dropout of the default config has p = 0.1); the masks come from a counter-based hash,
    so they differ from torch's generator (same distribution, not the same samples).  One simplification: the three aliased
    `encode_input_features` calls share one encoding, hence one set of dropout masks.  Returns the loss terms as device scalars."""
    m = Model(fn, params, drop_seed, drop_p)
    m.n_site = first_drop_site  # sub-batches of one step use disjoint site ranges: independent masks
    f, ops = fn, fn.ops
    dev = batch["agent/valid"].device
    gv = batch["agent/valid"].bool()
    S, T_gt, A = gv.shape
    P = batch["map/valid"].shape[1]
    TL = batch["tl_stop/valid"].shape[2]
    M = S * A
    H = slice(0, n_hist)

    # ---- encode_input_features (aliased inputs encoded once) ----
    map_feature, pl_valid = m.map_encoder(batch)
    hv = gv[:, H]
    af_hist = m.encode_agents(hv, batch["agent/pos"][:, H], batch["agent/yaw_bbox"][:, H], batch["agent/vel"][:, H],
                              batch["agent/spd"][:, H], batch["agent/yaw_rate"][:, H], batch["agent/acc"][:, H],
                              batch["agent/size"].unsqueeze(1).expand(-1, n_hist, -1, -1),
                              batch["agent/type"].unsqueeze(1).expand(-1, n_hist, -1, -1))  # [S*11*A, D] scene-major
    # traffic lights of the history in FRAME-major order: the rollout reads one frame per step
    tlv_tm = batch["tl_stop/valid"][:, H].transpose(0, 1).contiguous()  # [11,S,TL]
    tl_hist = m.encode_tl(tlv_tm, batch["tl_stop/state"][:, H].transpose(0, 1), batch["tl_stop/pos"][:, H].transpose(0, 1),
                          batch["tl_stop/dir"][:, H].transpose(0, 1))  # [11*S*TL, D]
    kv_map = [m.kv_project(f"model.transformer_as2pl.layers.{i}", map_feature) for i in range(3)]
    kv_tl_hist = [m.kv_project(f"model.transformer_as2tl.layers.{i}", tl_hist) for i in range(3)]

    # ---- destination predictor (pred_goal, :374-379) ----
    logits = m.dest_logits(af_hist, hv, map_feature, S, n_hist, A, P)
    goal_gt = batch["agent/dest"].long()
    goal_valid0 = hv.any(1)  # get_gt_goal (goal_manager.py:64-66) == DestCategorical.valid

    # ---- latent encoders (:382-383) ----
    Tp = len(range(0, n_hist, down))  # frames 0, 5, 10
    fr_p = torch.arange(0, n_hist, down, device=dev)
    idx_p = (torch.arange(S, device=dev)[:, None, None] * n_hist + fr_p[None, :, None]) * A \
        + torch.arange(A, device=dev)[None, None, :]
    af_prior = f.gather_rows(af_hist, idx_p.reshape(-1))
    idx_tl = (fr_p[None, :, None] * S + torch.arange(S, device=dev)[:, None, None]) * TL \
        + torch.arange(TL, device=dev)[None, None, :]
    kv_tl_prior = [f.gather_rows(kv, idx_tl.reshape(-1)) for kv in kv_tl_hist]
    prior_mean, prior_valid = m.latent_encoder("prior", af_prior, hv[:, ::down], kv_map, pl_valid, kv_tl_prior,
                                               batch["tl_stop/valid"][:, 0:n_hist:down], S, Tp, A, P, TL)
    Tq = len(range(0, T_gt, down))  # frames 0, 5, ..., 90
    sel = lambda k: batch[k][:, ::down]  # noqa: E731  (frames 0, 5, ..., 90)
    af_post = m.encode_agents(sel("agent/valid"), sel("agent/pos"), sel("agent/yaw_bbox"), sel("agent/vel"), sel("agent/spd"),
                              sel("agent/yaw_rate"), sel("agent/acc"), batch["agent/size"].unsqueeze(1).expand(-1, Tq, -1, -1),
                              batch["agent/type"].unsqueeze(1).expand(-1, Tq, -1, -1))
    tl_post = m.encode_tl(sel("tl_stop/valid"), sel("tl_stop/state"), sel("tl_stop/pos"), sel("tl_stop/dir"))  # [S*Tq*TL, D]
    kv_tl_post = [m.kv_project(f"model.transformer_as2tl.layers.{i}", tl_post) for i in range(3)]
    post_mean, post_valid = m.latent_encoder("post", af_post, gv[:, ::down], kv_map, pl_valid, kv_tl_post,
                                             sel("tl_stop/valid"), S, Tq, A, P, TL)
    ls_prior = params("model.latent_encoder.latent_prior_dist.log_std")
    ls_post = params("model.latent_encoder.latent_post_dist.log_std")

    # ---- rollout (:390-400 -> reactive_replay :420-476 -> rollout :205-354) ----
    z = f.rsample(prior_mean if use_prior else post_mean, ls_prior if use_prior else ls_post, eps.reshape(M, -1).contiguous())
    bidx = torch.arange(S, device=dev).unsqueeze(1)
    goal_feature = f.gather_rows(map_feature, (bidx * P + goal_gt).reshape(-1))  # goal_manager.py:131-138

    # without dropout relu(mlp_in(.)) of the goal feature / latent sample are loop invariants of the rollout and are evaluated
    # once; with dropout the reference draws fresh masks at every decode step, so they are evaluated per step (Model.decode_tail)
    hoist = m.dp() is None
    goal_src = m.mlp_in_goal(goal_feature) if hoist else goal_feature
    lat_src = m.mlp_in_latent(z) if hoist else z

    tf_mask = teacher_forcing_mask(gv, 10, 10)  # teacher_forcing_training (traffic_bots.yaml:131-137)
    gt_state = torch.cat([batch["agent/pos"], batch["agent/yaw_bbox"], batch["agent/spd"]], -1)  # [S,T,A,4]
    a_type = batch["agent/type"].reshape(M, 3).to(U8).contiguous()
    size = batch["agent/size"].reshape(M, 3)
    # TrafficRuleChecker.__init__ (utils/traffic_rule_checker.py:74-98)
    dest_valid = batch["map/valid"][bidx, goal_gt].to(U8).contiguous()
    dest_type = batch["map/type"][bidx, goal_gt]
    dest_pos = batch["map/pos"][bidx, goal_gt].contiguous()
    dest_dir = batch["map/dir"][bidx, goal_gt].contiguous()  # normalised in the kernel (:93)
    dest_lane = dest_type[:, :, :4].any(-1).to(U8).contiguous()
    dest_edge = dest_type[:, :, 4].to(U8).contiguous()
    boundary = batch["map/boundary"].contiguous()

    valid = gv[:, 0].to(U8).contiguous()  # Dynamics.init with frame 0 (:251-259)
    killed = torch.zeros_like(valid)
    dest_reached = torch.zeros_like(valid)
    goal_valid = goal_valid0.to(U8).contiguous()
    state = Var(gt_state[:, 0].reshape(M, 4).contiguous())
    vel, acc, yaw_rate = batch["agent/vel"][:, 0], batch["agent/acc"][:, 0], batch["agent/yaw_rate"][:, 0]
    zeros = Var(ops.zeros((M, D), gt_state))
    hidden = [zeros, zeros, zeros]
    rewards: List[Var] = []
    pred_valid_l, rv_l, preds_l = [], [], []
    zeros_u8 = torch.zeros_like(valid)
    f_xy = params.buffers["pre_processing.input.pose_pe_agent.pe_xy.freqs"]
    f_yaw = params.buffers["pre_processing.input.pose_pe_agent.pe_yaw.freqs"]

    # ---- step-stacked backward of the decode front (see Model.decode_front): the forward of every step writes its buffers into
    # row block t of step-stacked buffers; afterwards the chain is RECORDED once over the stacked buffers, so its backward is one
    # pass with n_step x M rows per launch instead of n_step passes with M rows (the front is ~70 % of a step's launches).
    # Needs a back end with an allocation hook and the general attention kernel for the shared map keys (P > 32).
    if stack_backward is None:
        stack_backward = hasattr(ops, "alloc_hook") and P > 32
    pre_site = first_drop_site + 500000  # dropout sites of the front: the same ids at every step, element indices continue
    tail_site = first_drop_site + 700000  # ... and of the tail
    if stack_backward:
        stack = StepStack(ops, n_step)
        fn_s = Fn(ops, record=False)
        m_s = Model(fn_s, params, drop_seed, drop_p)
        fr = torch.clamp(torch.arange(n_step, device=dev), max=n_hist - 1)  # TL frame of step t: min(t - 1, 10) (:287)
        idx_steps = ((fr[:, None, None] * S + torch.arange(S, device=dev)[None, :, None]) * TL
                     + torch.arange(TL, device=dev)[None, None, :]).reshape(-1)
        kv_tl_steps = [f.gather_rows(kv, idx_steps) for kv in kv_tl_hist]  # [n_step * S * TL, 2D] (recorded: scatter-add backward)
        tlv_steps = tlv_tm[fr].to(U8).contiguous()  # [n_step, S, TL]
        # the same for the tail of a step (Model.decode_tail): its inputs repeated per step (recorded: scatter-add backward)
        stack_t = StepStack(ops, n_step)
        fn_t = Fn(ops, record=False)
        m_t = Model(fn_t, params, drop_seed, drop_p)
        idx_rep = torch.arange(M, device=dev).repeat(n_step)
        goal_rep, lat_rep = f.gather_rows(goal_src, idx_rep), f.gather_rows(lat_src, idx_rep)
        rollout_start = len(f.nodes)
        attr_l, pe_l, vflat_l, valid_l, gvalid_l, x_gru_l = [], [], [], [], [], []
        x_front = mean_all = None
        main_nodes, gru_nodes, dyn_nodes = f.nodes, [], []

    for t in range(1, n_step + 1):
        ovr = tf_mask[:, t].to(U8) if t < T_gt else zeros_u8
        tl_t = min(t - 1, n_hist - 1)
        vflat = valid.reshape(-1)
        # forward(): state embedding from the DETACHED state (:136-153; stale vel / acc / yaw_rate: SURVEY 8a a3)
        sd_ = state.data.view(S, A, 4)
        attr = torch.cat([vel, sd_[..., 3:4], yaw_rate, acc, size.view(S, A, 3), a_type.view(S, A, 3).to(torch.float32)], -1)
        pe = ops.pose_pe(sd_[..., :2].reshape(M, 2).contiguous(), sd_[..., 2].reshape(M).contiguous(), f_xy, f_yaw)
        attr = attr.reshape(M, 11).contiguous()
        # TrafficBots.forward (models/traffic_bots.py:201-241)
        if stack_backward:
            i = t - 1
            post_site, m_s.n_site, fn_s.stack_t = m.n_site, pre_site, i
            stack.begin(i)
            kv_t = [Var(kv.data[i * S * TL:(i + 1) * S * TL]) for kv in kv_tl_steps]
            x = m_s.decode_front(attr, pe, vflat, valid, kv_map, pl_valid, kv_t, tlv_steps[i], S, A, P, TL)
            stack.end()
            if x_front is None:
                x_front = Var(stack.bufs[-1].flatten(0, 1), True)  # the front's outputs of all steps
            assert x.data.data_ptr() == x_front.data[i * M:(i + 1) * M].data_ptr()
            x = f.row_slice(x_front, i * M, (i + 1) * M)
            attr_l.append(attr), pe_l.append(pe), vflat_l.append(vflat), valid_l.append(valid)
            m.n_site = post_site
        else:  # per-step recording; the dropout sites / element indices are those of the stacked schedule (same masks)
            kv_t = [f.row_slice(kv, tl_t * S * TL, (tl_t + 1) * S * TL) for kv in kv_tl_hist]
            post_site, m.n_site, f.stack_t = m.n_site, pre_site, t - 1
            x = m.decode_front(attr, pe, vflat, valid, kv_map, pl_valid, kv_t, tlv_tm[tl_t].to(U8).contiguous(), S, A, P, TL)
            m.n_site, f.stack_t = post_site, None
        if stack_backward:
            f.nodes = gru_nodes  # back-propagation through time: the GRU nodes of all steps, kept apart from ...
        x, hidden = m.gru_layers("model.agent_temporal", x, hidden, vflat)
        gflat = goal_valid.reshape(-1)
        if stack_backward:
            i = t - 1
            x_gru_l.append(x), gvalid_l.append(gflat)
            post_site, m_t.n_site, fn_t.stack_t = m.n_site, tail_site, i
            stack_t.begin(i)
            mean = m_t.decode_tail(Var(x.data), Var(goal_src.data), Var(lat_src.data), gflat, vflat, a_type, hoist)
            stack_t.end()
            if mean_all is None:
                mean_all = Var(stack_t.bufs[-1].flatten(0, 1), True)  # the action means of all steps
            assert mean.data.data_ptr() == mean_all.data[i * M:(i + 1) * M].data_ptr()
            mean = f.row_slice(mean_all, i * M, (i + 1) * M)
            f.nodes = dyn_nodes  # ... the dynamics / reward nodes of all steps
        else:
            post_site, m.n_site, f.stack_t = m.n_site, tail_site, t - 1
            mean = m.decode_tail(x, goal_src, lat_src, gflat, vflat, a_type, hoist)
            m.n_site, f.stack_t = post_site, None
        # Dynamics.update / override_states (utils/dynamics.py:74-149)
        pred = f.dynamics(state, mean, a_type, vflat)
        pred_valid = valid
        mo = ovr & (killed ^ 1)
        valid = valid | mo
        if t < T_gt:
            mm = mo.bool().unsqueeze(-1)
            state = f.select_rows(mo.reshape(-1).contiguous(), Var(gt_state[:, t].reshape(M, 4).contiguous()), pred)
            vel = torch.where(mm, batch["agent/vel"][:, t], vel)
            acc = torch.where(mm, batch["agent/acc"][:, t], acc)
            yaw_rate = torch.where(mm, batch["agent/yaw_rate"][:, t], yaw_rate)
        else:
            state = pred
        # rule check on the post-override state, kill, disable_goal_reached (:311-320)
        valid, killed, dest_reached, goal_valid = ops.sim_flags(
            state.data.view(S, A, 4), valid, gv[:, t].to(U8).contiguous() if t < T_gt else None, boundary, dest_pos, dest_dir,
            dest_valid, dest_lane, dest_edge, killed, dest_reached, goal_valid)
        # DifferentiableReward.get, IL part (utils/rewards.py:117-131)
        if t < T_gt:
            rv = (pred_valid & gv[:, t].to(U8)).reshape(-1).contiguous()
            rewards.append(f.reward(pred, gt_state[:, t].reshape(M, 4).contiguous(), rv))
        else:
            rv = pred_valid.reshape(-1)
            rewards.append(Var(ops.zeros((M, 1), gt_state)))
        pred_valid_l.append(pred_valid.reshape(-1))
        rv_l.append(rv)
        if return_buffers:
            preds_l.append(pred.data)

    if stack_backward:
        # record front and tail ONCE over the stacked buffers.  Order of the backward (= reverse of the node list): dynamics /
        # reward chain of all steps (fills the gradient of every step's action mean), tail (one pass), GRU chain of all steps
        # (back-propagation through time; fills the gradient of every step's front output), front (one pass), everything else.
        rops = ReplayOps(ops, stack)
        fn_r = Fn(rops)
        m_r = Model(fn_r, params, drop_seed, drop_p)
        m_r.n_site = pre_site
        vflat_all = torch.cat(vflat_l)
        out_r = m_r.decode_front(torch.cat(attr_l), torch.cat(pe_l), vflat_all, torch.cat(valid_l), kv_map, pl_valid,
                                 kv_tl_steps, tlv_steps.reshape(n_step * S, TL), n_step * S, A, P, TL, kv_shared=True)
        assert rops.done() and out_r.data.data_ptr() == x_front.data.data_ptr()
        fn_r.ops = ops
        out_r.grad, out_r.grad_fixed = x_front.grad, True  # filled by the GRU backward of every step (row slices)
        f.nodes = cat_nodes = []
        x_gru_all = f.cat_rows(x_gru_l)
        rops_t = ReplayOps(ops, stack_t)
        fn_rt = Fn(rops_t)
        m_rt = Model(fn_rt, params, drop_seed, drop_p)
        m_rt.n_site = tail_site
        out_t = m_rt.decode_tail(x_gru_all, goal_rep, lat_rep, torch.cat(gvalid_l), vflat_all, a_type.repeat(n_step, 1), hoist)
        assert rops_t.done() and out_t.data.data_ptr() == mean_all.data.data_ptr()
        fn_rt.ops = ops
        out_t.grad, out_t.grad_fixed = mean_all.grad, True  # filled by the dynamics backward of every step
        assert len(main_nodes) == rollout_start
        f.nodes = main_nodes + fn_r.nodes + gru_nodes + cat_nodes + fn_rt.nodes + dyn_nodes
        f.n_fwd += fn_s.n_fwd + fn_t.n_fwd

    # ---- TrainingMetrics.update / compute (models/metrics/training.py:62-158) ----
    t0 = loss_cfg["step_training_start"]
    pv = torch.stack(pred_valid_l, 1).bool()  # [M, n_step]
    pv[:, :t0] = False
    rvs = torch.stack(rv_l, 1).bool() & pv
    any_pv = pv.any(-1)
    kl_valid = (post_valid.bool() & any_pv)
    goal_rows = (goal_valid0.reshape(-1) & any_pv)
    rvs_u8 = rvs.to(U8).contiguous()
    # the three normalisers of TrainingMetrics.compute (:150-157): counts over the WHOLE batch
    counts = torch.stack([rvs.sum(), kl_valid.sum(), goal_rows.sum()]).to(torch.float32)

    def finish(counts_total: Tensor) -> Dict[str, Tensor]:
        """seeds the loss gradients with the given batch-wide counts [reward, kl, goal] (device tensor; no host synchronisation)
        and returns this (sub-)batch's share of the loss terms; `fn.backward()` follows."""
        s_r = (loss_cfg["w_diffbar_reward"] / counts_total[0]).reshape(1)
        r_all = torch.cat([r.data for r in rewards], 1)  # [M, n_step]
        loss_r = -s_r * ops.masked_sum(r_all, rvs_u8)  # w * (-sum r) / count
        for i, r in enumerate(rewards):
            if r.req:
                r.grad = ops.mask_scale(rvs_u8[:, i].contiguous(), -s_r)  # d loss / d r on the counted entries
        s_kl = (loss_cfg["w_vae_kl"] / counts_total[1]).reshape(1)
        kl_sum, dmq, dmp = ops.kl_fwd_bwd(post_mean.data, ls_post.data.view(-1), prior_mean.data, ls_prior.data.view(-1),
                                          kl_valid.to(U8).contiguous(), loss_cfg["kl_free_nats"], s_kl, ls_post.grad.view(-1),
                                          ls_prior.grad.view(-1))
        loss_kl = s_kl * kl_sum
        f._acc(post_mean, dmq)
        f._acc(prior_mean, dmp)
        s_g = (loss_cfg["w_goal"] / counts_total[2]).reshape(1)
        nll_sum, dlogits = ops.dest_nll(logits.data.view(S, A, P), dest_masks(batch, pl_valid), goal_valid0.to(U8).contiguous(),
                                        goal_gt, goal_rows.view(S, A).to(U8).contiguous(), s_g)
        loss_g = s_g * nll_sum
        f._acc(logits, dlogits.reshape(-1, 1))
        out = {"loss": loss_kl + loss_r + loss_g, "vae_kl": loss_kl, "diffbar_reward": loss_r, "goal_loss": loss_g}
        if return_buffers:
            out["preds"] = torch.stack(preds_l, 1).view(S, A, n_step, 4)
            out["pred_valid"] = torch.stack(pred_valid_l, 1).view(S, A, n_step)
            out["post_mean"] = post_mean.data.view(S, A, -1)
            out["prior_mean"] = prior_mean.data.view(S, A, -1)
            out["map_feature"] = map_feature.data.view(S, P, D)
        return out

    if defer_loss:  # the caller sums `counts` over its sub-batches first (TrainState.forward_backward with n_split > 1)
        return {"counts": counts, "finish": finish}
    return finish(counts)

"""Host-side driver of the CUDA hot path: owns the packed parameters and the device buffers, and exposes the two
operations the reference's training / eval loop needs -- scene encoding and the closed-loop rollout -- on torch
CUDA tensors in the reference's batch schema.  All arithmetic happens in `libtrafficbots_b200.so`; torch is used
for device memory and streams only.

Reference methods replaced (paths under the reference's `src/`):
  `Engine.encode_scene`  -> `SceneCentricInput.forward` + `TrafficBots.encode_input_features`
                            (data_modules/sc_input.py:98-140, models/traffic_bots.py:109-151)
  `Engine.rollout`       -> `WaymoMotion.rollout` with `WaymoMotion.forward` as the loop body
                            (pl_modules/waymo_motion.py:108-354)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import torch
from torch import Tensor

from . import _native as nt

VIOLATION_KEYS = ("outside_map", "outside_map_this_step", "goal_reached", "goal_reached_this_step", "dest_reached",
                  "dest_reached_this_step")
# the four checks that are off in the default config report their (all-False) sticky state (traffic_rule_checker.py:426-472)
DISABLED_VIOLATION_KEYS = ("collided", "collided_this_step", "run_road_edge", "run_road_edge_this_step",
                           "run_red_light", "run_red_light_this_step", "passive", "passive_this_step")


class SceneFeatures(dict):
    """dict returned by `encode_scene`: the reference's feature dict (`agent_feature(_valid)`, `map_feature(_valid)`,
    `tl_feature(_valid)`) plus the projected K|V caches under private keys (`_kv_map`, `_kv_tl`)."""


class Engine:
    def __init__(self, state_dict: Optional[Mapping[str, Tensor]], device: Optional[torch.device] = None,
                 packed: Optional[Tensor] = None, rollout_cluster: int = 0):
        """`packed`: share the packed parameter blob of another Engine (see `fork`); otherwise `state_dict` is packed.
        `rollout_cluster`: CTAs per scene-mode of the persistent decode kernel (`TbDims.n_cta_per_mode`; 0 = chosen by the
        library from the batch size, 1 when several batches are kept in flight on separate streams)."""
        self.lib = nt.lib()
        if not torch.cuda.is_available():
            raise nt.TbError("no CUDA device: trafficbots_b200 has no CPU implementation of the hot path")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rollout_cluster = int(rollout_cluster)
        if packed is not None:
            self.packed = packed
        else:
            self.packed = torch.empty(self.lib.tb_packed_weight_bytes() // 4, dtype=torch.float32, device=self.device)
            self.load_state_dict(state_dict)
        self._enc_ws: Optional[Tensor] = None
        self._state: Optional[Tensor] = None
        self._state_dims = None
        self.last_t = 0

    def fork(self, rollout_cluster: Optional[int] = None) -> "Engine":
        """a second Engine on the SAME packed parameters with its own workspaces / simulation state: one per batch kept in
        flight (`pipeline.ScenePipeline`).  Re-packing through either engine updates both."""
        return Engine(None, self.device, packed=self.packed,
                      rollout_cluster=self.rollout_cluster if rollout_cluster is None else rollout_cluster)

    # ------------------------------------------------------------------------------------------------ parameters
    def load_state_dict(self, state_dict: Mapping[str, Tensor]) -> None:
        """re-lays a reference `WaymoMotion.state_dict()` (fp32) into the kernel layout (`tb_pack_weights`)."""
        n = self.lib.tb_weight_count()
        keep = []
        ptrs = (C.c_void_p * n)()
        for i in range(n):
            name = self.lib.tb_weight_name(i).decode()
            rows, cols = self.lib.tb_weight_rows(i), self.lib.tb_weight_cols(i)
            if name not in state_dict:
                raise nt.TbError(f"state_dict is missing {name}")
            t = state_dict[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            want = (rows,) if cols == 0 else (rows, cols)
            if tuple(t.shape) != want:
                raise nt.TbError(f"{name}: shape {tuple(t.shape)}, expected {want}")
            keep.append(t)
            ptrs[i] = t.data_ptr()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)  # forked engines may still be reading the blob on other streams
            nt.check(self.lib.tb_pack_weights(ptrs, self.packed.data_ptr(), nt.current_stream_ptr()), "tb_pack_weights")
            # `keep` may be freed after this, and forked engines on other streams must see the new blob
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------------ encoding
    def _dims(self, S, K, A, P, TL, Th, Tg, T) -> nt.TbDims:
        return nt.TbDims(S, K, A, P, TL, Th, Tg, T, self.rollout_cluster)

    def encode_scene(self, batch: Mapping[str, Tensor], prefix: str = "history/", share_map: Optional[Mapping[str, Tensor]] = None
                     ) -> SceneFeatures:
        """batch: reference batch dict (CUDA tensors). Uses map/* and `{prefix}agent/*`, `{prefix}tl_stop/*`.
        `share_map`: features of an earlier `encode_scene` of the SAME scenes: the map is not encoded again (its features
        and K|V caches are taken from there), only the agent / traffic-light tensors under `prefix` are."""
        mv = batch["map/valid"]
        S, P, N = mv.shape
        if N != 20:
            raise nt.TbError("n_pl_node must be 20")
        if prefix == "sc/":  # the re-keyed history tensors of SceneCentricPreProcessing (`sc/agent_*`, `sc/tl_*`)
            ak, tk = (lambda k: f"sc/agent_{k}"), (lambda k: f"sc/tl_{k}")
        else:
            ak, tk = (lambda k: f"{prefix}agent/{k}"), (lambda k: f"{prefix}tl_stop/{k}")
        av = batch[ak("valid")]
        Th, A = av.shape[1], av.shape[2]
        TL = batch[tk("valid")].shape[2]
        dims = self._dims(S, 1, A, P, TL, Th, Th, 1)
        g = lambda k: batch[k]  # noqa: E731
        p = nt.dev_ptr
        sin = nt.TbSceneIn(
            p(mv, "u8", (S, P, 20), "map/valid") if share_map is None else None,
            p(g("map/type"), "u8", (S, P, 11), "map/type"),
            p(g("map/pos"), "f32", (S, P, 20, 2), "map/pos"), p(g("map/dir"), "f32", (S, P, 20, 2), "map/dir"),
            p(av, "u8", (S, Th, A), ak("valid")), p(g(ak("pos")), "f32", (S, Th, A, 2), ak("pos")),
            p(g(ak("yaw_bbox")), "f32", (S, Th, A, 1), ak("yaw_bbox")),
            p(g(ak("vel")), "f32", (S, Th, A, 2), ak("vel")), p(g(ak("spd")), "f32", (S, Th, A, 1), ak("spd")),
            p(g(ak("yaw_rate")), "f32", (S, Th, A, 1), ak("yaw_rate")),
            p(g(ak("acc")), "f32", (S, Th, A, 1), ak("acc")), p(g(ak("size")), "f32", (S, A, 3), ak("size")),
            p(g(ak("type")), "u8", (S, A, 3), ak("type")), p(g(tk("valid")), "u8", (S, Th, TL), tk("valid")),
            p(g(tk("state")), "u8", (S, Th, TL, 5), tk("state")),
            p(g(tk("pos")), "f32", (S, Th, TL, 2), tk("pos")), p(g(tk("dir")), "f32", (S, Th, TL, 2), tk("dir")))
        dev = self.device
        f = SceneFeatures()
        map_keys = ("map_feature", "map_feature_valid", "_kv_map", "_kv_map_tc", "_n_key_map")
        if share_map is not None:
            for k in map_keys:
                f[k] = share_map[k]
            if tuple(f["map_feature"].shape) != (S, P, 128):
                raise nt.TbError("encode_scene: share_map comes from a different batch shape")
        else:
            f["map_feature"] = torch.empty(S, P, 128, device=dev)
            f["map_feature_valid"] = torch.empty(S, P, dtype=torch.bool, device=dev)
            f["_kv_map"] = torch.empty(3, S, P, 256, device=dev)
            f["_kv_map_tc"] = torch.empty(self.lib.tb_kv_tc_bytes(C.byref(dims), 0), dtype=torch.uint8, device=dev)
            f["_n_key_map"] = torch.empty(S, dtype=torch.int32, device=dev)
        f["agent_feature"] = torch.empty(S, Th, A, 128, device=dev)
        f["agent_feature_valid"] = av
        f["tl_feature"] = torch.empty(S, Th, TL, 128, device=dev)
        f["tl_feature_valid"] = batch[tk("valid")]
        f["_kv_tl"] = torch.empty(3, S, Th, TL, 256, device=dev)
        f["_kv_tl_tc"] = torch.empty(self.lib.tb_kv_tc_bytes(C.byref(dims), 1), dtype=torch.uint8, device=dev)
        f["_n_key_tl"] = torch.empty(S, Th, dtype=torch.int32, device=dev)
        sout = nt.TbSceneOut(f["map_feature"].data_ptr(), f["map_feature_valid"].data_ptr(), f["agent_feature"].data_ptr(),
                             f["tl_feature"].data_ptr(), f["_kv_map"].data_ptr(), f["_kv_tl"].data_ptr(),
                             f["_kv_map_tc"].data_ptr(), f["_kv_tl_tc"].data_ptr(), f["_n_key_map"].data_ptr(),
                             f["_n_key_tl"].data_ptr())
        need = self.lib.tb_encode_workspace_bytes(C.byref(dims))
        if self._enc_ws is None or self._enc_ws.numel() < need:
            self._enc_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            nt.check(self.lib.tb_encode_scene(C.byref(dims), C.byref(sin), self.packed.data_ptr(), C.byref(sout),
                                              self._enc_ws.data_ptr(), nt.current_stream_ptr()), "tb_encode_scene")
        return f

    # ------------------------------------------------------------------------------------------------ building blocks
    def kv_project(self, block: int, layer: int, tgt: Tensor) -> Tensor:
        rows = tgt.numel() // 128
        kv = torch.empty(*tgt.shape[:-1], 256, device=self.device)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_kv_project(block, layer, nt.dev_ptr(tgt, "f32", name="tgt"), rows, self.packed.data_ptr(),
                                            kv.data_ptr(), nt.current_stream_ptr()), "tb_kv_project")
        return kv

    def xlayer(self, block: int, layer: int, src: Tensor, src_valid: Tensor, kv: Tensor, key_valid: Tensor,
               kv_share: int = 1, mask_self: bool = False) -> Tensor:
        nb, ns, _ = src.shape
        nk = kv.shape[-2]
        dst = torch.empty_like(src)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_xlayer(block, layer, nt.dev_ptr(src, "f32", name="src"),
                                        nt.dev_ptr(src_valid, "u8", (nb, ns), "src_valid"), nb, ns,
                                        nt.dev_ptr(kv, "f32", (nb // kv_share, nk, 256), "kv"),
                                        nt.dev_ptr(key_valid, "u8", (nb // kv_share, nk), "key_valid"), nk, kv_share,
                                        int(mask_self), self.packed.data_ptr(), dst.data_ptr(), nt.current_stream_ptr()),
                     "tb_xlayer")
        return dst

    def xlayer_tc(self, block: int, layer: int, src: Tensor, src_valid: Tensor, key_blocks: Tensor, n_key: Tensor,
                  n_key_max: int, kv_share: int = 1) -> Tensor:
        """`tb_xlayer_tc`: the layer on the tensor pipe against compacted key blocks.  src [nb, ns, 128]; key_blocks uint8
        [nb / kv_share, ceil(n_key_max / 64), 65536]; n_key int32 [nb / kv_share]."""
        nb, ns, _ = src.shape
        nT = (n_key_max + 63) // 64
        dst = torch.empty_like(src)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_xlayer_tc(block, layer, nt.dev_ptr(src, "f32", name="src"),
                                           nt.dev_ptr(src_valid, "u8", (nb, ns), "src_valid"), nb, ns,
                                           nt.dev_ptr(key_blocks, "u8", (nb // kv_share, nT, 65536), "key_blocks"),
                                           nt.dev_ptr(n_key, "i32", (nb // kv_share,), "n_key"), n_key_max, kv_share,
                                           self.packed.data_ptr(), dst.data_ptr(), nt.current_stream_ptr()), "tb_xlayer_tc")
        return dst

    # ------------------------------------------------------------------------------------------------ pre-rollout heads
    def gru_sequence(self, which: int, mode: int, x: Tensor, valid: Tensor, t_stride: int = 1):
        """`MultiAgentGRULoop` over the frames + temporal aggregation (`tb_gru_sequence`).  x [B,T,A,128], valid [B,T,A]
        -> (out [B,A,128], out_valid [B,A])."""
        B, T, A, _ = x.shape
        out = torch.empty(B, A, 128, device=self.device)
        ov = torch.empty(B, A, dtype=torch.bool, device=self.device)
        ws = torch.empty(self.lib.tb_gru_workspace_bytes(B, A), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_gru_sequence(which, mode, nt.dev_ptr(x, "f32", (B, T, A, 128), "x"),
                                              nt.dev_ptr(valid, "u8", (B, T, A), "valid"), B, T, A, t_stride,
                                              self.packed.data_ptr(), ws.data_ptr(), out.data_ptr(), ov.data_ptr(),
                                              nt.current_stream_ptr()), "tb_gru_sequence")
        return out, ov

    def latent_encoder(self, feat: Mapping[str, Tensor], posterior: bool = False, temporal_down_sample_rate: int = 5):
        """`LatentEncoder.forward` (models/latent_encoder.py:95-147): every `temporal_down_sample_rate`-th frame of the
        agent features through the policy's agent->map / agent->traffic-light blocks (shared weights), the latent encoder's
        own interaction block and GRU, max over the valid frames, and the mean MLP.  Returns (mean [S,A,16], valid [S,A]);
        the std is the constant `exp(log_std)` parameter.  `feat`: the dict of `encode_scene` (posterior: encoded with the
        full-horizon tensors)."""
        d = temporal_down_sample_rate
        av = feat["agent_feature_valid"][:, ::d].contiguous()
        x = feat["agent_feature"][:, ::d].contiguous()
        tv = feat["tl_feature_valid"][:, ::d].contiguous()
        S, T, A, _ = x.shape
        TL = tv.shape[2]
        kv_tl = feat["_kv_tl"][:, :, ::d].contiguous()  # [3,S,T,TL,256]
        x = x.view(S, T * A, 128)
        avf = av.view(S, T * A)
        avb = av.view(S * T, A)
        if "_kv_map_tc" in feat:  # tensor-core layers on the compacted key blocks of encode_scene
            P = feat["map_feature"].shape[1]
            nT_map, nT_tl = (P + 63) // 64, (TL + 63) // 64
            Th = feat["tl_feature_valid"].shape[1]
            kmap = feat["_kv_map_tc"].view(3, S, nT_map, 65536)
            ktl = feat["_kv_tl_tc"].view(3, S, Th, nT_tl, 65536)[:, :, ::d].contiguous()
            nk_tl = feat["_n_key_tl"][:, ::d].contiguous().view(S * T)
            for L in range(3):
                x = self.xlayer_tc(nt.BLOCK_AS2PL, L, x, avf, kmap[L], feat["_n_key_map"], P)
            x = x.view(S * T, A, 128)
            for L in range(3):
                x = self.xlayer_tc(nt.BLOCK_AS2TL, L, x, avb, ktl[L].reshape(S * T, nT_tl, 65536), nk_tl, TL)
        else:
            for L in range(3):
                x = self.xlayer(nt.BLOCK_AS2PL, L, x, avf, feat["_kv_map"][L], feat["map_feature_valid"])
            x = x.view(S * T, A, 128)
            for L in range(3):
                x = self.xlayer(nt.BLOCK_AS2TL, L, x, avb, kv_tl[L].reshape(S * T, TL, 256), tv.view(S * T, TL))
        block = nt.BLOCK_LATENT_POST_INT if posterior else nt.BLOCK_LATENT_PRIOR_INT
        x0 = x
        for L in range(3):  # tgt = the block input for all layers (agent_interaction.py:52), eye mask
            kv = self.kv_project(block, L, x0)
            x = self.xlayer(block, L, x, avb, kv, avb, mask_self=True)
        single = avb.sum(-1, keepdim=True) == 1  # scenes with exactly one valid agent bypass the block (:61-77)
        x = torch.where(single.unsqueeze(-1), x0, x)
        agg, v = self.gru_sequence(nt.GRU_LATENT_POST if posterior else nt.GRU_LATENT_PRIOR, 0, x.view(S, T, A, 128), av)
        mean = torch.empty(S, A, 16, device=self.device)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_mlp_head(nt.MLP_LATENT_POST_MEAN if posterior else nt.MLP_LATENT_PRIOR_MEAN, agg.data_ptr(),
                                          v.data_ptr(), S * A, self.packed.data_ptr(), mean.data_ptr(), nt.current_stream_ptr()),
                     "tb_mlp_head")
        return mean, v

    def dest_predictor(self, feat: Mapping[str, Tensor], agent_type: Tensor, map_type: Tensor):
        """`DestPredictor.forward`, mode mlp (models/goal_manager.py:228-246,294-333).  Returns (probs [S,A,P],
        logp [S,A,P], valid [S,A])."""
        af, av = feat["agent_feature"], feat["agent_feature_valid"]
        S, T, A, _ = af.shape
        P = feat["map_feature"].shape[1]
        tgt, v = self.gru_sequence(nt.GRU_DEST, 1, af, av)
        need = self.lib.tb_dest_workspace_bytes(S, A, P)
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        logp = torch.empty(S, A, P, device=self.device)
        probs = torch.empty(S, A, P, device=self.device)
        p = nt.dev_ptr
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_dest_logits(S, A, P, p(feat["map_feature"], "f32", (S, P, 128), "map_feature"),
                                             p(feat["map_feature_valid"], "u8", (S, P), "map_feature_valid"),
                                             p(map_type, "u8", (S, P, 11), "map/type"), tgt.data_ptr(), v.data_ptr(),
                                             p(agent_type, "u8", (S, A, 3), "agent/type"), self.packed.data_ptr(), ws.data_ptr(),
                                             logp.data_ptr(), probs.data_ptr(), nt.current_stream_ptr()), "tb_dest_logits")
        return probs, logp, v

    # ------------------------------------------------------------------------------------------------ rollout
    def _rollout_structs(self, feat, gt, tf_mask, agent_type, agent_size, raw_map, latent_sample, latent_logp, dest,
                         goal_valid, goal_gt, n_mode, n_step):
        S, P, _ = feat["map_feature"].shape
        Th, TL = feat["tl_feature_valid"].shape[1:]
        Tg, A = gt["valid"].shape[1:]
        B = S * n_mode
        dims = self._dims(S, n_mode, A, P, TL, Th, Tg, n_step)
        p = nt.dev_ptr
        rin = nt.TbRolloutIn(
            p(feat["map_feature"], "f32", (S, P, 128), "map_feature"), p(feat["map_feature_valid"], "u8", (S, P), "map_feature_valid"),
            p(feat["_kv_map"], "f32", (3, S, P, 256), "_kv_map"), p(feat["_kv_tl"], "f32", (3, S, Th, TL, 256), "_kv_tl"),
            p(feat["tl_feature_valid"], "u8", (S, Th, TL), "tl_feature_valid"),
            p(gt["valid"], "u8", (S, Tg, A), "gt valid"), p(gt["pos"], "f32", (S, Tg, A, 2), "gt pos"),
            p(gt["yaw_bbox"], "f32", (S, Tg, A, 1), "gt yaw_bbox"), p(gt["spd"], "f32", (S, Tg, A, 1), "gt spd"),
            p(gt["vel"], "f32", (S, Tg, A, 2), "gt vel"), p(gt["acc"], "f32", (S, Tg, A, 1), "gt acc"),
            p(gt["yaw_rate"], "f32", (S, Tg, A, 1), "gt yaw_rate"), p(tf_mask, "u8", (S, Tg, A), "tf_mask"),
            p(agent_type, "u8", (S, A, 3), "agent_type"), p(agent_size, "f32", (S, A, 3), "agent_size"),
            p(raw_map["boundary"], "f32", (S, 4), "map/boundary"), p(raw_map["valid"], "u8", (S, P, 20), "map/valid"),
            p(raw_map["type"], "u8", (S, P, 11), "map/type"), p(raw_map["pos"], "f32", (S, P, 20, 2), "map/pos"),
            p(raw_map["dir"], "f32", (S, P, 20, 2), "map/dir"), p(goal_gt, "f32", (S, A, 4), "goal_gt", optional=True),
            p(latent_sample, "f32", (B, A, 16), "latent_sample"), p(latent_logp, "f32", (B, A), "latent_logp"),
            p(dest, "i64", (B, A), "dest"), p(goal_valid, "u8", (B, A), "goal_valid"),
            feat["_kv_map_tc"].data_ptr() if "_kv_map_tc" in feat else None,
            feat["_kv_tl_tc"].data_ptr() if "_kv_tl_tc" in feat else None,
            feat["_n_key_map"].data_ptr() if "_n_key_map" in feat else None,
            feat["_n_key_tl"].data_ptr() if "_n_key_tl" in feat else None)
        return dims, rin

    def _ensure_state(self, dims: nt.TbDims) -> Tensor:
        need = self.lib.tb_rollout_state_bytes(C.byref(dims))
        if need == 0:
            raise nt.TbError("tb_rollout_state_bytes: unsupported dimensions")
        if self._state is None or self._state.numel() < need:
            self._state = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._state_dims = dims
        return self._state

    def alloc_outputs(self, B: int, A: int, T: int, trace: bool = False) -> Dict[str, Tensor]:
        dev = self.device
        o = {
            "preds": torch.empty(B, A, T, 4, device=dev), "valid": torch.empty(B, A, T, dtype=torch.bool, device=dev),
            "override_masks": torch.empty(B, A, T, dtype=torch.bool, device=dev),
            "diffbar_rewards": torch.empty(B, A, T, device=dev),
            "diffbar_rewards_valid": torch.empty(B, A, T, dtype=torch.bool, device=dev),
            "action_log_probs": torch.empty(B, A, T, device=dev), "latent_log_probs": torch.empty(B, A, T, device=dev),
            "_violations": torch.empty(6, B, A, T, dtype=torch.bool, device=dev),
        }
        if trace:
            o["trace/policy_feature"] = torch.empty(B, A, T, 128, device=dev)
            o["trace/action_mean"] = torch.empty(B, A, T, 2, device=dev)
        return o

    @staticmethod
    def _out_struct(o: Dict[str, Tensor]) -> nt.TbRolloutOut:
        tp = o.get("trace/policy_feature")
        ta = o.get("trace/action_mean")
        return nt.TbRolloutOut(o["preds"].data_ptr(), o["valid"].data_ptr(), o["override_masks"].data_ptr(),
                               o["diffbar_rewards"].data_ptr(), o["diffbar_rewards_valid"].data_ptr(),
                               o["action_log_probs"].data_ptr(), o["latent_log_probs"].data_ptr(),
                               o["_violations"].data_ptr(), tp.data_ptr() if tp is not None else None,
                               ta.data_ptr() if ta is not None else None)

    def rollout(self, feat: Mapping[str, Tensor], gt: Mapping[str, Tensor], tf_mask: Tensor, agent_type: Tensor,
                agent_size: Tensor, raw_map: Mapping[str, Tensor], latent_sample: Tensor, latent_logp: Tensor,
                dest: Tensor, goal_valid: Tensor, goal_gt: Optional[Tensor], n_mode: int = 1, n_step: int = 90,
                trace: bool = False, out: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
        """Full closed-loop rollout (t = 1..n_step).  Per-scene tensors have leading dim S, per-scene-mode tensors
        (`latent_sample`, `latent_logp`, `dest`, `goal_valid`) have leading dim B = S * n_mode (scene-major).
        gt: dict with valid/pos/yaw_bbox/spd/vel/acc/yaw_rate, [S,Tg,A,.].  Returns the RolloutBuffer fields,
        `[B,A,T,.]`, violations as `violations/<key>`."""
        dims, rin = self._rollout_structs(feat, gt, tf_mask, agent_type, agent_size, raw_map, latent_sample,
                                          latent_logp, dest, goal_valid, goal_gt, n_mode, n_step)
        B, A = dims.n_scene * dims.n_mode, dims.n_agent
        if out is None:
            out = self.alloc_outputs(B, A, n_step, trace)
        state = self._ensure_state(dims)
        rout = self._out_struct(out)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rollout(C.byref(dims), C.byref(rin), self.packed.data_ptr(), state.data_ptr(),
                                         C.byref(rout), nt.current_stream_ptr()), "tb_rollout")
        self.last_t = n_step
        return self._finish(out)

    def rule_checks(self, out: Dict[str, Tensor], gt: Mapping[str, Tensor], agent_type: Tensor, agent_size: Tensor,
                    raw_map: Mapping[str, Tensor], tl: Optional[Mapping[str, Tensor]], enable: Mapping[str, bool], n_mode: int = 1,
                    w_collision: float = 0.0, reduce_collision_with_max: bool = True, collision_size_scale: float = 1.1
                    ) -> Dict[str, Tensor]:
        """`tb_rule_checks`: the optional traffic-rule checks and the collision reward on the finished rollout
        `out` (the dict returned by `rollout`).  `tl`: dict(valid [S,Ttl,TL], pos [S,Ttl,TL,2], state [S,Ttl,TL,5]) -- the
        tensors the reference hands to `TrafficRuleChecker`.  Adds `violations/<key>` for the 8 optional keys to `out` and
        updates `diffbar_rewards` in place when `w_collision` > 0."""
        B, A, T = out["valid"].shape
        S, P = raw_map["valid"].shape[:2]
        Tg = gt["valid"].shape[1]
        mask = sum(bit for k, bit in nt.RULE_BITS.items() if enable.get(k))
        if (mask & 8) and not (mask & 4):
            raise nt.TbError("enable_check_passive needs enable_check_run_red_light (in the reference the combination raises a "
                             "NameError: traffic_rule_checker.py:441-442,457-464)")
        need_tl = bool(mask & 12)
        if need_tl and tl is None:
            raise nt.TbError("rule_checks: traffic-light tensors are required for run_red_light / passive")
        TL = tl["valid"].shape[2] if tl is not None else 1
        Ttl = tl["valid"].shape[1] if tl is not None else 1
        dims = self._dims(S, n_mode, A, P, TL, Tg, Tg, T)
        if B != S * n_mode:
            raise nt.TbError(f"rule_checks: {B} scene-modes, expected {S} x {n_mode}")
        p = nt.dev_ptr
        viol = torch.empty(8, B, A, T, dtype=torch.bool, device=self.device)
        rin = nt.TbRuleIn(
            p(out["preds"], "f32", (B, A, T, 4), "preds"), p(out["valid"], "u8", (B, A, T), "valid"),
            p(out["override_masks"], "u8", (B, A, T), "override_masks"),
            p(out["violations/outside_map_this_step"], "u8", (B, A, T), "outside_map_this_step"),
            p(gt["valid"], "u8", (S, Tg, A), "gt valid"), p(gt["pos"], "f32", (S, Tg, A, 2), "gt pos"),
            p(gt["yaw_bbox"], "f32", (S, Tg, A, 1), "gt yaw_bbox"), p(gt["spd"], "f32", (S, Tg, A, 1), "gt spd"),
            p(agent_type, "u8", (S, A, 3), "agent_type"), p(agent_size, "f32", (S, A, 3), "agent_size"),
            p(raw_map["valid"], "u8", (S, P, 20), "map/valid"), p(raw_map["type"], "u8", (S, P, 11), "map/type"),
            p(raw_map["pos"], "f32", (S, P, 20, 2), "map/pos"), p(raw_map["dir"], "f32", (S, P, 20, 2), "map/dir"),
            p(tl["valid"], "u8", (S, Ttl, TL), "tl valid") if tl is not None else None,
            p(tl["pos"], "f32", (S, Ttl, TL, 2), "tl pos") if tl is not None else None,
            p(tl["state"], "u8", (S, Ttl, TL, 5), "tl state") if tl is not None else None,
            Ttl, mask, collision_size_scale, float(w_collision), int(bool(reduce_collision_with_max)))
        rout = nt.TbRuleOut(viol.data_ptr(), out["diffbar_rewards"].data_ptr() if w_collision > 0 else None,
                            out["diffbar_rewards_valid"].data_ptr())
        ws = torch.empty(self.lib.tb_rule_workspace_bytes(C.byref(dims)), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rule_checks(C.byref(dims), C.byref(rin), C.byref(rout), ws.data_ptr(), nt.current_stream_ptr()),
                     "tb_rule_checks")
        for i, k in enumerate(nt.OPT_VIOLATION_KEYS):
            out["violations/" + k] = viol[i]
        return out

    def begin_rollout(self, *args, n_mode: int = 1, n_step: int = 90, **kw) -> Dict:
        """`tb_rollout_init` only; the steps are then driven one by one with `step(ctx)` (per-step `forward` use)."""
        dims, rin = self._rollout_structs(*args, n_mode, n_step, **kw)
        out = self.alloc_outputs(dims.n_scene * dims.n_mode, dims.n_agent, n_step)
        state = self._ensure_state(dims)
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rollout_init(C.byref(dims), C.byref(rin), self.packed.data_ptr(), state.data_ptr(),
                                              nt.current_stream_ptr()), "tb_rollout_init")
        self.last_t = 0
        return {"dims": dims, "rin": rin, "out": out, "rout": self._out_struct(out), "keep": (args, kw), "n_step": n_step}

    def step(self, ctx: Dict) -> int:
        t = self.last_t + 1
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rollout_steps(C.byref(ctx["dims"]), C.byref(ctx["rin"]), self.packed.data_ptr(),
                                               self._state.data_ptr(), C.byref(ctx["rout"]), t, t, nt.current_stream_ptr()),
                     "tb_rollout_steps")
        self.last_t = t
        return t

    def steps(self, ctx: Dict, t_first: int, t_last: int) -> None:
        """decode steps t_first..t_last of an opened rollout in one library call (`tb_rollout_steps`)."""
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rollout_steps(C.byref(ctx["dims"]), C.byref(ctx["rin"]), self.packed.data_ptr(),
                                               self._state.data_ptr(), C.byref(ctx["rout"]), t_first, t_last,
                                               nt.current_stream_ptr()), "tb_rollout_steps")
        self.last_t = t_last

    def reinit(self, ctx: Dict) -> None:
        """`tb_rollout_init` again on an opened rollout (back to frame 0): measurement aid."""
        with torch.cuda.device(self.device):
            nt.check(self.lib.tb_rollout_init(C.byref(ctx["dims"]), C.byref(ctx["rin"]), self.packed.data_ptr(),
                                              self._state.data_ptr(), nt.current_stream_ptr()), "tb_rollout_init")
        self.last_t = 0

    def _finish(self, out: Dict[str, Tensor]) -> Dict[str, Tensor]:
        res = {k: v for k, v in out.items() if not k.startswith("_")}
        for i, k in enumerate(VIOLATION_KEYS):
            res[f"violations/{k}"] = out["_violations"][i]
        # copies: the state buffer is reused (overwritten in place) by the next rollout on this engine
        res["hidden"] = self.state_field(nt.STATE_HIDDEN).clone()
        res["final_state"] = self.state_field(nt.STATE_AGENT_STATE).clone()
        res["final_valid"] = self.state_field(nt.STATE_VALID).clone()
        return res

    def state_field(self, field: int) -> Tensor:
        """typed VIEW into the simulation-state buffer (see `tb_state_field` in the header): valid until the next rollout
        on this engine starts; clone it to keep it."""
        d = self._state_dims
        B, A = d.n_scene * d.n_mode, d.n_agent
        off = self.lib.tb_rollout_state_offset(C.byref(d), field)
        raw = self._state
        f32 = lambda n: raw[off: off + 4 * n].view(torch.float32)  # noqa: E731
        u8 = lambda n: raw[off: off + n].view(torch.bool)  # noqa: E731
        if field == nt.STATE_AGENT_STATE:
            return f32(B * A * 4).view(B, A, 4)
        if field == nt.STATE_VALID:
            return u8(2 * B * A).view(2, B, A)[(self.last_t + 1) & 1]
        if field in (nt.STATE_KILLED, nt.STATE_GOAL_VALID):
            return u8(B * A).view(B, A)
        if field == nt.STATE_VEL:
            return f32(B * A * 2).view(B, A, 2)
        if field in (nt.STATE_ACC, nt.STATE_YAW_RATE):
            return f32(B * A).view(B, A)
        if field == nt.STATE_STICKY:
            return u8(3 * B * A).view(3, B, A)
        if field == nt.STATE_HIDDEN:
            return f32(3 * B * A * 128).view(3, B * A, 128)
        raise nt.TbError(f"unknown state field {field}")


def gt_from_batch(batch: Mapping[str, Tensor], n_frame: Optional[int] = None, prefix: str = "agent/") -> Dict[str, Tensor]:
    """the GT tensors the rollout overrides with (waymo_motion.py:434-466,523-548); `n_frame=11` gives test mode."""
    keys = ("valid", "pos", "yaw_bbox", "spd", "vel", "acc", "yaw_rate")
    if n_frame is None:
        return {k: batch[prefix + k] for k in keys}
    return {k: batch[prefix + k][:, :n_frame].contiguous() for k in keys}


def raw_map_from_batch(batch: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    return {k: batch["map/" + k] for k in ("boundary", "valid", "type", "pos", "dir")}

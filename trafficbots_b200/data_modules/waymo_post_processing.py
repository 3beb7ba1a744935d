"""`WaymoPostProcessing` -- mirror of the reference module (`src/data_modules/waymo_post_processing.py:8-81`) on
`tb_post_process`: same constructor keywords, same `forward(valid, scores, trajs, agent_type)` and `pred_dict` keys / shapes.
The mode selection (`mtr_nms`, top-k), `mpa_nms` (a triple Python loop in the reference) and the temperature softmax run
in one kernel with one CTA per (scene, agent); `trajs` may be the strided view `rollout_buffer.preds[:, :, :, t0:]` of the
rollout's own output, which is read in place.  `aggr_thresh` (k-means aggregation, default off) is not implemented."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch
from torch import Tensor, nn

from .. import _native as nt
from ..config import UnsupportedConfig


class WaymoPostProcessing(nn.Module):
    def __init__(self, k_pred: int = 6, score_temperature: float = 1e2, mpa_nms_thresh: Sequence[float] = (),
                 mtr_nms_thresh: Sequence[float] = (), aggr_thresh: Sequence[float] = (), n_iter_em: int = 3,
                 use_ade: bool = True) -> None:
        super().__init__()
        if len(list(aggr_thresh)) > 0:
            raise UnsupportedConfig("waymo_post_processing.aggr_thresh (k-means trajectory aggregation) is not implemented")
        for name, th in (("mpa_nms_thresh", mpa_nms_thresh), ("mtr_nms_thresh", mtr_nms_thresh)):
            if len(list(th)) not in (0, 3):
                raise UnsupportedConfig(f"waymo_post_processing.{name} must be empty or [veh, ped, cyc]")
        self.k_pred = int(k_pred)
        self.score_temperature = float(score_temperature)
        self.mpa_nms_thresh = [float(x) for x in mpa_nms_thresh]
        self.mtr_nms_thresh = [float(x) for x in mtr_nms_thresh]
        self.aggr_thresh = []
        self.n_iter_em = n_iter_em
        self.use_ade = bool(use_ade)

    def _cfg(self) -> nt.TbPostCfg:
        c = nt.TbPostCfg()
        c.k_pred, c.score_temperature, c.use_ade = self.k_pred, self.score_temperature, int(self.use_ade)
        c.n_mtr, c.n_mpa = len(self.mtr_nms_thresh), len(self.mpa_nms_thresh)
        for i, v in enumerate(self.mtr_nms_thresh):
            c.mtr_nms_thresh[i] = v
        for i, v in enumerate(self.mpa_nms_thresh):
            c.mpa_nms_thresh[i] = v
        return c

    def forward(self, valid: Tensor, scores: Tensor, trajs: Tensor, agent_type: Tensor) -> Dict[str, Optional[Tensor]]:
        """valid [S,A]; scores [S,A,n_pred] (not normalised); trajs [S,A,n_pred,Tf,4] (any strides with a dense last dim and
        consecutive steps); agent_type [S,A,3] -> pred_dict like the reference (:33-81) + `mode_idx`."""
        if not trajs.is_cuda:
            raise nt.TbError("WaymoPostProcessing: CUDA tensors expected (the hot path has no CPU implementation)")
        S, A, n, Tf, d = trajs.shape
        if d != 4:
            raise UnsupportedConfig("WaymoPostProcessing: trajectories must be (x, y, yaw, spd)")
        if trajs.dtype != torch.float32 or trajs.stride(4) != 1 or trajs.stride(3) != 4 or any(st % 4 for st in trajs.stride()[:3]) \
                or trajs.data_ptr() % 16:
            trajs = trajs.float().contiguous()
        k = min(self.k_pred, n)
        dev = trajs.device
        scores = scores.to(torch.float32).contiguous()
        valid = valid.contiguous()
        agent_type = agent_type.contiguous()
        w_trajs = torch.empty(S, Tf, A, k, 2, device=dev)
        w_yaw = torch.empty(S, Tf, A, k, 1, device=dev)
        w_spd = torch.empty(S, Tf, A, k, 1, device=dev)
        w_scores = torch.empty(S, A, k, device=dev)
        mode_idx = torch.empty(S, A, k, dtype=torch.int32, device=dev)
        cfg = self._cfg()
        p = nt.dev_ptr
        with torch.cuda.device(dev):
            nt.check(nt.lib().tb_post_process(S, A, n, Tf, trajs.data_ptr(), trajs.stride(0), trajs.stride(1), trajs.stride(2),
                                              p(scores, "f32", (S, A, n), "scores"), p(valid, "u8", (S, A), "valid"),
                                              p(agent_type, "u8", (S, A, 3), "agent_type"), C.byref(cfg), w_trajs.data_ptr(),
                                              w_yaw.data_ptr(), w_spd.data_ptr(), w_scores.data_ptr(), mode_idx.data_ptr(),
                                              nt.current_stream_ptr()), "tb_post_process")
        return {"waymo_trajs": w_trajs, "waymo_yaw_bbox": w_yaw, "waymo_spd": w_spd, "waymo_scores": w_scores,
                "waymo_valid": valid.unsqueeze(1).expand(-1, Tf, -1), "mode_idx": mode_idx}

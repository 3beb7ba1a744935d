"""`RolloutBuffer` with the attribute names / shapes of the reference's (`src/utils/buffer.py:8-123`) so that the
reference's metrics and post-processing consume it unchanged.  The CUDA rollout writes every field in its final
`[n_batch, n_agent, n_step, ...]` layout, so there is no per-step `add` / `finish` stacking here: the buffer is
created finished, from the tensors the kernels filled."""
from __future__ import annotations

from typing import Dict

import torch
from torch import Tensor

VIOLATION_ORDER = ("outside_map", "collided", "run_road_edge", "run_red_light", "passive", "goal_reached", "dest_reached")


class RolloutBuffer:
    def __init__(self, step_start: int, step_end: int, step_current: int, fields: Dict[str, Tensor]) -> None:
        self.step_start = step_start
        self.step_end = step_end
        self.step_future_start = step_current + 1 - step_start
        self.preds: Tensor = fields["preds"]  # [n_batch, n_agent, n_step, 4]
        self.valid: Tensor = fields["valid"]  # [n_batch, n_agent, n_step]
        self.override_masks: Tensor = fields["override_masks"]
        self.diffbar_rewards: Tensor = fields["diffbar_rewards"]
        self.diffbar_rewards_valid: Tensor = fields["diffbar_rewards_valid"]
        self.latent_log_probs: Tensor = fields["latent_log_probs"]
        self.action_log_probs: Tensor = fields["action_log_probs"]
        self.latents = []
        # the 14 keys of TrafficRuleChecker.check (traffic_rule_checker.py:499-515); the four checks that are off in the
        # default config report their sticky all-False state
        never = torch.zeros_like(self.valid)
        self.violations: Dict[str, Tensor] = {}
        for name in VIOLATION_ORDER:
            for key in (name, name + "_this_step"):
                self.violations[key] = fields.get("violations/" + key, never)
        self.vis_dicts: Dict[str, Tensor] = {}

    def flatten_repeat(self, n_repeat: int) -> None:
        """[n_scene * n_repeat, n_agent, ...] -> [n_scene, n_agent, n_repeat, ...] (views, no copy)."""
        def fr(x: Tensor) -> Tensor:
            return x.view(x.shape[0] // n_repeat, n_repeat, *x.shape[1:]).transpose(1, 2)

        for name in ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid", "latent_log_probs",
                     "action_log_probs"):
            setattr(self, name, fr(getattr(self, name)))
        self.violations = {k: fr(v) for k, v in self.violations.items()}
        self.vis_dicts = {k: fr(v) for k, v in self.vis_dicts.items()}

"""`ScenePipeline` -- keeps several scene batches in flight on one GPU.

Why: a batch of BASELINE.json configs[1] (32 scenes, K = 1) is 32 scene-modes = 32 CTAs of the persistent decode kernel,
which run for ~21 ms on 32 of the 148 SMs while the rest of the GPU idles (round 1 filled the idle SMs by running the
same program redundantly in 4-CTA clusters).  The reference's eval loop (`validation_step` per batch of the DataLoader,
`src/pl_modules/waymo_motion.py:574-601,683-690`) is a stream of independent batches, so the B200-native schedule is to keep
`depth` batches in flight, each on its own CUDA stream with its own forked `Engine` (same packed parameters, own
workspaces / simulation state) and 1-CTA clusters: the encode + heads of batch i+1.. run on the SMs the decode kernels of
batches i-3..i leave free.  Every batch's pinned-host -> device copy and device -> pinned-host read-back are queued on the
batch's own stream, so they overlap the other slots' kernels.  Results are bit-identical to sequential execution
(`tests/test_gpu_pipeline.py`): the slots share nothing but read-only parameters.

    pipe = ScenePipeline(module, depth=4)
    tickets = [pipe.submit(host_batch) for host_batch in loader]     # blocks only when all slots are busy
    for t in tickets: out = pipe.result(t)                           # dict of pinned host tensors (valid until the slot is reused)
"""
from __future__ import annotations

from typing import Callable, Dict, List, Mapping, Optional

import torch
from torch import Tensor

from . import engine as E

StepFn = Callable[[object, Mapping[str, Tensor]], Mapping[str, Tensor]]

RESULT_FIELDS = ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid", "action_log_probs",
                 "latent_log_probs")


def joint_future_step(module, cb: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    """the `joint_future_pred` leg of `validation_step` (waymo_motion.py:581-598) on the `WaymoMotion` surface:
    encode_input_features -> latent_encoder (prior) -> pred_goal -> joint_future_pred.  Returns the RolloutBuffer fields
    `[S, A, K, T, ...]` (+ the violation maps) as a flat dict of device tensors."""
    feat = module.model.encode_input_features(cb)
    latent = module.model.latent_encoder(**feat)
    goal = module.model.goal_manager.pred_goal(agent_type=cb["agent/type"], map_type=cb["map/type"], agent_state=None, **feat)
    goal_valid = cb["history/agent/valid"].any(1)
    buf, goal_sample, goal_logp = module.joint_future_pred(cb, feat, latent, goal, goal_valid, require_vis_dict=False)
    out = {name: getattr(buf, name) for name in RESULT_FIELDS}
    for name, v in buf.violations.items():
        out["violations/" + name] = v
    out["goal_sample"] = goal_sample
    out["goal_log_probs"] = goal_logp
    return out


class _Slot:
    def __init__(self, eng: E.Engine, stream: torch.cuda.Stream):
        self.eng = eng
        self.stream = stream
        self.dev_in: Dict[str, Tensor] = {}
        self.host_out: Dict[str, Tensor] = {}
        self.done = torch.cuda.Event(enable_timing=True)
        self.start = torch.cuda.Event(enable_timing=True)
        self.busy = False
        self.ticket = -1
        self.keep = None  # device results of the batch in flight (kept alive until the read-back has run)
        self.borrowed = set()  # data_ptrs of caller-owned resident inputs (never written by a later host batch)


class ScenePipeline:
    def __init__(self, module, depth: int = 4, step_fn: StepFn = joint_future_step, rollout_cluster: Optional[int] = None,
                 read_back: bool = True) -> None:
        """module: a `trafficbots_b200` `WaymoMotion` on a CUDA device.  `rollout_cluster`: CTAs per scene-mode of the decode
        kernel in every slot (default 1 for depth > 1: no redundant cluster ranks; None / 0 with depth 1 = library default)."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.module = module
        self.depth = depth
        self.step_fn = step_fn
        self.read_back = read_back
        base = module.engine()
        self.device = base.device
        cl = rollout_cluster if rollout_cluster is not None else (1 if depth > 1 else 0)
        self.rollout_cluster = cl
        self.slots: List[_Slot] = [_Slot(base.fork(cl), torch.cuda.Stream(self.device)) for _ in range(depth)]
        self._next_ticket = 0
        self._where: Dict[int, _Slot] = {}
        self.latencies: List[float] = []  # ms from a batch's first queued operation to its last, in collection order

    # ------------------------------------------------------------------------------------------------------
    def submit(self, host_batch: Mapping[str, Tensor]) -> int:
        """queues one batch (pinned host tensors, reference batch schema): H2D copy, the step, D2H read-back -- all on the
        slot's stream.  Returns a ticket for `result`.  Blocks (host side) only if the slot's previous batch has not been
        collected yet and is still running."""
        slot = self.slots[self._next_ticket % self.depth]
        if slot.busy:  # not collected: wait for it, the caller may still call result() for its ticket afterwards
            slot.done.synchronize()
        ticket = self._next_ticket
        self._next_ticket += 1
        # the module's own parameter check (re-pack if they changed) runs on the caller's stream, before the slot is entered
        self.module.use_engine(None)
        self.module.engine()
        # work the caller queued on its own stream (e.g. producing a device-resident batch) is ordered before the slot's
        slot.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(slot.stream):
            slot.start.record(slot.stream)
            for k, v in host_batch.items():
                if v.is_cuda:  # already resident (the caller keeps it alive and unchanged until the batch is collected)
                    slot.dev_in[k] = v
                    continue
                d = slot.dev_in.get(k)
                if d is None or d.shape != v.shape or d.dtype != v.dtype or d.data_ptr() in slot.borrowed:
                    d = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    slot.dev_in[k] = d
                d.copy_(v, non_blocking=True)
            slot.borrowed = {v.data_ptr() for v in host_batch.values() if v.is_cuda}
            self.module.use_engine(slot.eng)
            try:
                res = self.step_fn(self.module, slot.dev_in)
            finally:
                self.module.use_engine(None)
            if self.read_back:
                for k, v in res.items():
                    if k.startswith("_"):  # device-only results (e.g. the WOMD records that are all-gathered over NCCL)
                        continue
                    h = slot.host_out.get(k)
                    if h is None or h.shape != v.shape or h.dtype != v.dtype:
                        h = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                        slot.host_out[k] = h
                    h.copy_(v, non_blocking=True)
            slot.keep = res
            slot.done.record(slot.stream)
        slot.busy = True
        slot.ticket = ticket
        self._where[ticket] = slot
        return ticket

    def result(self, ticket: int) -> Mapping[str, Tensor]:
        """waits for the batch of `ticket` and returns its results: pinned host tensors (`read_back=True`; valid until the
        slot is reused `depth` submissions later) or the device tensors."""
        slot = self._where.pop(ticket)
        if slot.ticket != ticket:
            raise RuntimeError(f"ticket {ticket}: its slot has been reused (collect results within `depth` submissions)")
        slot.done.synchronize()
        slot.busy = False
        self.latencies.append(slot.start.elapsed_time(slot.done))
        if len(self.latencies) > 4096:
            del self.latencies[:2048]
        return slot.host_out if self.read_back else slot.keep

    def slot_of(self, ticket: int) -> _Slot:
        """the slot that holds `ticket` (its `done` event and device results `keep`), before `result(ticket)` is called."""
        return self._where[ticket]

    def drain(self) -> None:
        for s in self.slots:
            if s.busy:
                s.done.synchronize()
                s.busy = False
        self._where.clear()

    # bytes moved per batch (after at least one submit)
    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.slots[0].dev_in.values())

    def d2h_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.slots[0].host_out.values())

"""Benchmark of the TrafficBots hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic scenes (a "scene" = encode + prior
latent + destination prediction + K rollouts): scene encoding (`tb_encode_scene`), the pre-rollout heads (prior latent
encoder, destination predictor) and the 90-step closed-loop rollout of every scene-mode (`tb_rollout`).  Workload at N=1 = BASELINE.json
configs[1]: 32 scenes, 64 agents, 1024 map polylines, 91 frames, K=1.  With N>1 GPUs every rank processes its
own batch of 32 scenes (scene sharding, weak scaling, no data-path collective).

`--impl reference` times the reference's CPU implementation of the same path (the oracle port,
`oracle/trafficbots_oracle.py`, which is pinned to the unmodified reference) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "scenes/sec (64 agents, 91-step closed-loop rollout)"
WORKLOAD = dict(n_scene=32, n_agent=64, n_pl=1024, n_mode=1, n_step=90)


# ----------------------------------------------------------------------------------------------------------
# algorithmic work, D = F = 128
# ----------------------------------------------------------------------------------------------------------
def flops_front(A, P, TL, D=128):
    """k_step_front per scene-mode and step: agent encoder + 3 agent->map layers + 3 agent->TL layers (their K|V are
    pre-projected) + the interaction K|V projection of the step."""
    enc = 2 * A * (11 * 32 + 32 * 32)
    as2pl = 3 * (4 * A * D * D + 4 * A * P * D + 4 * A * D * D)
    as2tl = 3 * (4 * A * D * D + 4 * A * TL * D + 4 * A * D * D)
    kv_int = 3 * 4 * A * D * D
    return enc + as2pl + as2tl + kv_int


def flops_back(A, D=128):
    inter = 3 * (4 * A * D * D + 4 * A * A * D + 4 * A * D * D)
    gru = 3 * 12 * A * D * D
    # add_goal / add_latent mlp_out: Linear(256,128) + Linear(128,128); the z half of the first Linear is step-invariant and is
    # computed once per rollout by k_rollout_init, so it is NOT counted per step (SURVEY 8d's F_step counts it: 6AD^2 each)
    add = 2 * (6 - 2) * A * D * D
    head = 3 * 2 * A * (D * D + 2 * D)
    return inter + gru + add + head


def flops_map_encoder(P, D=128):
    node = 2 * 20 * P * (31 * 32 + 32 * 32)
    dense = 3 * (6 * 20 * P * D * D + 4 * 20 * P * 20 * D + 2 * 20 * P * D * D + 4 * 20 * P * D * D)
    glob = 6 * P * D * D + 4 * P * P * D + 2 * P * D * D + 4 * P * D * D
    return node + dense + glob


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the busiest half of the samples = "under load"
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(n_scene, n_agent, n_pl, n_mode, seed):
    """synthetic batch (trafficbots_b200.synthetic)."""
    from trafficbots_b200 import synthetic
    batch = synthetic.make_batch(n_scene, n_agent=n_agent, n_pl=n_pl, seed=seed)
    return batch, {}


USED_KEYS = ("map/valid", "map/type", "map/pos", "map/dir", "map/boundary", "agent/valid", "agent/pos", "agent/yaw_bbox",
             "agent/spd", "agent/vel", "agent/acc", "agent/yaw_rate", "agent/type", "agent/size", "agent/goal",
             "history/agent/valid", "history/agent/pos", "history/agent/yaw_bbox", "history/agent/vel", "history/agent/spd",
             "history/agent/yaw_rate", "history/agent/acc", "history/agent/size", "history/agent/type",
             "history/tl_stop/valid", "history/tl_stop/state", "history/tl_stop/pos", "history/tl_stop/dir")


def run_step(eng, cb, ex, n_mode, n_step, out=None):
    """device-resident step through the thin C-ABI driver: encode -> prior latent -> destination -> rollout.  With K = 1
    the single mode is the deterministic one (prior mean, arg-max destination: waymo_motion.py:489-500)."""
    from trafficbots_b200 import engine as E, host
    feat = eng.encode_scene(cb)
    lat_mean, _ = eng.latent_encoder(feat)
    probs, _logp, _ = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])
    dest = probs.argmax(-1)
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    goal_valid = cb["history/agent/valid"].any(1)
    lat_logp = ex["latent_logp"]
    return eng.rollout(feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), lat_mean, lat_logp, dest,
                       goal_valid, cb["agent/goal"], n_mode=n_mode, n_step=n_step, out=out)


def run_step_public(module, cb, ex):
    """the call sequence a user of the reference makes (validation_step's joint_future_pred leg, waymo_motion.py:581-598)
    through the `WaymoMotion` surface: encode_input_features -> latent_encoder -> pred_goal -> joint_future_pred."""
    feat = module.model.encode_input_features(cb)
    latent = module.model.latent_encoder(**feat)
    goal = module.model.goal_manager.pred_goal(agent_type=cb["agent/type"], map_type=cb["map/type"], agent_state=None, **feat)
    goal_valid = cb["history/agent/valid"].any(1)
    buf, _goal_sample, _goal_logp = module.joint_future_pred(cb, feat, latent, goal, goal_valid, require_vis_dict=False)
    return buf


# ----------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_scene, repeats=1, seed=1234):
    """the oracle port of the reference's CPU path (encode_scene + latent prior + destination predictor + rollout as the
    reference implements them: K|V re-projected every step) on `n_scene` scenes of the bench workload; returns scenes/s
    and seconds."""
    import trafficbots_oracle as orc
    from trafficbots_b200 import weights
    W = WORKLOAD
    sd = weights.init_state_dict(2023)
    batch, _ = make_inputs(n_scene, W["n_agent"], W["n_pl"], 1, seed)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():  # encode -> prior latent -> destination predictor -> rollout (K = 1: the deterministic mode)
            orc.joint_future_pred(sd, batch, k=1, sample_seed=0, step_end=W["n_step"])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_scene / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the per-step sample so that (steps + warmup) samples finish in ~2 minutes
    rate1, t1 = cpu_reference_rate(1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(max(1, min(WORKLOAD["n_scene"], budget / max(t1, 1e-3) * 0.8)))
    for _ in range(args.warmup):
        cpu_reference_rate(n)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_reference_rate(n)
        times.append(dt)
    total = sum(times)
    value = n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n} scenes/step (bounded sample of the 32-scene batch), 64 agents, 1024 polylines, K=1, "
                               "encode + prior latent + destination predictor + 90-step closed-loop rollout, reference CPU algorithm (oracle port)"},
        "cpu_baseline": {"value": value, "unit": "scenes/s", "cores": cores, "kind": "port",
                         "sample": f"{n} scenes x {args.steps} steps, torch {torch.__version__} CPU fp32, {cores} threads"},
        "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from trafficbots_b200 import _native as nt, engine as E, host, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU implementation; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    W = WORKLOAD
    S, A, P, K, T = W["n_scene"], W["n_agent"], W["n_pl"], W["n_mode"], W["n_step"]
    sd = weights.init_state_dict(2023)
    eng = E.Engine(sd, dev)
    batch, ex = make_inputs(S, A, P, K, seed=1000 + 100 * rank)
    host_batch = host.pin_batch({k: batch[k] for k in USED_KEYS})
    cb = host.batch_to_device(host_batch, dev)
    cex = {"latent_logp": torch.zeros(S * K, A, device=dev)}  # log-prob of the deterministic latent: not part of the timed math
    out = eng.alloc_outputs(S * K, A, T)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    h2d = sum(v.numel() * v.element_size() for v in host_batch.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        run_step(eng, cb, cex, K, T, out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = eng.lib.tb_launch_count()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations
        ev[i][0].record()
        run_step(eng, cb, cex, K, T, out)
        ev[i][1].record()
    barrier()
    launches = eng.lib.tb_launch_count() - n0
    ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end to end: pinned host buffers in, results back to the host, through the same public call ---------
    from trafficbots_b200 import config as tb_config, parallel
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion
    module = WaymoMotion(**tb_config.default_config(n_joint_future=K))
    module.load_state_dict(sd)
    module = module.to(dev).eval()
    host_res = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items() if not k.startswith("_")}
    host_res["violations"] = torch.empty(6, S * K, A, T, dtype=torch.bool).pin_memory()
    d2h = sum(v.numel() * v.element_size() for v in host_res.values())

    stager = host.SceneStager(dev)

    def e2e_step(last):
        cbe = stager.get()  # this step's inputs: pinned host -> device, queued on the staging stream
        stager.submit(host_batch)  # the next step's inputs travel while this step computes (one copy per step, steady state)
        buf = run_step_public(module, cbe, None)  # RolloutBuffer after flatten_repeat: [S, A, K, T, ...]
        pairs = [(host_res[name], getattr(buf, name).squeeze(2)) for name in
                 ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid", "action_log_probs",
                  "latent_log_probs")]
        pairs += [(host_res["violations"][i], buf.violations[name].squeeze(2)) for i, name in enumerate(E.VIOLATION_KEYS)]
        stager.read_back(pairs)  # results -> pinned host, overlapping the next step's kernels
        if world > 1:  # the metrics reduction: one packed all-gather per step
            local = parallel.pack_scene_metrics(buf.preds, buf.valid, buf.violations, buf.diffbar_rewards, cbe["agent/pos"],
                                                cbe["agent/valid"])
            parallel.all_gather_scenes(local, world * S)
        if last:
            stager.join()  # the last step's read-back (and the copy it started) end inside its timed region
        return buf
    stager.submit(host_batch)
    for i in range(2):
        e2e_step(False)
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev2[i][0].record()
        e2e_step(i == args.steps - 1)
        ev2[i][1].record()
    barrier()
    wall_e2e = time.perf_counter() - t_wall0
    ms_e2e = sum(a.elapsed_time(b) for a, b in ev2)
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel: the persistent decode kernel (all 90 steps of every scene-mode in ONE launch), timed alone ---
    feat = eng.encode_scene(cb)
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    lat_mean, _ = eng.latent_encoder(feat)
    dest = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])[0].argmax(-1)
    rollout_ms = eng.profile_rollout(feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb),
                                     lat_mean, cex["latent_logp"], dest, cb["history/agent/valid"].any(1),
                                     cb["agent/goal"], n_mode=K, n_step=T, out=out)
    torch.cuda.synchronize()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
        B = S * K
        f_roll = (flops_front(A, P, 40) + flops_back(A)) * B * T  # per launch: all scene-modes of this rank, all steps
        ach = f_roll / (rollout_ms * 1e-3) / 1e12
        f_total = (flops_front(A, P, 40) + flops_back(A)) * B * T + flops_map_encoder(P) * S
        # CPU baseline: the oracle port on a bounded sample (about 10-30 s of CPU work)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_cpu = 32  # the whole batch of the GPU arm: ~10 s on 16 cores
        if world == 1:
            cpu_rate, cpu_s = cpu_reference_rate(n_cpu)
        line = {
            "metric": METRIC, "value": world * S * args.steps / (ms * 1e-3), "unit": "scenes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{S} scenes/GPU/step, {A} agents, {P} map polylines, 40 TL, K={K}, encode_scene + {T}-step "
                                   "closed-loop rollout (BASELINE.json configs[1]), incl. the pre-rollout heads (prior latent encoder, "
                                   "destination predictor; the K = 1 mode is the deterministic one)",
                       "l2": "256 MiB flush write between timed iterations", "timing": "CUDA events per step, summed, max over ranks",
                       "e2e_staging": "double-buffered (host.SceneStager): the pinned-host -> device copy of step i+1 and the device -> "
                                      "pinned-host read-back of step i-1 run on a side stream under step i's kernels; every step's "
                                      "copies are issued and completed inside the timed steps",
                       "weights": "seeded random init (no checkpoint distributable)"},
            "e2e": {"value": world * S * args.steps / (ms_e2e * 1e-3), "unit": "scenes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_rollout_tc16 (persistent decode kernel: 90 steps x (embed, 9 attention layers, 3 GRU layers, "
                                   "add_goal, add_latent, action head, dynamics/rule-check tail), one 4-CTA cluster per scene-mode)",
                         "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                         "traffic": 11.64e9 if (world == 1 and S == 32) else None,
                         "traffic_note": "dram read 10.12 GB + write 1.52 GB per launch, ncu --set full (profiles/r1o_k_rollout_tc16.txt: 17.9 ms, tensor "
                                         "pipe active 35.3 %, issue active 34.9 %, L2 hit 67.9 %): the 88 MB of key blocks stream through L2 at every "
                                         "one of the 90 steps; ~650 GB/s = 10 % of HBM peak, the kernel is bound by its serial GEMM -> epilogue chain",
                         "peak_source": peak_src, "flops_per_launch": f_roll,
                         "avg_launch_ms": rollout_ms, "whole_step_tflops": f_total / (ms / args.steps * 1e-3) / 1e12,
                         "whole_step_frac": f_total / (ms / args.steps * 1e-3) / 1e12 / peak_tf,
                         "attention_frac": 4.0 * A * (P + 40 + A) * 128 * 3 * B * T / (rollout_ms * 1e-3) / 1e12 / peak_tf,
                         "hbm_frac": (11.64e9 / (rollout_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbps", 6553.6)))
                                     if (world == 1 and S == 32) else None,
                         "note": "algorithmic fp32-equivalent FLOPs (SURVEY 8d); the kernel issues 3 bf16 MMAs per logical "
                                 "product (bf16x3) on M=128 tiles holding 64 agents, on 4 x B = 128 of the 148 SMs"},
            "cpu_baseline": ({"value": cpu_rate, "unit": "scenes/s", "cores": cores, "kind": "port",
                              "sample": f"{n_cpu} scenes of the same workload, 1 pass ({cpu_s:.1f} s), torch CPU fp32"}
                             if world == 1 else None),  # timed at N = 1 only
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

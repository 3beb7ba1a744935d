"""Side measurements of the other BASELINE.json configurations on ONE GPU (bench.py stays on configs[1]):

  python tools/bench_configs.py --scenes 32 --agents 64 --pl 1024 --k 6      # per-GPU slice of configs[2] (K = 6 modes)
  python tools/bench_configs.py --scenes 148 --agents 128 --pl 2048 --k 1    # configs[4], the 128-agent stress shape

Times the public call sequence (encode_input_features -> latent_encoder -> pred_goal -> joint_future_pred on the
`WaymoMotion` surface) with the batch resident in HBM, CUDA events around every step, 256 MiB L2 flush between steps;
prints one JSON line with scenes/s, scene-modes/s and the tensor-roofline fraction of SURVEY 8d's algorithmic FLOPs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=32)
    ap.add_argument("--agents", type=int, default=64)
    ap.add_argument("--pl", type=int, default=1024)
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from trafficbots_b200 import config as tb_config, host, weights
    from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    S, A, P, K = args.scenes, args.agents, args.pl, args.k
    sd = weights.init_state_dict(2023)
    module = WaymoMotion(**tb_config.default_config(n_joint_future=K))
    module.load_state_dict(sd)
    module = module.to(dev).eval()
    batch, _ = bench.make_inputs(S, A, P, K, seed=4242)
    cb = host.batch_to_device({k: batch[k] for k in bench.USED_KEYS}, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    for _ in range(args.warmup):
        bench.run_step_public(module, cb, None)
    torch.cuda.synchronize()
    n0 = module.engine().lib.tb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i)
        ev[i][0].record()
        buf = bench.run_step_public(module, cb, None)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    f_roll = (bench.flops_front(A, P, 40) + bench.flops_back(A)) * S * K * 90
    f_total = f_roll + bench.flops_map_encoder(P) * S
    line = {
        "workload": f"{S} scenes x K={K} modes, {A} agents, {P} polylines, 90 steps, 1 GPU, device-resident inputs",
        "ms_per_step": ms, "scenes_per_s": S / (ms * 1e-3), "scene_modes_per_s": S * K / (ms * 1e-3),
        "algorithmic_tflops": f_total / (ms * 1e-3) / 1e12, "tensor_frac_of_1400": f_total / (ms * 1e-3) / 1e12 / 1400.0,
        "preds_shape": list(buf.preds.shape), "finite": bool(torch.isfinite(buf.preds).all()),
        "launches_per_step": (module.engine().lib.tb_launch_count() - n0) / args.steps,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

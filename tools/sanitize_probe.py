"""Reduced configuration for `compute-sanitizer` (memcheck / racecheck / synccheck): one small scene, a TWO-step rollout through the
persistent decode kernel (cluster size from TB_CLUSTER), the scene encoder, the pre-rollout heads, the optional rule checks and
the post-processing kernels.  The sanitizers slow the persistent tcgen05 kernels by 2-3 orders of magnitude, hence the two steps.

  compute-sanitizer --tool racecheck python tools/sanitize_probe.py [--steps 2] [--agents 8] [--pl 64]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--agents", type=int, default=8)
    ap.add_argument("--pl", type=int, default=64)
    ap.add_argument("--scenes", type=int, default=1)
    args = ap.parse_args()
    from trafficbots_b200 import engine as E, host, synthetic, weights
    from trafficbots_b200.data_modules.waymo_post_processing import WaymoPostProcessing
    from trafficbots_b200.models.metrics.womd import WOMDMetrics
    sd = weights.init_state_dict(2023)
    batch = synthetic.make_batch(args.scenes, n_agent=args.agents, n_pl=args.pl, seed=11, area_scale=0.3, plant_red_light=True)
    eng = E.Engine(sd, "cuda:0")
    cb = host.batch_to_device(batch, "cuda:0")
    feat = eng.encode_scene(cb)
    lat, _ = eng.latent_encoder(feat)
    probs, _, _ = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    S, A = args.scenes, args.agents
    out = eng.rollout(feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), lat, torch.zeros(S, A, device="cuda:0"),
                      probs.argmax(-1), cb["history/agent/valid"].any(1), cb["agent/goal"], n_mode=1, n_step=args.steps)
    tl = {k: cb[f"history/tl_stop/{k}"] for k in ("valid", "pos", "state")}
    eng.rule_checks(out, gt, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), tl,
                    {"collided": True, "run_road_edge": True, "run_red_light": True, "passive": True}, w_collision=0.5)
    if args.steps >= 90:
        pp = WaymoPostProcessing(k_pred=6, mpa_nms_thresh=[2.5, 1.0, 1.5])
        d = pp(out["valid"].any(-1), torch.ones(S, A, 1, device="cuda:0"), out["preds"].unsqueeze(2)[:, :, :, 10:], cb["agent/type"])
        WOMDMetrics().update(cb, d["waymo_trajs"], d["waymo_scores"])
    torch.cuda.synchronize()
    print("sanitize_probe: finished,", int(eng.lib.tb_launch_count()), "kernels, preds finite:", bool(torch.isfinite(out["preds"]).all()))


if __name__ == "__main__":
    main()

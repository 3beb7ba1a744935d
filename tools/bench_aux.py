"""Timing of the kernels around the decode path that round 2 added: `tb_rule_checks` (optional traffic-rule
checks + collision reward), `tb_post_process`, `tb_womd_pack` -- CUDA events, L2 flushed between repeats, on the shapes of
BASELINE.json configs[1] / configs[2]; beside each, the CPU restatement of the reference (`oracle/`) on the host cores for the
same work, and the algorithmic bytes / achieved GB/s against the measured HBM peak (these kernels are byte / integer / geometry
work: no tensor cores).  Prints one JSON line per kernel.

  python tools/bench_aux.py [--cpu]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402


def timed(fn, flush, repeats=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(repeats):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time the CPU restatement (slow for the rule checks)")
    args = ap.parse_args()
    from trafficbots_b200 import engine as E, host, synthetic, weights
    from trafficbots_b200.data_modules.waymo_post_processing import WaymoPostProcessing
    from trafficbots_b200.models.metrics.womd import WOMDMetrics
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6500.0))
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    S, A, P, T = 32, 64, 1024, 90
    sd = weights.init_state_dict(2023)
    batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=1000)
    cb = host.batch_to_device(batch, dev)
    eng = E.Engine(sd, dev)
    feat = eng.encode_scene(cb)
    lat, _ = eng.latent_encoder(feat)
    dest = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])[0].argmax(-1)
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    out = eng.rollout(feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), lat, torch.zeros(S, A, device=dev), dest,
                      cb["history/agent/valid"].any(1), cb["agent/goal"], n_mode=1, n_step=T)
    tl = {k: cb[f"history/tl_stop/{k}"] for k in ("valid", "pos", "state")}
    on = {"collided": True, "run_road_edge": True, "run_red_light": True, "passive": True}
    rm = E.raw_map_from_batch(cb)
    ms = timed(lambda: eng.rule_checks(dict(out), gt, cb["agent/type"], cb["agent/size"], rm, tl, on, w_collision=0.5), flush)
    # work: per (scene, step): A^2 box pairs x 32 line-point tests, A x valid road-edge segments culls, A x lane nodes, A x TL, A^2 x 25 circles
    line = {"kernel": "tb_rule_checks (k_rule_compact + k_rule_step + k_rule_sticky), all four checks + collision reward",
            "shape": f"{S} scenes x {A} agents x {P} polylines x {T} steps", "ms": ms, "scene_steps_per_s": S * T / (ms * 1e-3),
            "note": "runs once per rollout, after it (the reference evaluates the checks inside the step loop, A x 20 P segment "
                    "tests per step in PyTorch); share of the 8.9 ms scene step: %.1f %%" % (100 * ms / 8.9)}
    if args.cpu:
        import rule_checks_oracle as rco
        sub = {k: v[:2] for k, v in batch.items()}
        rs = rco.init_rules(sub["agent/type"], sub["agent/size"], sub["map/valid"], sub["map/type"], sub["map/pos"], sub["map/dir"],
                            sub["history/tl_stop/valid"], sub["history/tl_stop/pos"], sub["history/tl_stop/state"], on)
        st = torch.cat([sub["agent/pos"], sub["agent/yaw_bbox"], sub["agent/spd"]], -1)
        t0 = time.perf_counter()
        for t in range(1, 11):
            rco.check_optional(rs, t, sub["agent/valid"][:, t], st[:, t])
        dt = (time.perf_counter() - t0) / 10 / 2  # per scene-step
        line["cpu_port_scene_steps_per_s"] = 1.0 / dt
        line["cpu_cores"] = os.cpu_count()
    print(json.dumps(line), flush=True)

    # ---- post-processing + WOMD packing on the K = 6 shape ------------------------------------------------------------------
    K = 6
    valid, scores, trajs = synthetic.make_mode_trajectories(S, A, K, seed=7, n_step=T)
    raw = trajs.transpose(1, 2).reshape(S * K, A, T, 4).contiguous().to(dev)  # the rollout kernels' output layout
    view = raw.view(S, K, A, T, 4).transpose(1, 2)[:, :, :, 10:]
    v_d, s_d = valid.to(dev), scores.to(dev)
    for name, pp in (("default (temperature softmax only)", WaymoPostProcessing(k_pred=6)),
                     ("mpa_nms", WaymoPostProcessing(k_pred=6, mpa_nms_thresh=[2.5, 1.0, 1.5]))):
        ms = timed(lambda: pp(v_d, s_d, view, cb["agent/type"]), flush)
        nbytes = S * A * K * 80 * (16 + 16)  # trajectories read (x, y, yaw, spd) + written (xy | yaw | spd)
        if "mpa" in name:
            nbytes += S * A * K * 80 * 16  # second pass over the selected trajectories for the pairwise distances (L2)
        line = {"kernel": "tb_post_process, " + name, "shape": f"{S} scenes x {A} agents x K={K} x 80 steps", "ms": ms,
                "algorithmic_bytes": nbytes, "gb_per_s": nbytes / (ms * 1e-3) / 1e9, "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / hbm}
        if args.cpu:
            import post_oracle as po
            t0 = time.perf_counter()
            po.post_process(valid, scores, trajs[:, :, :, 10:], batch["agent/type"], 6, 1e2, pp.mpa_nms_thresh, [], True)
            line["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
        print(json.dumps(line), flush=True)
    pp = WaymoPostProcessing(k_pred=6)
    d = pp(v_d, s_d, view, cb["agent/type"])
    m = WOMDMetrics()
    ms = timed(lambda: (m.update(cb, d["waymo_trajs"], d["waymo_scores"]), m.reset()), flush)
    rec_bytes, _ = m.record_layout(A, K)
    nbytes = S * rec_bytes + S * A * 91 * (8 + 8 + 4 + 1)  # records written + GT read
    line = {"kernel": "tb_womd_pack", "shape": f"{S} scenes x {A} agents x K={K}", "ms": ms, "record_bytes_per_scene": rec_bytes,
            "algorithmic_bytes": nbytes, "gb_per_s": nbytes / (ms * 1e-3) / 1e9, "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / hbm,
            "note": "one CTA per scene (32 CTAs): launch-latency sized, not bandwidth sized"}
    if args.cpu:
        import post_oracle as po
        t0 = time.perf_counter()
        po.womd_pack(batch, d["waymo_trajs"].cpu(), d["waymo_scores"].cpu())
        line["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

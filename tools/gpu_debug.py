"""Development aid (GPU box): runs every stage of the CUDA path against the oracle and prints the error of each
stage instead of stopping at the first failure.  Usage: python tools/gpu_debug.py [case ...]"""
from __future__ import annotations

import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import trafficbots_oracle as orc  # noqa: E402
from golden_util import load_case  # noqa: E402
from trafficbots_b200 import engine as E, host, weights  # noqa: E402


def md(a, b):
    return float((a.detach().cpu().float() - b.detach().cpu().float()).abs().max())


def stage_xlayer(eng, sd):
    g = torch.Generator().manual_seed(1)
    for block, prefix in ((2, "model.transformer_as2pl"), (4, "model.agent_interaction.transformer")):
        for (ns, nk, mask_self) in ((64, 1024, False), (8, 40, False), (37, 37, True)):
            src = torch.randn(2, ns, 128, generator=g)
            tgt = src.clone() if mask_self else torch.randn(2, nk, 128, generator=g)
            sv = torch.rand(2, ns, generator=g) < 0.8
            kvd = sv.clone() if mask_self else torch.rand(2, nk, generator=g) < 0.7
            if not mask_self:
                kvd[0] = False
            pfx = prefix + ".layers.1"
            kv = eng.kv_project(block, 1, tgt.cuda())
            t2 = orc.layer_norm(tgt, sd, pfx + ".norm_tgt")
            kv_ref = torch.nn.functional.linear(t2, sd[pfx + ".attn.in_proj_weight"][128:], sd[pfx + ".attn.in_proj_bias"][128:])
            out = eng.xlayer(block, 1, src.cuda(), sv.cuda(), kv, kvd.cuda(), mask_self=mask_self)
            ref = orc.xlayer(sd, pfx, src, ~sv, tgt, ~kvd, torch.eye(ns, dtype=torch.bool) if mask_self else None)
            print(f"  xlayer block {block} ns={ns} nk={nk} self={mask_self}: kv {md(kv, kv_ref):.2e}  out {md(out, ref):.2e}"
                  f"  (nan: {bool(torch.isnan(out).any())})")


def stage_case(name):
    gold, sd, batch, meta = load_case(name)
    eng = E.Engine(sd, "cuda")
    cb = host.batch_to_device(batch, "cuda")
    t0 = time.time()
    feat = eng.encode_scene(cb)
    torch.cuda.synchronize()
    print(f"  encode_scene: {time.time() - t0:.3f}s")
    ref = orc.encode_scene(sd, batch)
    print("  map_feature_valid equal:", torch.equal(feat["map_feature_valid"].cpu(), ref["map_feature_valid"]))
    for k in ("map_feature", "agent_feature", "tl_feature"):
        print(f"  {k}: vs oracle {md(feat[k], ref[k]):.2e}  vs golden {md(feat[k], gold['enc/' + k]):.2e}")
    K, S = meta["K"], meta["S"]
    A = batch["agent/type"].shape[1]
    oref = orc.joint_future_pred(sd, batch, k=K, sample_seed=meta["sseed"], return_trace=True)
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    lat = oref["latent_sample"].cuda()
    lat_logp = oref["latent_log_probs"][:, :, :, 0].transpose(1, 2).reshape(S * K, A).contiguous().cuda()
    dest = oref["goal_sample"].transpose(1, 2).reshape(S * K, A).contiguous().cuda()
    goal_valid = batch["history/agent/valid"].any(1).repeat_interleave(K, 0).cuda()
    t0 = time.time()
    out = eng.rollout(feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), lat, lat_logp, dest,
                      goal_valid, cb["agent/goal"], n_mode=K, n_step=90, trace=True)
    torch.cuda.synchronize()
    print(f"  rollout: {time.time() - t0:.3f}s")
    sh = lambda x: x.cpu().view(S, K, *x.shape[1:]).transpose(1, 2)  # noqa: E731
    for k in ("valid", "override_masks", "diffbar_rewards_valid"):
        print(f"  {k} equal: {torch.equal(sh(out[k]), oref[k])}")
    for k in ("outside_map", "outside_map_this_step", "goal_reached", "goal_reached_this_step", "dest_reached", "dest_reached_this_step"):
        eq = torch.equal(sh(out['violations/' + k]), oref[k])
        print(f"  violations/{k} equal: {eq}  (any: {bool(oref[k].any())})")
    p, q = sh(out["preds"]), oref["preds"]
    per_t = (p - q).abs().amax(dim=(0, 1, 2, 4))
    print("  preds max|diff| per step:", " ".join(f"{v:.1e}" for v in per_t.tolist()))
    pf = (sh(out["trace/policy_feature"]) - oref["trace/policy_feature"]).abs().amax(dim=(0, 1, 2, 4))
    print("  policy_feature max|diff| per step:", " ".join(f"{v:.1e}" for v in pf.tolist()))
    am = (sh(out["trace/action_mean"]) - oref["trace/action_mean"]).abs().amax(dim=(0, 1, 2, 4))
    print("  action_mean max|diff| per step:", " ".join(f"{v:.1e}" for v in am[:12].tolist()))
    for k in ("diffbar_rewards", "action_log_probs", "latent_log_probs"):
        print(f"  {k}: {md(sh(out[k]), oref[k]):.2e}")
    print(f"  hidden: {md(out['hidden'], oref['hidden']):.2e}   vs golden preds {md(sh(out['preds']), gold['jfp/preds']):.2e}")


def main():
    cases = sys.argv[1:] or ["cfg1_s1_a8_p64_k1", "s3_a8_p64_k2", "s1_a64_p1024_k1"]
    sd = weights.init_state_dict(11)
    eng = E.Engine(sd, "cuda")
    print("== building blocks")
    try:
        stage_xlayer(eng, sd)
    except Exception:
        traceback.print_exc()
    for c in cases:
        print("== case", c)
        try:
            stage_case(c)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()

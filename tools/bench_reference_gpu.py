"""Secondary baseline (BASELINE.md 3 / SURVEY 8d): the UNMODIFIED reference's own PyTorch path on ONE B200 -- fp32 and fp16
autocast -- for the same step the bench times (pre_processing -> encode_input_features x 2 -> latent prior -> pred_goal ->
joint_future_pred, K = 1 or 6), on the same synthetic scenes.  Needs the reference sources (tools/install_reference.sh ->
baseline/_ref, shipped to the GPU box by gpurun).  Reported in DESIGN.md as `gpu_reference_baseline`; never the headline.

  python tools/bench_reference_gpu.py [--scenes 32] [--k 1] [--steps 3] [--amp]
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
import torch._dynamo  # noqa: E402,F401  (torch.optim imports it lazily; must precede ref_loader's module stubs)


def train(model, batch, dev, args):
    """the reference's training_step + loss.backward() + clip + Adam step on the GPU, module as shipped (dropout 0.1) -- the
    training counterpart of the secondary baseline."""
    import ref_train
    model.train()  # as shipped: dropout 0.1
    model.log = lambda *a, **k: None
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    b = {k: v.to(dev) for k, v in batch.items()}
    times = []
    with torch.autocast("cuda", dtype=torch.float16, enabled=args.amp):
        for i in range(args.steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = model.training_step({k: v.clone() for k, v in b.items()}, i)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            torch.cuda.synchronize()
            if i > 0:
                times.append(time.perf_counter() - t0)
    best = min(times)
    print(json.dumps({"impl": "reference on GPU (unmodified PyTorch training_step + backward + Adam)",
                      "precision": "fp16 autocast" if args.amp else "fp32", "scenes": args.scenes, "agents": args.agents,
                      "polylines": args.pl, "s_per_step": best, "scenes_per_s": args.scenes / best, "gpu": torch.cuda.get_device_name(0),
                      "loss": float(loss), "peak_memory_gib": torch.cuda.max_memory_allocated() / 2 ** 30}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=32)
    ap.add_argument("--agents", type=int, default=64)
    ap.add_argument("--pl", type=int, default=1024)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--train", action="store_true", help="training_step + backward (BASELINE.json configs[3]) instead of the eval step")
    args = ap.parse_args()
    import ref_loader
    from trafficbots_b200 import synthetic, weights
    if not ref_loader.reference_available():
        print(json.dumps({"unavailable": "reference sources not found (run tools/install_reference.sh in the build container)"}))
        return
    dev = torch.device("cuda", 0)
    model = ref_loader.build_reference(n_agent=args.agents, n_pl=args.pl, n_joint_future=args.k)
    model.load_state_dict(weights.init_state_dict(2023), strict=True)
    model = model.to(dev).eval()
    batch = synthetic.make_batch(args.scenes, n_agent=args.agents, n_pl=args.pl, seed=1000)
    if args.train:
        return train(model, batch, dev, args)

    @torch.no_grad()
    def step():
        b = {k: v.to(dev) for k, v in batch.items()}
        b = model.pre_processing(b)
        input_dict = {k.split("input/")[-1]: v for k, v in b.items() if "input/" in k}
        prior_dict = {k.split("latent_prior/")[-1]: v for k, v in b.items() if "latent_prior/" in k}
        feat = model.model.encode_input_features(**input_dict)
        feat_prior = model.model.encode_input_features(**prior_dict)
        goal_valid = input_dict["agent_valid"].any(1)
        goal = model.model.goal_manager.pred_goal(agent_type=b["ref/agent_type"], map_type=b["ref/map_type"],
                                                  agent_state=b["ref/agent_state"], **feat)
        latent = model.model.latent_encoder(**feat_prior)
        buf, _, _ = model.joint_future_pred(batch=b, input_feature_dict=feat, latent=latent, goal=goal, goal_valid=goal_valid,
                                            require_vis_dict=False)
        return buf

    times = []
    with torch.autocast("cuda", dtype=torch.float16, enabled=args.amp):
        step()
        torch.cuda.synchronize()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            buf = step()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    best = min(times)
    print(json.dumps({"impl": "reference on GPU (unmodified PyTorch path)", "precision": "fp16 autocast" if args.amp else "fp32",
                      "scenes": args.scenes, "k": args.k, "agents": args.agents, "polylines": args.pl,
                      "s_per_step": best, "scenes_per_s": args.scenes / best, "gpu": torch.cuda.get_device_name(0),
                      "finite": bool(torch.isfinite(buf.preds).all()), "note": "encodes the map twice per step like test_step "
                      "(validation_step: three times); one process, default stream, no CUDA graphs -- the reference as shipped"}))


if __name__ == "__main__":
    main()

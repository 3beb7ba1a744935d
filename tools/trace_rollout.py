"""Development aid (GPU box): clock64 marks of the persistent tensor-core rollout kernel (CTA 0, worker thread 0) for the
bench workload; prints the average cycles per phase of a decode step.  Usage: python tools/trace_rollout.py [n_scene]"""
from __future__ import annotations

import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import bench  # noqa: E402
from trafficbots_b200 import engine as E, host, weights  # noqa: E402

PHASES = ("embed", "9 attention layers", "GRU x3", "add_goal/add_latent/head", "tail")


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    sd = weights.init_state_dict(2023)
    eng = E.Engine(sd, "cuda")
    batch, _ = bench.make_inputs(S, 64, 1024, 1, seed=1000)
    cb = host.batch_to_device(batch, "cuda")
    feat = eng.encode_scene(cb)
    gt = E.gt_from_batch(cb)
    tf = host.teacher_forcing_mask(gt["valid"], 10, 10)
    lat_mean, _ = eng.latent_encoder(feat)
    dest = eng.dest_predictor(feat, cb["agent/type"], cb["map/type"])[0].argmax(-1)
    args = (feat, gt, tf, cb["agent/type"], cb["agent/size"], E.raw_map_from_batch(cb), lat_mean,
            torch.zeros(S, 64, device="cuda"), dest, cb["history/agent/valid"].any(1), cb["agent/goal"])
    for _ in range(2):
        eng.rollout(*args, n_mode=1, n_step=90)
    torch.cuda.synchronize()
    trace = torch.zeros(1024 + 2800, dtype=torch.int64, device="cuda")
    eng.lib.tb_debug_set_trace(C.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.rollout(*args, n_mode=1, n_step=90)
    e1.record()
    torch.cuda.synchronize()
    eng.lib.tb_debug_set_trace(C.c_void_p(0))
    full = trace.cpu().tolist()
    tr = full[:1024]
    n = max(i for i, v in enumerate(tr) if v) + 1
    print(f"rollout (init + 90 steps): {e0.elapsed_time(e1):.3f} ms; n_key_map[0] = {int(feat['_n_key_map'][0])}; marks {n}; "
          f"kernel cycles {tr[n - 1] - tr[0]}")
    per = len(PHASES)
    steps = (n - 1) // per
    tot = [0] * per
    for st in range(steps):
        for i in range(per):
            tot[i] += tr[st * per + i + 1] - tr[st * per + i]
    for i, name in enumerate(PHASES):
        print(f"  {name:28s} {tot[i] / steps:10.0f} cycles/step")
    print(f"  {'step':28s} {sum(tot) / steps:10.0f} cycles/step")


    for name, base in (("group 0 leader (tid 0)", 1024), ("group 1 leader (tid 128 or 256)", 1024 + 1400)):
        d = full[base:base + 1400]
        pairs = [(d[2 * i], d[2 * i + 1]) for i in range(700) if d[2 * i]]
        print(f"detailed marks, {name}: id:+cycles since previous mark")
        line = []
        for i, (pid, clk) in enumerate(pairs):
            line.append(f"{pid}:{clk - pairs[i - 1][1] if i else 0}")
        for i in range(0, len(line), 12):
            print("   " + "  ".join(line[i:i + 12]))


if __name__ == "__main__":
    main()

"""Development aid (GPU box): clock64 marks of the tensor-core decode kernel (CTA 0, worker thread 0) for the bench
workload; prints the cycle deltas between consecutive marks.  Usage: python tools/trace_front.py [n_scene]"""
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
    ctx = eng.begin_rollout(*args, n_mode=1, n_step=90)
    for _ in range(20):
        eng.step(ctx)
    torch.cuda.synchronize()
    trace = torch.zeros(512, dtype=torch.int64, device="cuda")
    eng.lib.tb_debug_set_trace(C.c_void_p(trace.data_ptr()))
    eng.step(ctx)
    torch.cuda.synchronize()
    eng.lib.tb_debug_set_trace(C.c_void_p(0))
    tr = trace.cpu().tolist()
    n = max(i for i, v in enumerate(tr) if v) + 1
    print("n_key_map[0] =", int(feat["_n_key_map"][0]), " marks:", n, " total cycles:", tr[n - 1] - tr[0])
    d = [tr[i + 1] - tr[i] for i in range(n - 1)]
    for i in range(0, len(d), 16):
        print(f"{i:4d}: " + " ".join(f"{v:6d}" for v in d[i:i + 16]))


if __name__ == "__main__":
    main()

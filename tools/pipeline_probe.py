"""Probe: throughput of the full scene step (encode + heads + rollout) with D batches in flight on D CUDA streams, for
decode-kernel cluster sizes 1 / 2 / 4 (TB_CLUSTER).  Prints one JSON line per (cluster, depth) combination.

  python tools/pipeline_probe.py [--depths 1,2,4,5,6] [--clusters 4,2,1] [--batches 24]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depths", default="1,2,4,5,6")
    ap.add_argument("--clusters", default="4,2,1")
    ap.add_argument("--batches", type=int, default=24)
    ap.add_argument("--prio", type=int, default=0)
    args = ap.parse_args()
    from trafficbots_b200 import engine as E, host, weights

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    W = bench.WORKLOAD
    S, A, P, K, T = W["n_scene"], W["n_agent"], W["n_pl"], W["n_mode"], W["n_step"]
    sd = weights.init_state_dict(2023)
    dmax = max(int(x) for x in args.depths.split(","))
    engs = [E.Engine(sd, dev) for _ in range(dmax)]
    cbs, outs = [], []
    for d in range(dmax):
        batch, _ = bench.make_inputs(S, A, P, K, seed=1000 + d)
        cbs.append(host.batch_to_device({k: batch[k] for k in bench.USED_KEYS}, dev))
        outs.append(engs[d].alloc_outputs(S * K, A, T))
    cex = {"latent_logp": torch.zeros(S * K, A, device=dev)}
    streams = [torch.cuda.Stream(dev) for _ in range(dmax)]
    for cl in [int(x) for x in args.clusters.split(",")]:
        os.environ["TB_CLUSTER"] = str(cl)
        for depth in [int(x) for x in args.depths.split(",")]:
            for eng in engs:
                eng._state = None  # scratch size depends on the cluster size
            n = args.batches
            for phase in range(2):  # warm-up pass, then the timed pass
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                lat = []
                t_host0 = time.perf_counter()
                e0.record()
                for s in streams[:depth]:
                    s.wait_event(e0)
                for i in range(n):
                    d = i % depth
                    with torch.cuda.stream(streams[d]):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        bench.run_step(engs[d], cbs[d], cex, K, T, outs[d])
                        b.record()
                        lat.append((a, b))
                t_host = time.perf_counter() - t_host0
                cur = torch.cuda.current_stream()
                for s in streams[:depth]:
                    cur.wait_stream(s)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            lats = sorted(a.elapsed_time(b) for a, b in lat)
            print(json.dumps({"cluster": cl, "depth": depth, "batches": n, "ms_per_batch": ms / n,
                              "scenes_per_s": S * n / (ms * 1e-3), "latency_ms_median": lats[len(lats) // 2],
                              "latency_ms_max": lats[-1], "host_issue_ms_per_batch": 1e3 * t_host / n}), flush=True)


if __name__ == "__main__":
    main()

"""How long does the HOST need to queue one batch of the pipelined scene step (BASELINE.json configs[1])?  If that is close to
the per-batch time of the bench, the schedule is bound by the Python thread issuing the launches, not by the GPU.
Usage: python tools/host_probe.py [depth]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from trafficbots_b200 import config as tb_config, host, weights  # noqa: E402
from trafficbots_b200.pipeline import ScenePipeline  # noqa: E402
from trafficbots_b200.pl_modules.waymo_motion import WaymoMotion  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    cfg = bench.CONFIGS[1]
    dev = torch.device("cuda", 0)
    module = WaymoMotion(**tb_config.default_config(n_joint_future=1))
    module.load_state_dict(weights.init_state_dict(2023))
    module = module.to(dev).eval()
    batch, _ = bench.make_inputs(cfg["n_scene"], cfg["n_agent"], cfg["n_pl"], 1, seed=1000)
    hb = host.pin_batch({k: batch[k] for k in bench.USED_KEYS})
    db = host.batch_to_device(hb, dev)
    for name, b, rb in (("device-resident batch, results stay on the device", db, False), ("pinned host batch in, results read back", hb, True)):
        pipe = ScenePipeline(module, depth=depth, read_back=rb)
        for _ in range(3):  # warm-up: three batches per slot
            tickets = [pipe.submit(b) for _ in range(depth)]
            for t in tickets:
                pipe.result(t)
        torch.cuda.synchronize()
        n = 40
        host_s = 0.0
        t_all = time.perf_counter()
        tickets = []
        for i in range(n):
            if i >= depth:
                pipe.result(tickets[i - depth])
            t0 = time.perf_counter()
            slot = pipe.slots[i % depth]
            slot.done.synchronize()  # exclude waiting for the slot from the host time of submit()
            t0 = time.perf_counter()
            tickets.append(pipe.submit(b))
            host_s += time.perf_counter() - t0
        for t in tickets[n - depth:]:
            pipe.result(t)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_all
        print(f"depth {depth}, {name}: host time to queue one batch {1e3 * host_s / n:.2f} ms; wall per batch {1e3 * wall / n:.2f} ms "
              f"({cfg['n_scene'] * n / wall:.0f} scenes/s)", flush=True)


if __name__ == "__main__":
    main()

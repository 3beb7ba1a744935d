"""times Engine.encode_scene (polyline encoder + map self-attention + K|V projections) on the bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from trafficbots_b200 import engine as E, host, weights

dev = torch.device("cuda", 0)
eng = E.Engine(weights.init_state_dict(2023), dev)
batch, _ = bench.make_inputs(32, 64, 1024, 1, seed=1000)
cb = host.batch_to_device({k: batch[k] for k in bench.USED_KEYS}, dev)
for _ in range(3):
    eng.encode_scene(cb)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.encode_scene(cb)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"encode_scene: median {sorted(ts)[5]:.3f} ms, min {min(ts):.3f} ms")

"""Micro-benchmarks of the training primitives at the shapes of one decode step of BASELINE.json configs[3]
(16 scenes x 64 agents = 1024 rows, 1024 map keys).  Back-to-back launches on one stream, CUDA events, per-launch average."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from trafficbots_b200.train.cuda_ops import CudaOps  # noqa: E402


def timeit(name, fn, n=200, flops=None, bytes_=None):
    """n dependent launches captured in one CUDA graph (the way the training step runs them): no host time in the number."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.2f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:58s} {us:9.2f} us{extra}", flush=True)


def ncu_pass():
    """one eager launch of the heavy training kernels at their largest shapes (for `ncu --set full -k regex:k_tr_`)."""
    dev = "cuda:0"
    ops = CudaOps(dev, check=False)
    torch.manual_seed(0)
    R = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    M = 327680
    x, w, b, dy = R(M, 128), R(128, 128), R(128), R(M, 128)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    for _ in range(2):
        y = ops.linear_fwd(x, w, b, False)
        ops.linear_bwd(dy, x, w, b, y, False, dw, db, True)
        lw, lb = R(128), R(128)
        yl, st = ops.layernorm_fwd(x, lw, lb, False)
        ops.layernorm_bwd(dy, x, lw, lb, st, yl, False, torch.zeros(128, device=dev), torch.zeros(128, device=dev))
        q, kv, do = R(16, 64, 128), R(16, 1024, 256), R(16, 64, 128)
        kvalid = (torch.rand(16, 1024, device=dev) < 0.9).to(torch.uint8)
        o, p, dead = ops.attention_fwd(q, kv, kvalid, False)
        ops.attention_bwd(do, q, kv, kvalid, False, p)
    torch.cuda.synchronize()


def main():
    if os.environ.get("TB_NCU"):
        return ncu_pass()
    dev = "cuda:0"
    ops = CudaOps(dev, check=False)
    torch.manual_seed(0)
    R = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    for M in (1024, 327680):
        x, w, b, dy = R(M, 128), R(128, 128), R(128), R(M, 128)
        y = ops.linear_fwd(x, w, b, False)
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        n = 200 if M == 1024 else 20
        timeit(f"linear_fwd  M={M} K=128 N=128", lambda: ops.linear_fwd(x, w, b, False), n, flops=2 * M * 128 * 128)
        timeit(f"linear_bwd  M={M} K=128 N=128 (dx+dw+db)", lambda: ops.linear_bwd(dy, x, w, b, y, False, dw, db, True), n,
               flops=4 * M * 128 * 128)
        w3, b3 = R(384, 128), R(384)
        timeit(f"linear_fwd  M={M} K=128 N=384", lambda: ops.linear_fwd(x, w3, b3, False), n, flops=2 * M * 128 * 384)
        lw, lb = R(128), R(128)
        yl, st = ops.layernorm_fwd(x, lw, lb, False)
        timeit(f"layernorm_fwd M={M}", lambda: ops.layernorm_fwd(x, lw, lb, False), n, bytes_=8 * M * 128)
        dlw, dlb = torch.zeros(128, device=dev), torch.zeros(128, device=dev)
        timeit(f"layernorm_bwd M={M}", lambda: ops.layernorm_bwd(dy, x, lw, lb, st, yl, False, dlw, dlb), n, bytes_=12 * M * 128)
        keep = (torch.rand(M, device=dev) < 0.9).to(torch.uint8)
        timeit(f"add_mask M={M}", lambda: ops.add_mask_fwd(x, dy, keep), n, bytes_=12 * M * 128)
        gi, gh = R(M, 384), R(M, 384)
        timeit(f"gru_gates_fwd M={M}", lambda: ops.gru_gates_fwd(gi, gh, x), n, bytes_=(8 * 384 + 8 * 128) * M)
        timeit(f"axpy M={M}", lambda: ops.add_(x, dy), n, bytes_=12 * M * 128)
    for (B, S, T, eye, what) in ((16, 64, 1024, False, "rollout agent->map"), (16, 64, 40, False, "rollout agent->TL"),
                                 (16, 64, 64, True, "rollout interaction"), (16384, 20, 20, False, "polyline encoder"),
                                 (16, 1216, 1024, False, "posterior latent agent->map"))[:int(os.environ.get("TB_N_ATT", "5"))]:
        q, kv, do = R(B, S, 128), R(B, T, 256), R(B, S, 128)
        kvalid = (torch.rand(B, T, device=dev) < 0.9).to(torch.uint8)
        o, p, dead = ops.attention_fwd(q, kv, kvalid, eye)
        n = 100 if B * S * T < 4e6 else 10
        fl = 4.0 * B * S * T * 128
        timeit(f"attention_fwd B={B} S={S} T={T} ({what})", lambda: ops.attention_fwd(q, kv, kvalid, eye), n, flops=fl)
        timeit(f"attention_bwd B={B} S={S} T={T} ({what})", lambda: ops.attention_bwd(do, q, kv, kvalid, eye, p), n, flops=2 * fl)


if __name__ == "__main__":
    main()

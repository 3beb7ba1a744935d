"""Times the training step (BASELINE.json configs[3] per-GPU slice: 16 scenes x 64 agents x 1024 polylines, 90 steps)
phase by phase.  Usage: python tools/bench_train.py [n_scene] [n_agent] [n_pl] [repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from trafficbots_b200 import synthetic, weights  # noqa: E402
from trafficbots_b200.train import graph, trainer  # noqa: E402
from trafficbots_b200.train.tape import Fn  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    P = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    dev = "cuda:0"
    sd = weights.init_state_dict(2023)
    ts = trainer.TrainState(sd, device=dev)
    ts.ops.check = False
    batch = {k: v.to(dev) for k, v in synthetic.make_batch(S, n_agent=A, n_pl=P, seed=7).items()}
    torch.manual_seed(0)
    L = ts.ops.L
    for i in range(rep + 1):
        use_prior, eps = ts.draw_noise(S, A)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        n0 = L.tb_launch_count()
        t0 = time.perf_counter()
        ts.flat_g.zero_()
        fn = Fn(ts.ops)
        out = graph.training_forward(fn, ts.params, batch, eps.to(dev), use_prior, n_step=int(os.environ.get("TB_TRAIN_STEPS", "90")))
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        n1 = L.tb_launch_count()
        fn.backward()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        n2 = L.tb_launch_count()
        ts.optimizer_step()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        print(f"iter {i}: forward host {1e3 * (t1 - t0):.0f} ms (+{1e3 * (t2 - t1):.0f} ms GPU tail, {n1 - n0} launches, {fn.n_fwd} ops), "
              f"backward host {1e3 * (t3 - t2):.0f} ms (+{1e3 * (t4 - t3):.0f} ms GPU tail, {n2 - n1} launches), adam {1e3 * (t5 - t4):.1f} ms, "
              f"total {1e3 * (t5 - t0):.0f} ms = {S / (t5 - t0):.1f} scenes/s, peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB, "
              f"loss {float(out['loss']):.4f}", flush=True)


def graph_mode():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    P = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    dev = "cuda:0"
    ts = trainer.TrainState(weights.init_state_dict(2023), device=dev, dropout_p=float(os.environ.get("TB_TRAIN_DROPOUT", "0")),
                            n_split=int(os.environ.get("TB_TRAIN_SPLIT", "1")))
    ts.ops.check = False
    batch = {k: v.to(dev) for k, v in synthetic.make_batch(S, n_agent=A, n_pl=P, seed=7).items()}
    torch.manual_seed(0)
    t0 = time.perf_counter()
    ts.capture(batch)
    eager = ts.forward_backward(batch, torch.zeros(S, A, 16), False)
    g_eager = ts.flat_g.clone()
    out = ts.replay(batch, torch.zeros(S, A, 16), False)
    torch.cuda.synchronize()
    print(f"n_split {ts.n_split}, dropout {ts.dropout_p}: capture (posterior graph): {time.perf_counter() - t0:.1f} s; graph == eager: loss {float(out['loss']):.6f} vs "
          f"{float(eager['loss']):.6f}, max grad diff {float((ts.flat_g - g_eager).abs().max()):.2e} "
          f"(scale {float(g_eager.abs().max()):.2e}); memory reserved {torch.cuda.memory_reserved() / 2 ** 30:.1f} GiB", flush=True)
    for i in range(rep):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        eps = torch.randn(S, A, 16)
        e0.record()
        out = ts.replay(batch, eps, False)
        e1.record()
        ts.optimizer_step()
        e2.record()
        torch.cuda.synchronize()
        print(f"graph iter {i}: fwd+bwd {e0.elapsed_time(e1):.1f} ms, adam {e1.elapsed_time(e2):.2f} ms -> {S / (e0.elapsed_time(e2) * 1e-3):.1f} "
              f"scenes/s, loss {float(out['loss']):.4f}", flush=True)


def curve_mode(n_iter):
    """n_iter optimizer steps (graph replay, dropout 0.1, lr 3e-4, clip 5) on two alternating synthetic batches: the loss terms."""
    S, A, P = 16, 64, 1024
    dev = "cuda:0"
    ts = trainer.TrainState(weights.init_state_dict(2023), device=dev, dropout_p=0.1)
    ts.ops.check = False
    batches = [{k: v.to(dev) for k, v in synthetic.make_batch(S, n_agent=A, n_pl=P, seed=7 + i).items()} for i in range(2)]
    torch.manual_seed(0)
    ts.capture(batches[0])
    hist = []
    for i in range(n_iter):
        out = ts.training_step(batches[i % 2], graph=True)
        hist.append([float(out[k]) for k in ("loss", "vae_kl", "diffbar_reward", "goal_loss")] + [float(out["grad_sq_norm"]) ** 0.5])
    for i, h in enumerate(hist):
        if i < 3 or i % 10 == 9:
            print(f"step {i + 1:3d}: loss {h[0]:.4f} = kl {h[1]:.4f} + reward {h[2]:.4f} + goal {h[3]:.4f}; |grad| {h[4]:.2f}", flush=True)
    first, last = sum(h[0] for h in hist[:10]) / 10, sum(h[0] for h in hist[-10:]) / 10
    print(f"mean loss of the first 10 steps {first:.4f}, of the last 10 steps {last:.4f}")


if __name__ == "__main__":
    if os.environ.get("TB_TRAIN_CURVE"):
        curve_mode(int(os.environ["TB_TRAIN_CURVE"]))
        sys.exit(0)
    if os.environ.get("TB_TRAIN_GRAPH"):
        graph_mode()
        sys.exit(0)
    main()

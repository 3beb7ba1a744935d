"""Helpers of the training tests: load a gradient fixture (oracle/make_golden_train.py) and compare a gradient set with it."""
import os

import numpy as np
import torch

import ref_train
from make_golden_train import N_PROJ, fingerprint
from trafficbots_b200 import synthetic, weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN_CASES = ("train_s2_a8_p64_post", "train_s3_a8_p64_prior")


def load_train_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    S, A, P, seed, wseed, nseed = [int(x) for x in z["meta__case"]]
    p_prior = float(z["meta__p_prior"])
    sd = weights.init_state_dict(wseed)
    batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed)
    use_prior, eps = ref_train.draw_training_noise(nseed, S, A, p_prior=p_prior)
    terms = {k[len("term__"):]: float(z[k]) for k in z.files if k.startswith("term__")}
    grads = {k[len("grad__"):]: z[k] for k in z.files if k.startswith("grad__")}
    return dict(S=S, A=A, P=P, sd=sd, batch=batch, use_prior=use_prior, eps=eps, terms=terms, grads=grads,
                latent_sample=torch.from_numpy(z["latent_sample"]))


def compare_grads(grads, gold, rel=1e-3):
    """every parameter's gradient against the reference fingerprint: norm, projections and sampled entries within
    `rel` of the tensor's scale (its L2 norm).  Returns the worst relative deviation."""
    worst = 0.0
    total_sq = sum(float(fp[0]) ** 2 for fp in gold.values())
    floor = 1e-6 * total_sq ** 0.5  # gradients that are zero up to rounding (e.g. the last bias of the destination MLP)
    for name, g in grads.items():
        if name not in gold:
            assert float(g.abs().max()) == 0.0, f"{name}: the reference has no gradient here"
            continue
        fp = fingerprint(name, g)
        ref = gold[name]
        scale = float(ref[0]) + floor
        n = g.numel()
        dev_norm = abs(fp[0] - ref[0]) / scale
        dev_proj = np.abs(fp[1:1 + N_PROJ] - ref[1:1 + N_PROJ]).max() / (scale * 4.0)  # ~N(0, norm^2) projections
        dev_smp = np.abs(fp[1 + N_PROJ:] - ref[1 + N_PROJ:]).max() / (scale / n ** 0.5 * 8.0 + floor)
        dev = max(dev_norm, dev_proj, dev_smp)
        assert dev <= rel, f"{name}: gradient deviates from the reference by {dev:.2e} of its scale (norm {ref[0]:.3e})"
        worst = max(worst, dev)
    return worst

"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/train_*.npz` from the UNMODIFIED reference's `training_step` + backward.

Run in the build container (needs `/root/reference`):   python oracle/make_golden_train.py

Per case: seeded synthetic scenes + seeded parameters are loaded into the reference `WaymoMotion` (train mode, every dropout
probability 0), `training_step(batch, 0)` is executed and back-propagated (`oracle/ref_train.py`).  Stored: the loss terms,
and per parameter a gradient FINGERPRINT (the full gradients would be 13.6 MB per case): L2 norm, projections on four seeded
random directions and 32 seeded sample entries.  `fingerprint()` is shared with the tests.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

CASES = {
    # name: (n_scene, n_agent, n_pl, scene_seed, weight_seed, noise_seed, p_prior)
    "train_s2_a8_p64_post": (2, 8, 64, 100, 2023, 5, 0.1),    # posterior latent in the rollout (the 90 % branch, :386-387)
    "train_s3_a8_p64_prior": (3, 8, 64, 100, 2023, 6, 1.0),   # prior latent in the rollout (:384-385); zero-TL + single-agent scenes
}
N_PROJ, N_SAMPLE = 4, 32


def fingerprint(name: str, g: torch.Tensor) -> np.ndarray:
    """[norm, N_PROJ projections, N_SAMPLE entries] of a gradient tensor (fp64 accumulate), seeded by the parameter name."""
    g = g.detach().double().cpu().reshape(-1)
    gen = torch.Generator().manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(name)) % (2 ** 31))
    proj = torch.randn(N_PROJ, g.numel(), generator=gen, dtype=torch.float64) @ g
    idx = torch.randint(0, g.numel(), (N_SAMPLE,), generator=gen)
    return torch.cat([g.norm().reshape(1), proj, g[idx]]).numpy()


def main() -> None:
    import ref_loader
    import ref_train
    from trafficbots_b200 import synthetic, weights
    from make_golden import checksum
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (S, A, P, seed, wseed, nseed, p_prior) in CASES.items():
        model = ref_loader.build_reference(n_agent=A, n_pl=P, n_joint_future=1)
        sd = weights.init_state_dict(wseed)
        model.load_state_dict(sd, strict=True)
        batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed)
        terms, grads, extra = ref_train.run_reference_training(model, batch, nseed, p_prior)
        arrays = {"term__" + k: np.array(float(v)) for k, v in terms.items()}
        for k, g in grads.items():
            if g is not None:
                arrays["grad__" + k] = fingerprint(k, g)
        arrays["latent_sample"] = extra["latent_sample"].numpy()
        arrays["meta__case"] = np.array([S, A, P, seed, wseed, nseed], dtype=np.int64)
        arrays["meta__p_prior"] = np.array(p_prior)
        arrays["meta__checksum_batch"] = np.array(checksum(batch))
        arrays["meta__checksum_weights"] = np.array(checksum(sd))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", {k: float(v) for k, v in terms.items()})


if __name__ == "__main__":
    main()

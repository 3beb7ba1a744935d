"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's `training_step` + backward on a synthetic batch
(build container only; reference `src/pl_modules/waymo_motion.py:356-418`, Lightning's `loss.backward()`).

Every dropout probability is set to 0 (BASELINE.json configs[3]: "dropout 0 for the parity run"); the module stays in
train() mode.  RNG protocol (global torch CPU generator, seeded by the caller): `torch.rand(1)` decides prior vs posterior
(:384), then `Normal.rsample` draws the latent noise [S, A, 16] at the first decode step (models/traffic_bots.py:196-199).
`draw_training_noise` reproduces the two draws for the implementation under test.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import nn


def zero_dropout(model: nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.GRU):
            m.dropout = 0.0


def draw_training_noise(seed: int, n_scene: int, n_agent: int, latent_dim: int = 16, p_prior: float = 0.1) -> Tuple[bool, torch.Tensor]:
    torch.manual_seed(seed)
    use_prior = bool(torch.rand(1) < p_prior)
    eps = torch.empty(n_scene, n_agent, latent_dim).normal_()
    return use_prior, eps


def run_reference_training(model, batch: Dict[str, torch.Tensor], seed: int, p_prior: float = 0.1, dropout: bool = False):
    """-> (loss terms dict, gradients by `named_parameters()` name, extra tensors).  dropout=False: every dropout probability
    set to 0 (the parity configuration); True: the module as shipped (p = 0.1 everywhere; used for timing only)."""
    model.train()
    if not dropout:
        zero_dropout(model)
    model.hparams.p_training_rollout_prior = p_prior
    for p in model.parameters():
        p.grad = None
    logged = {}
    model.log = lambda k, v, **kw: logged.__setitem__(k, v.detach().clone() if torch.is_tensor(v) else v)
    torch.manual_seed(seed)
    loss = model.training_step({k: v.clone() for k, v in batch.items()}, 0)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    terms = {k.split("/")[-1]: v for k, v in logged.items()}
    extra = {"latent_sample": model.model.latent_sample.detach().clone()}
    return terms, grads, extra

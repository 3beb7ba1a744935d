"""Distribution objects handed between the pre-rollout heads and the rollout: the diagonal Gaussian personality
latent and the categorical destination (interface of reference `src/models/modules/distributions.py:9-59,150-201`:
`.sample(deterministic)`, `.log_prob(x)`, `.repeat_interleave_(k, dim)`, `.mean / .stddev / .probs / .valid`).
Sampling is host-side bookkeeping on tiny tensors with torch's generator, like the reference."""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
from torch import Tensor


class DiagGaussian:
    def __init__(self, mean: Tensor, log_std: Tensor, valid: Optional[Tensor] = None) -> None:
        self.mean = mean
        self.stddev = log_std.exp().expand_as(mean)
        self.valid = valid

    def repeat_interleave_(self, repeats: int, dim: int) -> None:
        self.mean = self.mean.repeat_interleave(repeats, dim)
        self.stddev = self.stddev.repeat_interleave(repeats, dim)
        if self.valid is not None:
            self.valid = self.valid.repeat_interleave(repeats, dim)

    def _rsample(self) -> Tensor:
        # the reference's draw (`Independent(Normal(mean, std), 1).rsample()`, distributions.py:26,49): same generator calls
        return torch.distributions.Normal(self.mean, self.stddev, validate_args=False).rsample()

    def sample(self, deterministic: Union[bool, Tensor]) -> Tensor:
        if isinstance(deterministic, Tensor):
            return torch.where(deterministic.unsqueeze(-1), self.mean, self._rsample())
        return self.mean if deterministic else self._rsample()

    def log_prob(self, sample: Tensor) -> Tensor:
        var = self.stddev ** 2
        return (-((sample - self.mean) ** 2) / (2 * var) - self.stddev.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)


class DestCategorical:
    def __init__(self, probs: Optional[Tensor] = None, logits: Optional[Tensor] = None, valid: Optional[Tensor] = None):
        if probs is None:
            probs = torch.softmax(logits, dim=-1)
        self.probs = probs / probs.sum(-1, keepdim=True)
        self.valid = valid

    def repeat_interleave_(self, repeats: int, dim: int) -> None:
        self.probs = self.probs.repeat_interleave(repeats, dim)
        if self.valid is not None:
            self.valid = self.valid.repeat_interleave(repeats, dim)

    def sample(self, deterministic: Union[bool, Tensor]) -> Tensor:
        det = self.probs.argmax(-1)
        if deterministic is True:
            return det
        rnd = torch.multinomial(self.probs.reshape(-1, self.probs.shape[-1]), 1, True).reshape(self.probs.shape[:-1])
        if deterministic is False:
            return rnd
        return torch.where(deterministic, det, rnd)

    def log_prob(self, sample: Tensor) -> Tensor:
        eps = torch.finfo(self.probs.dtype).eps
        return torch.log(self.probs.clamp(min=eps, max=1 - eps)).gather(-1, sample.unsqueeze(-1)).squeeze(-1)

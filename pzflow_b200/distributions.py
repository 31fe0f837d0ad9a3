"""Latent distributions of the hot path: Uniform and CentBeta13.

Mirrors pzflow/distributions.py (LatentDist 17-30, CentBeta13 137-229, Uniform
395-470): same constructors, ``info`` tuples and ``_params``.  Inside
``Flow._log_prob`` the latent log-density is the tail of the fused chain
kernel; the stand-alone ``log_prob`` / ``sample`` below are the thin device
versions of the reference methods (they are not on the hot path).
CentBeta, Normal, Tdist and Joint are out of scope.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import plan as _plan

Pytree = tuple


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _plan.N.NativeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _seed(seed) -> int:
    return int(np.random.randint(1e18)) if seed is None else int(seed)


class LatentDist(ABC):
    """Base class for latent distributions."""

    info = ("LatentDist", ())

    @abstractmethod
    def log_prob(self, params: Pytree, inputs):
        """Calculate log-probability of the inputs."""

    @abstractmethod
    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """Sample from the distribution."""


class CentBeta13(LatentDist):
    """A centered Beta distribution with alpha, beta = 13 (support [-B, B] per dimension)."""

    def __init__(self, input_dim: int, B: float = 5) -> None:
        self.input_dim = input_dim
        self.B = B
        self._params = tuple([(0.0, 0.0) for i in range(input_dim)])
        self.info = ("CentBeta13", (input_dim, B))
        self.a = 13
        self.b = 13

    def log_prob(self, params: Pytree, inputs):
        """sum_i beta.logpdf(u_i; 13, 13, loc=-B, scale=2B) (distributions.py:165-194)."""
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        loc, scale = -float(self.B), 2.0 * float(self.B)
        betaln = 2.0 * math.lgamma(13.0) - math.lgamma(26.0)
        y = (u - loc) / scale
        lp = (-betaln + (12.0 * torch.log(y) + 12.0 * torch.log1p(-y))) - math.log(scale)
        lp = torch.where((u > loc + scale) | (u < loc), torch.full_like(lp, -math.inf), lp)
        out = lp.sum(dim=1)
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None) -> torch.Tensor:
        gen = torch.Generator(device=_device())
        gen.manual_seed(_seed(seed) % (2 ** 63))
        conc = torch.full((nsamples, self.input_dim, 2), 13.0, device=_device())
        g = torch._standard_gamma(conc, generator=gen)
        s = g[..., 0] / (g[..., 0] + g[..., 1])
        return (2.0 * float(self.B) * (s - 0.5)).contiguous()

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """[nsamples, input_dim] draws, 2B(Beta(13,13) - 0.5) (distributions.py:196-229)."""
        return self.sample_device(nsamples, seed).cpu().numpy()


class Uniform(LatentDist):
    """A multivariate uniform distribution with support [-B, B]."""

    def __init__(self, input_dim: int, B: float = 5) -> None:
        self.input_dim = input_dim
        self.B = B
        self._params = ()
        self.info = ("Uniform", (input_dim, B))

    def log_prob(self, params: Pytree, inputs):
        """-D log(2B) inside the support, -inf outside (distributions.py:414-441)."""
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        mask = ((u >= -self.B) & (u <= self.B)).all(dim=-1)
        val = -self.input_dim * math.log(2 * self.B)
        out = torch.where(mask, torch.full(mask.shape, val, device=u.device),
                          torch.full(mask.shape, -math.inf, device=u.device))
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None) -> torch.Tensor:
        return _plan.sample_uniform(nsamples, self.input_dim, float(self.B), _seed(seed), _device())

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """[nsamples, input_dim] draws from U(-B, B) (distributions.py:443-470)."""
        return self.sample_device(nsamples, seed).cpu().numpy()

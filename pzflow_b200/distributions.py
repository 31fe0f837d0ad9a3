"""Latent distributions of the hot path: Uniform, CentBeta13, and Normal, CentBeta, Tdist, Joint.

Mirrors pzflow/distributions.py (LatentDist 17-30, CentBeta 45-135, CentBeta13 137-229, Normal
232-303, Uniform 395-470): same constructors, ``info`` tuples and ``_params``.  Inside
``Flow._log_prob`` the latent log-density is the tail of the fused chain
kernel; the stand-alone ``log_prob`` / ``sample`` below are the thin device
versions of the reference methods (they are not on the hot path).
All of the reference's latents are covered (Tdist 306-392 and Joint 473-562 included).
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import plan as _plan
from . import prng as _prng

Pytree = tuple


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _plan.N.NativeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _seed(seed) -> int:
    return int(np.random.randint(1e18)) if seed is None else int(seed)


def _mahalanobis_and_logdet(x, cov):
    """Mahalanobis distance of x [N, D] and log-determinant of ``cov`` exactly as pzflow/distributions.py:33-42 writes
    them (eigen-decomposition; the only caller passes the identity, distributions.py:349-350).  Not on the hot path: the
    Tdist log-density inside the fused kernels uses |u|^2 and 0."""
    x = torch.as_tensor(x).to(_device(), torch.float32)
    vals, vecs = torch.linalg.eigh(torch.as_tensor(np.asarray(cov, dtype=np.float32)).to(x.device))
    U = vecs * torch.sqrt(1 / vals[..., None])
    maha = torch.square(U @ x[..., None]).reshape(x.shape[0], -1).sum(dim=1)
    return maha, torch.log(vals).sum(dim=-1)


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

    def sample_device(self, nsamples: int, seed: int = None, *, device=None, row_offset: int = 0,
                      total_rows: int = None) -> torch.Tensor:
        """Draws on ``device`` (default: the current one); ``row_offset`` = index of the first row in the caller's whole
        sample, so that row shards drawn on several devices equal one big draw."""
        shapes = np.full(self.input_dim, 13.0, np.float32)
        return _plan.sample_centbeta(nsamples, shapes, shapes, float(self.B), _seed(seed), _device(device), row_offset)

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

    def sample_device(self, nsamples: int, seed: int = None, *, device=None, row_offset: int = 0,
                      total_rows: int = None) -> torch.Tensor:
        """Rows [row_offset, row_offset + nsamples) of the ``total_rows`` draws of ``seed``, on ``device``.  In the
        default stream layout these are jax's own numbers -- ``random.uniform(PRNGKey(seed), (total_rows, D), -B, B)``
        (distributions.py:464-469) -- so ``Flow.sample(n, seed=s)`` returns the reference's rows (pzflow_b200.prng)."""
        total = nsamples + row_offset if total_rows is None else int(total_rows)
        lay = _prng.layout()
        if lay == "philox" or (lay == "threefry" and total * self.input_dim > _prng.MAX_THREEFRY_WORDS):
            return _plan.sample_uniform(nsamples, self.input_dim, float(self.B), _seed(seed), _device(device), row_offset)
        return _plan.sample_uniform_jax(nsamples, self.input_dim, float(self.B), _seed(seed), _device(device), row_offset,
                                        total, lay)

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """[nsamples, input_dim] draws from U(-B, B) (distributions.py:443-470)."""
        return self.sample_device(nsamples, seed).cpu().numpy()


class Normal(LatentDist):
    """A multivariate Gaussian distribution with mean zero and unit variance."""

    def __init__(self, input_dim: int) -> None:
        self.input_dim = input_dim
        self._params = ()
        self.info = ("Normal", (input_dim,))

    def log_prob(self, params: Pytree, inputs):
        """multivariate_normal.logpdf(inputs, 0, I) (distributions.py:255-275)."""
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        out = -0.5 * (self.input_dim * math.log(2.0 * math.pi) + (u * u).sum(dim=1))
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None, *, device=None, row_offset: int = 0,
                      total_rows: int = None) -> torch.Tensor:
        """Draws on the device.  In the default stream layout these are jax's own numbers: the reference's
        ``random.multivariate_normal(PRNGKey(seed), 0, I, (nsamples,))`` (pzflow/distributions.py:297-303) is
        ``random.normal(key, (nsamples, D))`` pushed through the identity Cholesky factor."""
        total = row_offset + nsamples if total_rows is None else int(total_rows)
        lay = _prng.layout()
        if lay == "philox" or (lay == "threefry" and total * self.input_dim > _prng.MAX_THREEFRY_WORDS):
            return _plan.sample_normal(nsamples, self.input_dim, _seed(seed), _device(device), row_offset)
        return _plan.sample_normal_jax(nsamples, self.input_dim, _seed(seed), _device(device), row_offset, total, lay)

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """[nsamples, input_dim] standard normal draws (distributions.py:277-303)."""
        return self.sample_device(nsamples, seed).cpu().numpy()


class CentBeta(LatentDist):
    """A centered Beta distribution with learned alpha, beta per dimension (support [-B, B])."""

    def __init__(self, input_dim: int, B: float = 5) -> None:
        self.input_dim = input_dim
        self.B = B
        self._params = tuple([(0.0, 0.0) for i in range(input_dim)])
        self.info = ("CentBeta", (input_dim, B))

    @staticmethod
    def _ab(params: Pytree, dev):
        a = torch.tensor([math.exp(float(np.asarray(p[0]))) for p in params], dtype=torch.float32, device=dev)
        b = torch.tensor([math.exp(float(np.asarray(p[1]))) for p in params], dtype=torch.float32, device=dev)
        return a, b

    def log_prob(self, params: Pytree, inputs):
        """sum_i beta.logpdf(u_i; exp(p_i0), exp(p_i1), loc=-B, scale=2B) (distributions.py:70-99)."""
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        a, b = self._ab(params, u.device)
        loc, scale = -float(self.B), 2.0 * float(self.B)
        betaln = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
        y = (u - loc) / scale
        lp = (-betaln + (torch.xlogy(a - 1, y) + torch.special.xlog1py(b - 1, -y))) - math.log(scale)
        lp = torch.where((u > loc + scale) | (u < loc), torch.full_like(lp, -math.inf), lp)
        out = lp.sum(dim=1)
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None, params: Pytree = None, *, device=None,
                      row_offset: int = 0, total_rows: int = None) -> torch.Tensor:
        params = self._params if params is None else params
        a = np.exp(np.asarray([float(np.asarray(p[0])) for p in params], np.float64)).astype(np.float32)
        b = np.exp(np.asarray([float(np.asarray(p[1])) for p in params], np.float64)).astype(np.float32)
        return _plan.sample_centbeta(nsamples, a, b, float(self.B), _seed(seed), _device(device), row_offset)

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        """[nsamples, input_dim] draws, 2B(Beta(a_i, b_i) - 0.5) (distributions.py:101-135)."""
        return self.sample_device(nsamples, seed, params).cpu().numpy()


class Tdist(LatentDist):
    """A multivariate T distribution with mean zero, unit scale matrix and learned degrees of freedom."""

    def __init__(self, input_dim: int) -> None:
        self.input_dim = input_dim
        self._params = float(np.log(np.float32(30.0)))
        self.info = ("Tdist", (input_dim,))

    def log_prob(self, params: Pytree, inputs):
        """distributions.py:330-358 with cov = I (maha = |u|^2, log_det = 0)."""
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        nu = math.exp(float(np.asarray(params)))
        t = 0.5 * (nu + self.input_dim)
        out = (math.lgamma(t) - math.lgamma(0.5 * nu) - self.input_dim / 2.0 * math.log(nu * math.pi)
               - t * torch.log(1 + (1.0 / nu) * (u * u).sum(dim=1)))
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None, params: Pytree = None, *, device=None,
                      row_offset: int = 0, total_rows: int = None) -> torch.Tensor:
        """mean + z / sqrt(chi2(nu) / nu) (distributions.py:360-392)."""
        nu = math.exp(float(np.asarray(self._params if params is None else params)))
        return _plan.sample_tdist(nsamples, self.input_dim, nu, _seed(seed), _device(device), row_offset)

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        return self.sample_device(nsamples, seed, params).cpu().numpy()


class Joint(LatentDist):
    """A joint distribution built from other distributions over consecutive columns (distributions.py:473-562)."""

    def __init__(self, *inputs) -> None:
        if inputs[0] == "Joint info":
            self.dists = [globals()[dist[0]](*dist[1]) for dist in inputs[1]]
        else:
            self.dists = inputs
        self._params = [dist._params for dist in self.dists]
        self.input_dim = sum([dist.input_dim for dist in self.dists])
        self.info = ("Joint", ("Joint info", [dist.info for dist in self.dists]))
        self._splits = np.cumsum([dist.input_dim for dist in self.dists])[:-1]

    def log_prob(self, params: Pytree, inputs):
        u = torch.as_tensor(inputs).to(_device(), torch.float32)
        parts = torch.tensor_split(u, [int(i) for i in self._splits], dim=1)
        out = torch.stack([self.dists[i].log_prob(params[i], parts[i]) for i in range(len(self.dists))], dim=1).sum(dim=1)
        return out if isinstance(inputs, torch.Tensor) else out.cpu().numpy()

    def sample_device(self, nsamples: int, seed: int = None, params: Pytree = None, *, device=None,
                      row_offset: int = 0, total_rows: int = None) -> torch.Tensor:
        params = self._params if params is None else params
        seeds = np.random.default_rng(_seed(seed)).integers(0, int(1e9), size=len(self.dists))
        cols = []
        kw = dict(device=device, row_offset=row_offset, total_rows=total_rows)
        for dist, p, sd in zip(self.dists, params, seeds):
            if isinstance(dist, (CentBeta, Tdist, Joint)):
                cols.append(dist.sample_device(nsamples, int(sd), p, **kw).reshape(nsamples, -1))
            else:
                cols.append(dist.sample_device(nsamples, int(sd), **kw).reshape(nsamples, -1))
        return torch.cat(cols, dim=1).contiguous()

    def sample(self, params: Pytree, nsamples: int, seed: int = None):
        return self.sample_device(nsamples, seed, params).cpu().numpy()

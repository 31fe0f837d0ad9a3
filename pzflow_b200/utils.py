"""Host glue of the hot path (mirror of pzflow/utils.py:11-19)."""
from __future__ import annotations

from . import bijectors


def build_bijector_from_info(info: tuple) -> tuple:
    """Build a Bijector from a Bijector_Info object (pzflow/utils.py:11-19)."""
    if info[0] == "Chain":
        return bijectors.Chain(*(build_bijector_from_info(i) for i in info[1]))
    try:
        factory = getattr(bijectors, info[0])
    except AttributeError:
        raise NotImplementedError(f"bijector {info[0]!r} is outside the B200 hot path") from None
    return factory(*info[1])


def gaussian_error_model(key, X, Xerr, nsamples):
    """Default Gaussian error model (pzflow/utils.py:52-61): X[:, None, :] + eps * Xerr[:, None, :],
    eps ~ N(0, 1) of shape [N, nsamples, D].  ``key`` is the integer seed here (the reference passes a
    jax PRNGKey); the draws are the library's counter-based stream (pzflow_b200.plan.error_eps), the same
    ones the fused kernels generate on the fly -- Flow.log_prob / posterior never call this function, they
    only check that the flow's error model IS this one and let the kernel do the expansion."""
    import numpy as np
    import torch

    from .plan import error_eps
    X = np.asarray(X, dtype=np.float32)
    Xerr = np.asarray(Xerr, dtype=np.float32)
    k = np.asarray(key).astype(np.uint64).ravel()   # an int seed, or the uint32 pair PRNGKey(seed) would be
    seed = int(k[0]) if k.size == 1 else (int(k[0]) << 32) | int(k[1])
    eps = error_eps(X.shape[0] * int(nsamples), X.shape[1], 0, seed).cpu().numpy()
    eps = eps.reshape(X.shape[0], int(nsamples), X.shape[1])
    return X[:, None, :] + eps * Xerr[:, None, :]

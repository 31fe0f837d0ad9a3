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
    """Placeholder so pickles that reference pzflow.utils.gaussian_error_model load;
    error convolution (pzflow/utils.py:52-61) is a "next" row."""
    raise NotImplementedError("err_samples / Gaussian error convolution is not part of the B200 hot path yet")

"""Host-side mirror of pzflow's Bijector protocol for the hot-path bijectors.

Same names, arguments, ``Bijector_Info`` tuples, params pytrees and error
behaviour as pzflow/bijectors.py (protocol 16-118, Chain 121-174, ColorTransform 177-311,
InvSoftplus 314-372, NeuralSplineCoupling 375-522, Reverse 525-556, Roll 559-598,
RollingSplineCoupling 601-665, Scale 668-714, ShiftBounds 717-777,
StandardScaler 817-861) -- but the Forward/Inverse
functions an ``InitFunction`` returns do not run stage by stage: each one
compiles the WHOLE bijector it belongs to (nested Chains included) into one
device plan and evaluates it with one fused CUDA kernel through the C-ABI.

Shuffle (780-814) draws its permutation from the jax key its InitFunction receives, and stores it nowhere: the key
schedule of the reference (PRNGKey(seed), one ``random.split`` per Chain stage, ``random.permutation``) is reproduced
on the host (pzflow_b200.prng, plan.resolve_shuffles), so a flow built with the same ``seed`` gets the reference's
permutation; like ``Roll`` it then costs nothing (folded into the column maps of the plan).  UniformDequantizer
(865-911) is provided AS THE REFERENCE EXECUTES IT: its forward map is the identity (line 897 discards the result of
``.at[].set()``), its inverse floors the listed columns.
"""
from __future__ import annotations

from functools import update_wrapper
from typing import Callable, Sequence, Tuple, Union

import numpy as np
import torch

from . import prng
from .plan import Plan, PlanCache

Pytree = Union[tuple, list]
Bijector_Info = Tuple[str, tuple]


class ForwardFunction:
    """(params, inputs[N,D], conditions=[N,C]) -> (outputs[N,D], log_det[N]); bijectors.py:16-44."""

    def __init__(self, func: Callable) -> None:
        self._func = func

    def __call__(self, params: Pytree, inputs, **kwargs):
        return self._func(params, inputs, **kwargs)


class InverseFunction:
    """(params, inputs[N,D], conditions=[N,C]) -> (outputs[N,D], log_det[N]); bijectors.py:47-75."""

    def __init__(self, func: Callable) -> None:
        self._func = func

    def __call__(self, params: Pytree, inputs, **kwargs):
        return self._func(params, inputs, **kwargs)


class InitFunction:
    """(rng, input_dim) -> (params, ForwardFunction, InverseFunction); bijectors.py:78-107."""

    def __init__(self, func: Callable) -> None:
        self._func = func

    def __call__(self, rng, input_dim: int, **kwargs):
        return self._func(rng, input_dim, **kwargs)


class Bijector:
    """Wrapper class for bijector functions; bijectors.py:110-118."""

    def __init__(self, func: Callable) -> None:
        self._func = func
        update_wrapper(self, func)

    def __call__(self, *args, **kwargs) -> Tuple[InitFunction, Bijector_Info]:
        return self._func(*args, **kwargs)


# --------------------------------------------------------------------------- #
def _generator(rng) -> np.random.Generator:
    """Accept an int seed, a NumPy Generator/SeedSequence or a jax-style uint32 key."""
    if isinstance(rng, np.random.Generator):
        return rng
    if isinstance(rng, np.random.SeedSequence):
        return np.random.default_rng(rng)
    if rng is None:
        return np.random.default_rng()
    arr = np.asarray(rng).astype(np.uint64).ravel()
    return np.random.default_rng(np.random.SeedSequence([int(v) for v in arr]))


def _jax_key(rng):
    """The jax-style key an InitFunction's ``rng`` stands for: an int seed is PRNGKey(seed), a uint32 pair is a key;
    NumPy generators (which no jax program can hold) get a key derived from their state."""
    if rng is None:
        return None
    if isinstance(rng, (int, np.integer)):
        return prng.key_from_seed(int(rng))
    if isinstance(rng, np.random.Generator):
        return None  # a nested stage of a Chain: only the outermost InitFunction resolves Shuffles
    if isinstance(rng, np.random.SeedSequence):
        return np.asarray(rng.generate_state(2), dtype=np.uint32)
    arr = np.asarray(rng)
    if arr.shape == (2,):
        return arr.astype(np.uint32)
    return None


def _split(rng: np.random.Generator) -> Tuple[np.random.Generator, np.random.Generator]:
    a, b = rng.bit_generator.seed_seq.spawn(2) if hasattr(rng.bit_generator, "seed_seq") else \
        np.random.SeedSequence(int(rng.integers(2 ** 63))).spawn(2)
    return np.random.default_rng(a), np.random.default_rng(b)


class _FusedFunctions:
    """Forward/inverse of one Bijector_Info, fused into one plan per params identity."""

    _MAX_PLANS = 4

    def __init__(self, info: Bijector_Info, input_dim: int, key=None) -> None:
        self.info = info
        self.input_dim = input_dim
        self.key = key  # the jax-style key the InitFunction was called with (Shuffle's permutation hangs on it)
        self._plans = PlanCache(self._MAX_PLANS)

    def plan(self, params: Pytree) -> Plan:
        device = torch.cuda.current_device() if torch.cuda.is_available() else -1
        return self._plans.get(params, device, lambda: Plan(self.info, params, self.input_dim, rng_key=self.key))

    @staticmethod
    def _out(t: torch.Tensor, like):
        return t if isinstance(like, torch.Tensor) else t.cpu().numpy()

    def forward(self, params, inputs, conditions=None, **kwargs):
        u, ld = self.plan(params).forward(inputs, conditions)
        return self._out(u, inputs), self._out(ld, inputs)

    def inverse(self, params, inputs, conditions=None, **kwargs):
        x, ld = self.plan(params).inverse(inputs, conditions)
        return self._out(x, inputs), self._out(ld, inputs)


def _fused(info: Bijector_Info, input_dim: int, key=None):
    f = _FusedFunctions(info, input_dim, key)
    return ForwardFunction(f.forward), InverseFunction(f.inverse)


def _dense_init(rng: np.random.Generator, in_dim: int, hidden_layers: int, hidden_dim: int, out_dim: int):
    """stax.serial(*(Dense, LeakyRelu) * hidden_layers, Dense) parameter layout
    (pzflow/utils.py:45-48): glorot-normal W[in,out], N(0, 1e-2) b.  Bit parity with
    jax.random initialisation is out of scope."""
    dims = [in_dim] + [hidden_dim] * hidden_layers + [out_dim]
    params = []
    for i in range(len(dims) - 1):
        fi, fo = dims[i], dims[i + 1]
        std = np.sqrt(2.0 / (fi + fo)) if fi + fo > 0 else 0.0
        W = rng.normal(0.0, std, size=(fi, fo)).astype(np.float32)
        b = rng.normal(0.0, 1e-2, size=(fo,)).astype(np.float32)
        params.append((W, b))
        if i < len(dims) - 2:
            params.append(())
    return params


# --------------------------------------------------------------------------- #
@Bijector
def Chain(*inputs: Sequence[Tuple[InitFunction, Bijector_Info]]) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that chains multiple InitFunctions into a single InitFunction
    (pzflow/bijectors.py:121-174).  The chain evaluates as ONE fused kernel."""
    init_funs = tuple(i[0] for i in inputs)
    bijector_info = ("Chain", tuple(i[1] for i in inputs))

    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        key = _jax_key(rng)
        rng = _generator(rng)
        all_params = []
        for init_f in init_funs:
            rng, layer_rng = _split(rng)
            param, _, _ = init_f(layer_rng, input_dim)
            all_params.append(param)
        forward_fun, inverse_fun = _fused(bijector_info, input_dim, key)
        return all_params, forward_fun, inverse_fun

    return init_fun, bijector_info


@Bijector
def NeuralSplineCoupling(K: int = 16, B: float = 5, hidden_layers: int = 2, hidden_dim: int = 128,
                         transformed_dim: int = None, n_conditions: int = 0,
                         periodic: bool = False) -> Tuple[InitFunction, Bijector_Info]:
    """A coupling layer bijection with rational quadratic splines
    (pzflow/bijectors.py:375-522)."""
    if not isinstance(periodic, bool):
        raise ValueError("`periodic` must be True or False.")
    bijector_info = ("NeuralSplineCoupling",
                     (K, B, hidden_layers, hidden_dim, transformed_dim, n_conditions, periodic))

    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        if transformed_dim is None:
            upper_dim = input_dim // 2
            lower_dim = input_dim - upper_dim
        else:
            upper_dim = input_dim - transformed_dim
            lower_dim = transformed_dim
        network_params = _dense_init(_generator(rng), upper_dim + n_conditions, hidden_layers, hidden_dim,
                                     (3 * K - 1 + int(periodic)) * lower_dim)
        forward_fun, inverse_fun = _fused(bijector_info, input_dim)
        return network_params, forward_fun, inverse_fun

    return init_fun, bijector_info


@Bijector
def Roll(shift: int = 1) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that rolls inputs along their last column (pzflow/bijectors.py:559-598)."""
    if not isinstance(shift, int):
        raise ValueError("shift must be an integer.")
    bijector_info = ("Roll", (shift,))

    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        forward_fun, inverse_fun = _fused(bijector_info, input_dim)
        return (), forward_fun, inverse_fun

    return init_fun, bijector_info


@Bijector
def RollingSplineCoupling(nlayers: int, shift: int = 1, K: int = 16, B: float = 5, hidden_layers: int = 2,
                          hidden_dim: int = 128, transformed_dim: int = None, n_conditions: int = 0,
                          periodic: bool = False) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that alternates NeuralSplineCouplings and Roll bijections
    (pzflow/bijectors.py:601-665)."""
    return Chain(*(NeuralSplineCoupling(K=K, B=B, hidden_layers=hidden_layers, hidden_dim=hidden_dim,
                                        transformed_dim=transformed_dim, n_conditions=n_conditions,
                                        periodic=periodic),
                   Roll(shift)) * nlayers)


@Bijector
def ShiftBounds(min: float, max: float, B: float = 5) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector shifts the bounds of inputs so the lie in the range (-B, B)
    (pzflow/bijectors.py:717-777)."""
    min = np.atleast_1d(np.asarray(min, dtype=np.float32))
    max = np.atleast_1d(np.asarray(max, dtype=np.float32))
    if len(min) != len(max):
        raise ValueError("Lengths of min and max do not match. "
                         + "Please provide either a single min and max, "
                         + "or a min and max for each dimension.")
    if (min > max).any():
        raise ValueError("All mins must be less than maxes.")
    bijector_info = ("ShiftBounds", (min, max, B))

    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        forward_fun, inverse_fun = _fused(bijector_info, input_dim)
        return (), forward_fun, inverse_fun

    return init_fun, bijector_info


@Bijector
def StandardScaler(means, stds) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that applies standard scaling to each input (pzflow/bijectors.py:817-861)."""
    bijector_info = ("StandardScaler", (means, stds))

    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        forward_fun, inverse_fun = _fused(bijector_info, input_dim)
        return (), forward_fun, inverse_fun

    return init_fun, bijector_info


def _parameter_free(bijector_info: Bijector_Info) -> InitFunction:
    @InitFunction
    def init_fun(rng, input_dim, **kwargs):
        forward_fun, inverse_fun = _fused(bijector_info, input_dim, _jax_key(rng))
        return (), forward_fun, inverse_fun

    return init_fun


@Bijector
def ColorTransform(ref_idx: int, mag_idx) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that calculates photometric colors from magnitudes (pzflow/bijectors.py:177-311):
    [.., mags ..] -> [other columns, reference magnitude, adjacent colors]."""
    if ref_idx <= 0:
        raise ValueError("ref_idx must be a positive integer.")
    if not isinstance(ref_idx, int):
        raise ValueError("ref_idx must be an integer.")
    if ref_idx not in mag_idx:
        raise ValueError("ref_idx must be in mag_idx.")
    bijector_info = ("ColorTransform", (ref_idx, mag_idx))
    return _parameter_free(bijector_info), bijector_info


@Bijector
def InvSoftplus(column_idx, sharpness=1) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that applies inverse softplus to the specified column(s)
    (pzflow/bijectors.py:314-372)."""
    idx = np.atleast_1d(column_idx)
    k = np.atleast_1d(sharpness)
    if len(idx) != len(k) and len(k) != 1:
        raise ValueError("Please provide either a single sharpness or one for each column index.")
    bijector_info = ("InvSoftplus", (column_idx, sharpness))
    return _parameter_free(bijector_info), bijector_info


@Bijector
def Reverse() -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that reverses the order of inputs (pzflow/bijectors.py:525-556)."""
    bijector_info = ("Reverse", ())
    return _parameter_free(bijector_info), bijector_info


@Bijector
def Scale(scale: float) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that multiplies inputs by a scalar (pzflow/bijectors.py:668-714)."""
    if isinstance(scale, np.ndarray):
        if scale.dtype != np.float32:
            raise ValueError("scale must be a float or array of floats.")
    elif not isinstance(scale, float):
        raise ValueError("scale must be a float or array of floats.")
    bijector_info = ("Scale", (scale,))
    return _parameter_free(bijector_info), bijector_info


@Bijector
def Shuffle() -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that randomly permutes inputs (pzflow/bijectors.py:780-814).  The permutation is
    ``jax.random.permutation(rng, arange(input_dim))`` of the key the InitFunction receives -- reproduced on the host
    (pzflow_b200.prng) -- and, like Roll, is folded into the plan's column maps."""
    bijector_info = ("Shuffle", ())
    return _parameter_free(bijector_info), bijector_info


@Bijector
def UniformDequantizer(column_idx) -> Tuple[InitFunction, Bijector_Info]:
    """Bijector that dequantizes discrete variables with uniform noise (pzflow/bijectors.py:865-911) -- as the
    reference executes it: forward is the identity (the noise is computed and dropped, line 897), inverse floors the
    column(s) ``column_idx``; log-det 0 both ways."""
    bijector_info = ("UniformDequantizer", (column_idx,))
    return _parameter_free(bijector_info), bijector_info

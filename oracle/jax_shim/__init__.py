"""A NumPy-backed stand-in for the slice of the ``jax`` API pzflow uses (TEST INFRASTRUCTURE).

jax / jaxlib / optax are not installed in the build container and cannot be
(no network), so the reference cannot be imported as is.  ``install()``
registers minimal ``jax``, ``jax.numpy``, ``jax.random``, ``jax.nn``,
``jax.scipy.*``, ``jax.example_libraries.stax`` and ``optax`` modules in
``sys.modules`` and puts /root/reference on ``sys.path``; after that
``import pzflow`` executes the reference's OWN, unmodified source
(pzflow/flow.py, bijectors.py, utils.py, distributions.py, flowEnsemble.py).

What this pins when its outputs are used as golden vectors: every line of the
reference's control flow -- column selection, condition scaling, the Chain
order and its reversal, roll direction, the upper/lower split, reshape/split of
the conditioner output, knot padding, the ``<=`` bin rule, the gathers, the
out-of-bounds masking, the log-det sums, the latent densities, the posterior
grid expansion / normalisation and the ensemble reductions.

What it does NOT pin: XLA's arithmetic.  Primitives here are NumPy float32
(``exp``/``log`` polynomials, matmul blocking and reduction order differ from
XLA's at the ulp level; random streams are NumPy's, not threefry).  Semantics
that differ between NumPy and JAX are restated explicitly below and marked.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
_F32 = np.float32


class ShimArray(np.ndarray):
    """ndarray with jax's functional-update ``.at[idx].set(v)``."""

    def __array_finalize__(self, obj):
        pass

    def astype(self, dtype, *args, **kwargs):
        # jax with x64 disabled: a request for float64 (e.g. ``.astype(float)``, pzflow/bijectors.py:896) yields float32
        if dtype is float or np.dtype(dtype) == np.float64:
            dtype = _F32
        return np.ndarray.astype(self, dtype, *args, **kwargs)

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, values, **kwargs):
        out = np.array(self._arr, copy=True)
        out[self._idx] = np.asarray(values)
        return _out(out)


def _in(x):
    """jax treats Python floats as weakly-typed float32 when x64 is disabled."""
    if isinstance(x, float):
        return _F32(x)
    return x


def _out(x):
    """Canonicalise results the way jax (x64 disabled) does: float64 -> float32."""
    if isinstance(x, (tuple, list)):
        return type(x)(_out(v) for v in x)
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            x = x.astype(_F32)
        return x.view(ShimArray)
    if isinstance(x, np.floating):
        return np.asarray(x, dtype=_F32).view(ShimArray)
    return x


def _wrap(fn):
    def wrapped(*args, **kwargs):
        args = tuple(_in(a) for a in args)
        kwargs = {k: _in(v) for k, v in kwargs.items()}
        with np.errstate(all="ignore"):
            return _out(fn(*args, **kwargs))
    wrapped.__name__ = getattr(fn, "__name__", "wrapped")
    return wrapped


# --------------------------------------------------------------------------- #
# jax.numpy                                                                   #
# --------------------------------------------------------------------------- #
def _make_jnp() -> types.ModuleType:
    jnp = types.ModuleType("jax.numpy")

    def _creation(fn):
        def f(*args, dtype=None, **kwargs):
            out = fn(*args, **kwargs) if dtype is None else fn(*args, dtype=dtype, **kwargs)
            if dtype is None and isinstance(out, np.ndarray) and out.dtype == np.float64:
                out = out.astype(_F32)
            return _out(out)
        return f

    jnp.array = _creation(np.array)
    jnp.asarray = _creation(np.asarray)
    jnp.zeros = _creation(np.zeros)
    jnp.ones = _creation(np.ones)
    jnp.linspace = _creation(np.linspace)
    jnp.arange = _creation(np.arange)
    jnp.identity = _creation(np.identity)

    def take_along_axis(arr, indices, axis, mode=None):
        # JAX semantics: negative indices wrap; out-of-bound gathers return NaN ("fill")
        arr, indices = np.asarray(arr), np.asarray(indices)
        n = arr.shape[axis]
        idx = np.where(indices < 0, indices + n, indices)
        bad = (idx < 0) | (idx >= n)
        out = np.take_along_axis(arr, np.clip(idx, 0, n - 1), axis)
        if np.issubdtype(out.dtype, np.floating):
            out = np.where(bad, np.asarray(np.nan, dtype=out.dtype), out)
        return _out(out)

    def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
        # JAX applies the substitutions sequentially to the running result
        x = np.asarray(x)
        info = np.finfo(x.dtype)
        posinf = info.max if posinf is None else posinf
        neginf = info.min if neginf is None else neginf
        with np.errstate(all="ignore"):
            out = np.where(np.isnan(x), np.asarray(nan, dtype=x.dtype), x)
            out = np.where(np.isposinf(out), np.asarray(posinf, dtype=x.dtype), out)
            out = np.where(np.isneginf(out), np.asarray(neginf, dtype=x.dtype), out)
        return _out(out.astype(x.dtype))

    jnp.take_along_axis = take_along_axis
    jnp.nan_to_num = nan_to_num
    jnp.ndarray = np.ndarray
    jnp.inf, jnp.nan, jnp.pi = np.inf, np.nan, np.pi
    jnp.float32, jnp.int32 = np.float32, np.int32
    jnp.linalg = np.linalg

    def __getattr__(name):
        fn = getattr(np, name)
        return _wrap(fn) if callable(fn) else fn

    jnp.__getattr__ = __getattr__
    return jnp


# --------------------------------------------------------------------------- #
# jax.random (NumPy streams -- NOT threefry)                                  #
# --------------------------------------------------------------------------- #
def _make_random(prng: str = "numpy") -> types.ModuleType:
    rnd = types.ModuleType("jax.random")
    if prng in ("threefry", "threefry_partitionable"):
        # jax's own streams (oracle/jax_prng.py: Threefry-2x32/20) for keys, split, uniform and permutation, in the
        # ORIGINAL layout ("threefry": jax < 0.5's default, pinned by the samples real JAX printed in
        # docs/tutorials/marginalization.ipynb) or the PARTITIONABLE one ("threefry_partitionable": the default of the
        # jax >= 0.5 the reference requires, pyproject.toml:13; pinned by the values jax's documentation prints);
        # normal / beta keep NumPy generators seeded by the key words here (random.normal IS restated since --
        # oracle/jax_prng.normal -- but the goldens cut in these modes only use it for initial weights, which they
        # store), so Uniform latents, key schedules and Shuffle are what is stream-exact in these modes.
        from .. import jax_prng as J
        part = prng == "threefry_partitionable"

        def _g(key):
            return np.random.default_rng([int(v) for v in np.asarray(key).ravel()])

        rnd.PRNGKey = J.PRNGKey
        rnd.split = lambda key, num=2: J.split(np.asarray(key), num, part)
        rnd.uniform = lambda key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0: _out(
            J.uniform(np.asarray(key), tuple(shape), minval, maxval, part))
        rnd.permutation = lambda key, x: _out(np.asarray(x)[J.permutation(np.asarray(key), len(np.asarray(x)), part)])
        rnd.normal = lambda key, shape=(), dtype=np.float32: _out(_g(key).standard_normal(size=shape).astype(_F32))
        rnd.beta = lambda key, a, b, shape=None, dtype=np.float32: _out(_g(key).beta(a, b, size=shape).astype(_F32))
        rnd.multivariate_normal = lambda key, mean, cov, shape=None: _out(
            _g(key).multivariate_normal(np.asarray(mean), np.asarray(cov), size=shape).astype(_F32))
        return rnd

    def PRNGKey(seed):
        seed = int(seed)
        return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)

    def _gen(key):
        return np.random.default_rng([int(v) for v in np.asarray(key).ravel()])

    def split(key, num=2):
        g = _gen(key)
        return [g.integers(0, 2 ** 32, size=2, dtype=np.uint32) for _ in range(num)]

    def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
        return _out(_gen(key).uniform(minval, maxval, size=shape).astype(_F32))

    def normal(key, shape=(), dtype=np.float32):
        return _out(_gen(key).standard_normal(size=shape).astype(_F32))

    def beta(key, a, b, shape=None, dtype=np.float32):
        return _out(_gen(key).beta(a, b, size=shape).astype(_F32))

    def permutation(key, x):
        return _out(_gen(key).permutation(np.asarray(x)))

    def multivariate_normal(key, mean, cov, shape=None):
        return _out(_gen(key).multivariate_normal(np.asarray(mean), np.asarray(cov), size=shape).astype(_F32))

    rnd.PRNGKey, rnd.split, rnd.uniform, rnd.normal = PRNGKey, split, uniform, normal
    rnd.beta, rnd.permutation, rnd.multivariate_normal = beta, permutation, multivariate_normal
    return rnd


# --------------------------------------------------------------------------- #
# jax.nn, jax.scipy, stax                                                     #
# --------------------------------------------------------------------------- #
def _make_nn() -> types.ModuleType:
    nn = types.ModuleType("jax.nn")

    def softmax(x, axis=-1):
        x = np.asarray(x)
        with np.errstate(all="ignore"):
            e = np.exp(x - x.max(axis=axis, keepdims=True))
            return _out(e / e.sum(axis=axis, keepdims=True))

    def softplus(x):
        # jax.nn.softplus = jnp.logaddexp(x, 0) = amax + log1p(exp(-|delta|)) (NaN delta -> x1 + x2)
        x = np.asarray(x)
        zero = x.dtype.type(0)
        with np.errstate(all="ignore"):
            amax, delta = np.maximum(x, zero), x - zero
            return _out(np.where(np.isnan(delta), x + zero, amax + np.log1p(np.exp(-np.abs(delta)))))

    def leaky_relu(x, negative_slope=0.01):
        x = np.asarray(x)
        return _out(np.where(x >= 0, x, x.dtype.type(negative_slope) * x))

    nn.softmax, nn.softplus, nn.leaky_relu = softmax, softplus, leaky_relu
    return nn


def _make_scipy():
    import scipy.special as sps

    sp = types.ModuleType("jax.scipy")
    special = types.ModuleType("jax.scipy.special")
    stats = types.ModuleType("jax.scipy.stats")
    integrate = types.ModuleType("jax.scipy.integrate")

    special.gammaln = lambda x: _out(sps.gammaln(np.asarray(x, dtype=_F32)).astype(_F32))

    class _Beta:
        @staticmethod
        def logpdf(x, a, b, loc=0, scale=1):
            # jax.scipy.stats.beta.logpdf: -betaln(a,b) + xlogy(a-1, y) + xlog1py(b-1, -y) - log(scale)
            x = np.asarray(x, dtype=_F32)
            a, b, loc, scale = _F32(a), _F32(b), _F32(loc), _F32(scale)
            with np.errstate(all="ignore"):
                betaln = (sps.gammaln(a).astype(_F32) + sps.gammaln(b).astype(_F32)) - sps.gammaln(a + b).astype(_F32)
                y = (x - loc) / scale
                lin = sps.xlogy(a - 1, y).astype(_F32) + sps.xlog1py(b - 1, -y).astype(_F32)
                lp = (-betaln + lin) - np.log(scale)
                lp = np.where((x > loc + scale) | (x < loc), _F32(-np.inf), lp)
            return _out(lp.astype(_F32))

    class _MVN:
        @staticmethod
        def logpdf(x, mean, cov):
            # jax.scipy.stats.multivariate_normal.logpdf for cov = identity (the only use: distributions.py:271-275)
            x, mean, cov = np.asarray(x, dtype=_F32), np.asarray(mean, dtype=_F32), np.asarray(cov, dtype=_F32)
            assert np.array_equal(cov, np.eye(cov.shape[0], dtype=_F32)), "jax_shim: identity covariance only"
            n = _F32(mean.shape[-1])
            y = x - mean
            return _out((_F32(-0.5) * (n * np.log(_F32(2 * np.pi)) + (y * y).sum(axis=-1, dtype=_F32))).astype(_F32))

    stats.beta, stats.multivariate_normal = _Beta, _MVN

    def trapezoid(y, x=None, dx=1.0, axis=-1):
        y = np.asarray(y)
        assert axis == -1
        d = np.diff(np.asarray(x)) if x is not None else y.dtype.type(dx)
        with np.errstate(all="ignore"):
            return _out(y.dtype.type(0.5) * (d * (y[..., 1:] + y[..., :-1])).sum(-1))

    integrate.trapezoid = trapezoid
    sp.special, sp.stats, sp.integrate = special, stats, integrate
    return sp, special, stats, integrate


def _make_stax(rnd, nn):
    stax = types.ModuleType("jax.example_libraries.stax")

    def Dense(out_dim):
        def init_fun(rng, input_shape):
            g = np.random.default_rng([int(v) for v in np.asarray(rng).ravel()])
            fan_in = input_shape[-1]
            std = np.sqrt(2.0 / (fan_in + out_dim)) if fan_in + out_dim else 0.0
            W = g.normal(0.0, std, size=(fan_in, out_dim)).astype(_F32)
            b = g.normal(0.0, 1e-2, size=(out_dim,)).astype(_F32)
            return input_shape[:-1] + (out_dim,), (_out(W), _out(b))

        def apply_fun(params, inputs, **kwargs):
            W, b = params
            return _out(np.asarray(inputs) @ np.asarray(W) + np.asarray(b))

        return init_fun, apply_fun

    def _elementwise(fn):
        return (lambda rng, input_shape: (input_shape, ())), (lambda params, inputs, **kw: fn(inputs))

    def serial(*layers):
        inits, applies = zip(*layers)

        def init_fun(rng, input_shape):
            params = []
            for init in inits:
                rng, layer_rng = rnd.split(rng)
                input_shape, p = init(layer_rng, input_shape)
                params.append(p)
            return input_shape, params

        def apply_fun(params, inputs, **kwargs):
            for fn, p in zip(applies, params):
                inputs = fn(p, inputs, **kwargs)
            return inputs

        return init_fun, apply_fun

    stax.Dense, stax.serial = Dense, serial
    stax.LeakyRelu = _elementwise(nn.leaky_relu)
    return stax


def install(reference_root: str = REFERENCE_ROOT, prng: str = "numpy"):
    """Register the stand-in modules and make the reference importable. Returns ``pzflow``.
    prng: "numpy" (the streams every round-1/2 golden was cut with), or "threefry" / "threefry_partitionable" (jax's own
    key schedule and uniform stream in the original / partitionable layout, see _make_random)."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__pzb_shim__", False):
        raise RuntimeError("a real jax is already imported; use it instead of the shim")
    jax = types.ModuleType("jax")
    jax.__pzb_shim__ = True
    jnp, rnd, nn = _make_jnp(), _make_random(prng), _make_nn()
    sp, special, stats, integrate = _make_scipy()
    ex = types.ModuleType("jax.example_libraries")
    stax = _make_stax(rnd, nn)
    ex.stax = stax
    jax.numpy, jax.random, jax.nn, jax.scipy, jax.example_libraries = jnp, rnd, nn, sp, ex
    jax.jit = lambda f, *a, **k: f

    def grad(*a, **k):
        raise NotImplementedError("jax_shim: autodiff is outside the hot path")

    jax.grad = grad
    optax = types.ModuleType("optax")
    optax.__getattr__ = lambda name: (lambda *a, **k: (_ for _ in ()).throw(
        NotImplementedError("jax_shim: optax is outside the hot path")))
    mods = {"jax": jax, "jax.numpy": jnp, "jax.random": rnd, "jax.nn": nn, "jax.scipy": sp,
            "jax.scipy.special": special, "jax.scipy.stats": stats, "jax.scipy.integrate": integrate,
            "jax.example_libraries": ex, "jax.example_libraries.stax": stax, "optax": optax}
    sys.modules.update(mods)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return importlib.import_module("pzflow")

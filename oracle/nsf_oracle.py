"""CPU oracle for the pzflow neural-spline-flow hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of the reference algorithm, op for op and in
the reference's own evaluation order, in float32 (JAX's default with x64
disabled).  It exists to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``pzflow_b200/`` does, and the
product path fails loudly when the CUDA library is missing.

Pinning status
--------------
The reference's arithmetic lives in jax/jaxlib 0.5.3 (poetry.lock:1498-1533),
which is not vendored under /root/reference and is not installable here, and
the reference's tests hold no known-answer vectors for this path (SURVEY.md
§8c).  The oracle is therefore pinned three ways, none of which is XLA itself:

1. against the reference's OWN Python code executed unmodified over a
   NumPy-backed stand-in for the ``jax`` API (``oracle/jax_shim``; fixtures
   produced by ``oracle/make_golden.py`` live in ``tests/golden/``) -- this
   pins control flow, slicing, roll direction, bin rule, masking, reductions;
2. against the reference test-suite's properties (bijectivity and log-det
   antisymmetry at atol 1e-6 on the fixed 3x7 input, tests/test_bijectors.py:
   7-13,71-86; batched == unbatched posterior, tests/test_flow.py:275-287;
   ensemble rules, tests/test_ensemble.py:20-56);
3. against the one JAX-produced value anchor in the reference tree: the two
   samples printed at docs/tutorials/marginalization.ipynb:133-135 must map
   into the Uniform(7,5) support and round-trip.

What is NOT pinned: XLA's transcendental polynomials and reduction order
("parity unpinned" at the ulp level; DESIGN.md says so too).

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

F32_MIN = np.float32(np.finfo(np.float32).min)


# --------------------------------------------------------------------------- #
# library semantics restated from the JAX docs (jax.nn / jnp)                 #
# --------------------------------------------------------------------------- #
def _softmax(x: np.ndarray) -> np.ndarray:
    """jax.nn.softmax over the last axis: exp(x - max) / sum(exp(x - max))."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def _softplus(x: np.ndarray) -> np.ndarray:
    """jax.nn.softplus = logaddexp(x, 0) = max(x,0) + log1p(exp(-|x|))."""
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _leaky_relu(x: np.ndarray) -> np.ndarray:
    """stax.LeakyRelu = jax.nn.leaky_relu(x, negative_slope=0.01)."""
    return np.where(x >= 0, x, x.dtype.type(0.01) * x)


def _nan_to_num(x: np.ndarray, nan: float) -> np.ndarray:
    """jnp.nan_to_num: the three substitutions are applied one after another to
    the running result (NaN -> nan, then +inf -> finfo.max, then -inf ->
    finfo.min), so with nan=-inf a NaN ends at finfo.min as well."""
    info = np.finfo(x.dtype)
    out = np.where(np.isnan(x), x.dtype.type(nan), x)
    out = np.where(np.isposinf(out), info.max, out)
    out = np.where(np.isneginf(out), info.min, out)
    return out.astype(x.dtype)


def _take_fill(a: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """jnp.take_along_axis(a, idx[..., None], -1)[..., 0] with JAX's default
    out-of-bounds mode ("fill": NaN for floats); negative indices wrap once."""
    n = a.shape[-1]
    idx = np.where(idx < 0, idx + n, idx)
    bad = (idx < 0) | (idx >= n)
    safe = np.clip(idx, 0, n - 1)
    out = np.take_along_axis(a, safe[..., None], -1)[..., 0]
    return np.where(bad, a.dtype.type(np.nan), out)


# --------------------------------------------------------------------------- #
# pzflow/utils.py                                                             #
# --------------------------------------------------------------------------- #
def dense_relu_network(params: Sequence, x: np.ndarray) -> np.ndarray:
    """utils.py:22-49 -- stax.serial(*(Dense, LeakyRelu)*hidden_layers, Dense).

    ``params`` is the stax list ``[(W[in,out], b[out]), (), ..., (W, b)]``;
    Dense is ``x @ W + b``.
    """
    h = x
    for p in params:
        if len(p) == 0:
            h = _leaky_relu(h)
        else:
            W, b = p
            h = h @ np.asarray(W, dtype=x.dtype) + np.asarray(b, dtype=x.dtype)
    return h


class BinTrace(list):
    """A ``bins`` list that also collects, per coupling stage, the searched knots [N, L, K+1] and the (masked) value
    they were compared with [N, L] (utils.py:145-152) -- what a test needs to tell a knot tie from a wrong bin."""

    def __init__(self):
        super().__init__()
        self.knots, self.values = [], []


def rational_quadratic_spline(inputs, W, H, D, B, periodic=False, inverse=False,
                              return_idx=False, trace=None):
    """utils.py:64-227, line for line.

    inputs [N, L]; W, H [N, L, K]; D [N, L, K-1] (K if periodic).
    Returns (outputs [N, L], log_det [N]) (+ idx [N, L] int32 if asked).
    """
    dt = inputs.dtype.type
    B = dt(B)
    # knot positions, utils.py:113-125
    xk = np.pad(-B + np.cumsum(W, axis=-1, dtype=inputs.dtype),
                [(0, 0)] * (W.ndim - 1) + [(1, 0)], mode="constant", constant_values=-B)
    yk = np.pad(-B + np.cumsum(H, axis=-1, dtype=inputs.dtype),
                [(0, 0)] * (H.ndim - 1) + [(1, 0)], mode="constant", constant_values=-B)
    # knot derivatives, utils.py:127-135
    if periodic:
        dk = np.pad(D, [(0, 0)] * (D.ndim - 1) + [(1, 0)], mode="wrap")
    else:
        dk = np.pad(D, [(0, 0)] * (D.ndim - 1) + [(1, 1)], mode="constant", constant_values=1)
    sk = H / W  # utils.py:137

    # utils.py:145-146
    out_of_bounds = (inputs <= -B) | (inputs >= B)
    masked = np.where(out_of_bounds, np.abs(inputs) - B, inputs)

    # utils.py:149-152 -- bin = (#knots <= x) - 1
    knots = yk if inverse else xk
    idx = np.sum(knots <= masked[..., None], axis=-1).astype(np.int32) - 1
    if trace is not None:
        trace.knots.append(knots)
        trace.values.append(masked)

    # utils.py:155-161
    in_xk = _take_fill(xk, idx)
    in_yk = _take_fill(yk, idx)
    in_dk = _take_fill(dk, idx)
    in_dkp1 = _take_fill(dk, idx + 1)
    in_wk = _take_fill(W, idx)
    in_hk = _take_fill(H, idx)
    in_sk = _take_fill(sk, idx)

    with np.errstate(all="ignore"):
        if inverse:
            # utils.py:163-175
            a = in_hk * (in_sk - in_dk) + (masked - in_yk) * (in_dkp1 + in_dk - 2 * in_sk)
            b = in_hk * in_dk - (masked - in_yk) * (in_dkp1 + in_dk - 2 * in_sk)
            c = -in_sk * (masked - in_yk)
            relx = 2 * c / (-b - np.sqrt(b ** 2 - 4 * a * c))
            outputs = relx * in_wk + in_xk
        else:
            # utils.py:198-206
            relx = (masked - in_xk) / in_wk
            num = in_hk * (in_sk * relx ** 2 + in_dk * relx * (1 - relx))
            den = in_sk + (in_dkp1 + in_dk - 2 * in_sk) * relx * (1 - relx)
            outputs = in_yk + num / den
        if not periodic:  # utils.py:177-178, 208-209
            outputs = np.where(out_of_bounds, inputs, outputs)

        # utils.py:182-196, 213-227
        dnum = in_dkp1 * relx ** 2 + 2 * in_sk * relx * (1 - relx) + in_dk * (1 - relx) ** 2
        dden = in_sk + (in_dkp1 + in_dk - 2 * in_sk) * relx * (1 - relx)
        log_det = 2 * np.log(in_sk) + np.log(dnum) - 2 * np.log(dden)
        if not periodic:
            log_det = np.where(out_of_bounds, dt(0), log_det)
        log_det = log_det.sum(axis=1, dtype=inputs.dtype)

    outputs = outputs.astype(inputs.dtype)
    log_det = (-log_det if inverse else log_det).astype(inputs.dtype)
    if return_idx:
        return outputs, log_det, idx
    return outputs, log_det


# --------------------------------------------------------------------------- #
# pzflow/bijectors.py                                                         #
# --------------------------------------------------------------------------- #
def _nsc_dims(input_dim, transformed_dim):
    """bijectors.py:463-470."""
    if transformed_dim is None:
        upper = input_dim // 2
        lower = input_dim - upper
    else:
        upper = input_dim - transformed_dim
        lower = transformed_dim
    return upper, lower


def nsc_spline_params(params, upper, conditions, K, B, lower_dim, n_conditions, periodic):
    """bijectors.py:480-492."""
    upper_dim = upper.shape[1]
    key = np.hstack((upper, conditions))[:, : upper_dim + n_conditions]
    out = dense_relu_network(params, key)
    out = out.reshape(-1, lower_dim, 3 * K - 1 + int(periodic))
    W, H, D = np.split(out, [K, 2 * K], axis=2)
    dt = upper.dtype.type
    W = dt(2 * B) * _softmax(W)
    H = dt(2 * B) * _softmax(H)
    D = _softplus(D)
    return W, H, D


def neural_spline_coupling(params, inputs, conditions, args, inverse, bins=None):
    """bijectors.py:494-520.  ``args`` is the Bijector_Info argument tuple
    (K, B, hidden_layers, hidden_dim, transformed_dim, n_conditions, periodic)."""
    K, B, _hl, _hd, transformed_dim, n_conditions, periodic = args
    upper_dim, lower_dim = _nsc_dims(inputs.shape[1], transformed_dim)
    upper, lower = inputs[:, :upper_dim], inputs[:, upper_dim:]
    W, H, D = nsc_spline_params(params, upper, conditions, K, B, lower_dim, n_conditions, periodic)
    lower, log_det, idx = rational_quadratic_spline(lower, W, H, D, B, periodic, inverse, return_idx=True,
                                                    trace=bins if isinstance(bins, BinTrace) else None)
    if bins is not None:
        bins.append(idx)
    return np.hstack((upper, lower)), log_det


def roll(inputs, shift, inverse):
    """bijectors.py:585-593 -- jnp.roll(x, +-shift, axis=-1), log_det 0."""
    out = np.roll(inputs, shift=-shift if inverse else shift, axis=-1)
    return out, np.zeros(inputs.shape[0], dtype=inputs.dtype)


def shift_bounds(inputs, mn, mx, B, inverse):
    """bijectors.py:741-773."""
    dt = inputs.dtype
    mn = np.atleast_1d(np.asarray(mn, dtype=dt))
    mx = np.atleast_1d(np.asarray(mx, dtype=dt))
    mean = (mx + mn) / 2
    half_range = (mx - mn) / 2
    Bf = dt.type(B)
    ones = np.ones(inputs.shape[0], dtype=dt)
    if inverse:
        out = inputs * half_range / Bf + mean
        ld = np.log(np.prod(half_range / Bf, dtype=dt)) * ones
    else:
        out = Bf * (inputs - mean) / half_range
        ld = np.log(np.prod(Bf / half_range, dtype=dt)) * ones
    return out.astype(dt), ld.astype(dt)


def standard_scaler(inputs, means, stds, inverse):
    """bijectors.py:848-857."""
    dt = inputs.dtype
    means = np.asarray(means, dtype=dt)
    stds = np.asarray(stds, dtype=dt)
    ones = np.ones(inputs.shape[0], dtype=dt)
    if inverse:
        out = inputs * stds + means
        ld = np.log(np.prod(stds, dtype=dt)) * ones
    else:
        out = (inputs - means) / stds
        ld = np.log(dt.type(1) / np.prod(stds, dtype=dt)) * ones
    return out.astype(dt), ld.astype(dt)


def reverse(inputs):
    """bijectors.py:541-553 -- inputs[:, ::-1] both ways, log_det 0."""
    return inputs[:, ::-1].copy(), np.zeros(inputs.shape[0], dtype=inputs.dtype)


def scale_bijector(inputs, scale, inverse):
    """bijectors.py:689-711 -- scale * x, log_det = log(scale ** D); inverse 1 / scale * x."""
    dt = inputs.dtype.type
    ld = np.log(dt(scale) ** dt(inputs.shape[-1])) * np.ones(inputs.shape[0], dtype=inputs.dtype)
    if inverse:
        return (dt(1) / dt(scale)) * inputs, -ld
    return dt(scale) * inputs, ld


def inv_softplus(inputs, column_idx, sharpness, inverse):
    """bijectors.py:350-369."""
    dt = inputs.dtype.type
    idx = np.atleast_1d(column_idx)
    k = np.atleast_1d(np.asarray(sharpness, dtype=inputs.dtype))
    outputs = inputs.copy()
    with np.errstate(all="ignore"):
        if inverse:
            outputs[:, idx] = np.log(dt(1) + np.exp(k * inputs[:, idx])) / k
            log_det = -np.log(dt(1) + np.exp(-k * inputs[:, idx])).sum(axis=1, dtype=inputs.dtype)
        else:
            outputs[:, idx] = np.log(dt(-1) + np.exp(k * inputs[:, idx])) / k
            log_det = np.log(dt(1) + np.exp(-k * outputs[:, idx])).sum(axis=1, dtype=inputs.dtype)
    return outputs, log_det


def color_transform(inputs, ref_idx, mag_idx, inverse):
    """bijectors.py:236-308."""
    input_dim = inputs.shape[1]
    mag_idx = np.asarray(mag_idx)
    front_idx = np.setdiff1d(np.arange(input_dim), mag_idx)
    mag0_idx = len(front_idx)
    new_idx = np.concatenate((front_idx, mag_idx))
    new_ref = np.where(new_idx == ref_idx)[0][0]
    zeros = np.zeros(inputs.shape[0], dtype=inputs.dtype)
    if not inverse:
        outputs = np.hstack((inputs[:, front_idx], inputs[:, ref_idx, None], -np.diff(inputs[:, mag_idx])))
        return outputs.astype(inputs.dtype), zeros
    outputs = np.hstack((inputs[:, 0:mag0_idx], inputs[:, mag0_idx, None],
                         np.cumsum(inputs[:, mag0_idx + 1:], axis=-1, dtype=inputs.dtype)))
    if ref_idx != mag_idx[0]:
        outputs[:, mag0_idx] = outputs[:, mag0_idx] + outputs[:, new_ref]
    outputs[:, mag0_idx + 1:] = outputs[:, mag0_idx, None] - outputs[:, mag0_idx + 1:]
    outputs = outputs[:, np.argsort(new_idx)]
    return outputs.astype(inputs.dtype), zeros


def shuffle(inputs, perm, inverse):
    """bijectors.py:798-810 -- inputs[:, perm] forward, inputs[:, argsort(perm)] inverse, log_det 0."""
    perm = np.asarray(perm, dtype=np.int64)
    idx = np.argsort(perm) if inverse else perm
    return inputs[:, idx].copy(), np.zeros(inputs.shape[0], dtype=inputs.dtype)


def uniform_dequantizer(inputs, column_idx, inverse):
    """bijectors.py:892-907 AS EXECUTED: forward returns inputs.astype(float) unchanged (line 897 drops the result of
    .at[].set()), inverse floors the listed columns; log_det 0."""
    out = inputs.copy()
    if inverse:
        idx = np.atleast_1d(column_idx)
        out[:, idx] = np.floor(inputs[:, idx])
    return out, np.zeros(inputs.shape[0], dtype=inputs.dtype)


def resolve_shuffles(info, key, input_dim, partitionable=False):
    """Give every ("Shuffle", ()) of ``info`` the permutation the reference draws for it: Chain.init_fun splits its key
    once per stage (bijectors.py:146-149), Shuffle.init_fun calls random.permutation(rng, arange(input_dim)) (795-797).
    key: jax-style uint32 pair (oracle.jax_prng.PRNGKey(seed); PRNGKey(0) for a flow loaded from a file, flow.py:186-187);
    partitionable: jax's stream layout (oracle/jax_prng.py)."""
    from . import jax_prng as J
    name, args = info
    if name == "Shuffle":
        return info if len(args) == 1 else ("Shuffle", (tuple(int(v) for v in J.permutation(key, input_dim, partitionable)),))
    if name != "Chain":
        return info
    out = []
    for sub in args:
        key, layer = J.split(key, 2, partitionable)
        out.append(resolve_shuffles(sub, layer, input_dim, partitionable))
    return ("Chain", tuple(out))


def apply_bijector(info, params, inputs, conditions, inverse=False, bins=None):
    """Interpret a Bijector_Info tuple the way utils.py:11-19 rebuilds it and
    bijectors.py:155-170 (Chain) runs it.  Returns (outputs, log_det)."""
    name, args = info
    if name == "Chain":
        # bijectors.py:155-160; the inverse walks params[::-1], funs[::-1]
        log_dets = np.zeros(inputs.shape[0], dtype=inputs.dtype)
        stages = list(zip(args, params))
        if inverse:
            stages = stages[::-1]
        for sub_info, sub_params in stages:
            inputs, ld = apply_bijector(sub_info, sub_params, inputs, conditions, inverse, bins)
            log_dets = log_dets + ld
        return inputs, log_dets
    if name == "NeuralSplineCoupling":
        return neural_spline_coupling(params, inputs, conditions, args, inverse, bins)
    if name == "RollingSplineCoupling":
        return apply_bijector(expand_rolling(args), params, inputs, conditions, inverse, bins)
    if name == "Roll":
        return roll(inputs, args[0], inverse)
    if name == "ShiftBounds":
        return shift_bounds(inputs, args[0], args[1], args[2], inverse)
    if name == "StandardScaler":
        return standard_scaler(inputs, args[0], args[1], inverse)
    if name == "Reverse":
        return reverse(inputs)
    if name == "Scale":
        return scale_bijector(inputs, args[0], inverse)
    if name == "InvSoftplus":
        return inv_softplus(inputs, args[0], args[1] if len(args) > 1 else 1, inverse)
    if name == "ColorTransform":
        return color_transform(inputs, args[0], args[1], inverse)
    if name == "Shuffle":
        if len(args) != 1:
            raise ValueError("oracle: Shuffle needs its permutation (resolve_shuffles)")
        return shuffle(inputs, args[0], inverse)
    if name == "UniformDequantizer":
        return uniform_dequantizer(inputs, args[0], inverse)
    raise NotImplementedError(f"oracle: bijector {name!r} is outside the hot path (SURVEY.md §2)")


def expand_rolling(args):
    """bijectors.py:651-665 -- RollingSplineCoupling(nlayers, shift, K, B,
    hidden_layers, hidden_dim, transformed_dim, n_conditions, periodic) is
    Chain(*(NSC(...), Roll(shift)) * nlayers)."""
    defaults = (None, 1, 16, 5, 2, 128, None, 0, False)
    full = tuple(args) + defaults[len(args):]
    nlayers, shift, K, B, hl, hd, td, nc, periodic = full
    nsc = ("NeuralSplineCoupling", (K, B, hl, hd, td, nc, periodic))
    return ("Chain", (nsc, ("Roll", (shift,))) * nlayers)


# --------------------------------------------------------------------------- #
# pzflow/distributions.py                                                     #
# --------------------------------------------------------------------------- #
def latent_log_prob(latent_info, u):
    """Uniform: distributions.py:414-441.  CentBeta13: distributions.py:165-194
    (jax.scipy.stats.beta.logpdf(x, 13, 13, loc=-B, scale=2B) per column,
    hstack, sum over axis 1)."""
    name, args = latent_info
    dt = u.dtype.type
    with np.errstate(all="ignore"):
        if name == "Uniform":
            input_dim, B = args
            mask = np.prod((u >= -B) & (u <= B), axis=-1)
            val = dt(-input_dim) * np.log(dt(2 * B))
            return np.where(mask, val, dt(-np.inf)).astype(u.dtype)
        if name == "CentBeta13":
            input_dim, B = args
            a = b = dt(13)
            # betaln(a, b) = gammaln(a) + gammaln(b) - gammaln(a + b), in the array dtype
            betaln = (dt(math.lgamma(13.0)) + dt(math.lgamma(13.0))) - dt(math.lgamma(26.0))
            loc, scale = dt(-B), dt(2 * B)
            cols = []
            for i in range(input_dim):
                x = u[:, i]
                y = (x - loc) / scale
                lin = (a - 1) * np.log(y) + (b - 1) * np.log1p(-y)
                lp = (-betaln + lin) - np.log(scale)
                lp = np.where((x > loc + scale) | (x < loc), dt(-np.inf), lp)
                cols.append(lp.reshape(-1, 1))
            return np.hstack(cols).sum(axis=1, dtype=u.dtype)
        if name == "Normal":
            # distributions.py:255-275 -- multivariate_normal.logpdf(u, zeros, identity)
            (input_dim,) = args
            return (dt(-0.5) * (dt(input_dim) * np.log(dt(2 * np.pi)) + (u * u).sum(axis=1, dtype=u.dtype))).astype(u.dtype)
    raise NotImplementedError(f"oracle: latent {name!r} is outside the hot path")


def tdist_log_prob(log_nu, u):
    """distributions.py:330-358 with cov = identity: maha = sum(u^2), log_det = 0."""
    from scipy import special as sps
    dt = u.dtype.type
    n = u.shape[1]
    nu = np.exp(dt(log_nu))
    maha = (u * u).sum(axis=1, dtype=u.dtype)
    t = dt(0.5) * (nu + dt(n))
    A = dt(sps.gammaln(t))
    Bc = dt(sps.gammaln(dt(0.5) * nu))
    Cc = dt(n / 2.0) * np.log(nu * dt(np.pi))
    E = -t * np.log(dt(1) + (dt(1.0) / nu) * maha)
    return ((A - Bc) - Cc - dt(0)) + E


def joint_log_prob(latent_info, latent_params, u):
    """distributions.py:528-541: split the columns, sum the sub-distributions' log-probs."""
    infos = latent_info[1][1]
    dims = [int(i[1][0]) for i in infos]
    parts = np.split(u, np.cumsum(dims)[:-1], axis=1)
    cols = [any_latent_log_prob(tuple(infos[i]), latent_params[i], parts[i]).reshape(-1, 1) for i in range(len(infos))]
    return np.hstack(cols).sum(axis=1, dtype=u.dtype)


def any_latent_log_prob(latent_info, latent_params, u):
    name = latent_info[0]
    with np.errstate(all="ignore"):
        if name == "CentBeta":
            return centbeta_log_prob(latent_params, u, latent_info[1][1])
        if name == "Tdist":
            return tdist_log_prob(latent_params, u)
        if name == "Joint":
            return joint_log_prob(latent_info, latent_params, u)
        return latent_log_prob(latent_info, u)


def centbeta_log_prob(latent_params, u, B):
    """distributions.py:70-99 -- sum_i beta.logpdf(u_i; exp(p_i0), exp(p_i1), loc=-B, scale=2B), with
    jax.scipy.stats.beta.logpdf = -betaln(a, b) + xlogy(a-1, y) + xlog1py(b-1, -y) - log(scale)."""
    from scipy import special as sps
    dt = u.dtype.type
    loc, scale = dt(-B), dt(2 * B)
    cols = []
    with np.errstate(all="ignore"):
        for i, (pa, pb) in enumerate(latent_params):
            a, b = np.exp(dt(pa)), np.exp(dt(pb))
            betaln = (dt(sps.gammaln(a)) + dt(sps.gammaln(b))) - dt(sps.gammaln(a + b))
            x = u[:, i]
            y = (x - loc) / scale
            lin = sps.xlogy(a - dt(1), y).astype(u.dtype) + sps.xlog1py(b - dt(1), -y).astype(u.dtype)
            lp = (-betaln + lin) - np.log(scale)
            lp = np.where((x > loc + scale) | (x < loc), dt(-np.inf), lp)
            cols.append(lp.reshape(-1, 1))
    return np.hstack(cols).sum(axis=1, dtype=u.dtype)


# --------------------------------------------------------------------------- #
# pzflow/flow.py, pzflow/flowEnsemble.py                                      #
# --------------------------------------------------------------------------- #
def flow_log_prob(bijector_info, latent_info, params, X, conditions=None, bins=None):
    """flow.py:397-407 -- forward, latent log-density + log_det, NaN -> zero
    probability.  ``params`` = (latent_params, bijector_params)."""
    X = np.ascontiguousarray(X)
    if conditions is None:
        conditions = np.zeros((X.shape[0], 1), dtype=X.dtype)  # flow.py:329
    with np.errstate(all="ignore"):
        u, log_det = apply_bijector(bijector_info, params[1], X, conditions, False, bins)
        lat = any_latent_log_prob(latent_info, params[0], u)
        lp = lat + log_det
    return _nan_to_num(lp.astype(X.dtype), nan=-np.inf)


def flow_inverse(bijector_info, params, U, conditions=None, bins=None):
    """flow.py:795 -- x = inverse(params[1], u, conditions)[0]."""
    if conditions is None:
        conditions = np.zeros((U.shape[0], 1), dtype=U.dtype)
    with np.errstate(all="ignore"):
        x, ld = apply_bijector(bijector_info, params[1], U, conditions, True, bins)
    return x, ld


def trapezoid(y, x):
    """jax.scipy.integrate.trapezoid: 0.5 * sum(dx * (y[1:] + y[:-1]))."""
    dx = np.diff(x)
    return (dx * (y[..., 1:] + y[..., :-1])).sum(-1, dtype=y.dtype) * y.dtype.type(0.5)


def flow_posterior(bijector_info, latent_info, params, rest, col_idx, grid,
                   conditions=None, normalize=True, nan_to_zero=True, batch_size=None):
    """flow.py:666-738 (err_samples=None, marg_rules=None branch).

    ``rest`` [Ng, D-1] holds the data columns with ``column`` removed;
    returns pdfs [Ng, G]."""
    dt = rest.dtype
    grid = np.asarray(grid, dtype=dt)
    nrows, G = rest.shape[0], grid.size
    batch_size = nrows if batch_size is None else batch_size
    pdfs = np.zeros((nrows, G), dtype=dt)
    for b0 in range(0, nrows, batch_size):
        batch = rest[b0:b0 + batch_size]
        cond = (np.zeros((batch.shape[0], 1), dtype=dt) if conditions is None
                else conditions[b0:b0 + batch_size])
        # flow.py:697-714 -- galaxy-major, grid-minor expansion
        full = np.hstack((np.repeat(batch[:, :col_idx], G, axis=0),
                          np.tile(grid, len(batch))[:, None],
                          np.repeat(batch[:, col_idx:], G, axis=0)))
        cond = np.repeat(cond, G, axis=0)
        lp = flow_log_prob(bijector_info, latent_info, params, full, cond).reshape(-1, G)
        with np.errstate(all="ignore"):
            pdfs[b0:b0 + batch_size] = np.exp(lp)
    with np.errstate(all="ignore"):
        if normalize:  # flow.py:732-734
            pdfs = pdfs / trapezoid(pdfs, grid).reshape(-1, 1)
        if nan_to_zero:  # flow.py:735-737
            pdfs = _nan_to_num(pdfs, nan=0.0)
    return pdfs


def ensemble_log_prob(flows, X, conditions=None, return_ensemble=False):
    """flowEnsemble.py:227-243.  ``flows`` = [(bijector_info, latent_info,
    params), ...]; naive log(mean(exp(lp)))."""
    ens = np.stack([flow_log_prob(bi, li, p, X, conditions) for bi, li, p in flows], axis=1)
    if return_ensemble:
        return ens
    with np.errstate(all="ignore"):
        return np.log(np.exp(ens).mean(axis=1, dtype=ens.dtype)).astype(ens.dtype)


def ensemble_posterior(flows, rest, col_idx, grid, conditions=None, normalize=True,
                       return_ensemble=False, nan_to_zero=True):
    """flowEnsemble.py:317-351 -- un-normalised per-flow pdfs, mean over flows,
    then trapezoid-normalise."""
    grid = np.asarray(grid, dtype=rest.dtype)
    ens = np.stack([flow_posterior(bi, li, p, rest, col_idx, grid, conditions,
                                   normalize=False, nan_to_zero=nan_to_zero)
                    for bi, li, p in flows], axis=1)  # [Ng, Nf, G]
    with np.errstate(all="ignore"):
        if return_ensemble:
            if normalize:
                flat = ens.reshape(-1, grid.size)
                flat = flat / trapezoid(flat, grid).reshape(-1, 1)
                ens = flat.reshape(rest.shape[0], -1, grid.size)
            return ens
        pdfs = ens.mean(axis=1, dtype=ens.dtype)
        if normalize:
            pdfs = pdfs / trapezoid(pdfs, grid).reshape(-1, 1)
    return pdfs


# --------------------------------------------------------------------------- #
# helpers shared by tests / bench (synthetic parameters)        #
# --------------------------------------------------------------------------- #
def cast_params(params, dtype):
    """Deep-cast a params pytree (for the float64 noise-floor twin)."""
    if isinstance(params, (list, tuple)):
        return type(params)(cast_params(p, dtype) for p in params)
    return np.asarray(params, dtype=dtype)


def synth_mlp_params(rng, in_dim, hidden_layers, hidden_dim, out_dim, bias_std=0.01,
                     out_scale=1.0):
    """stax-layout MLP parameters: W ~ N(0, 2/(fan_in+fan_out)), b ~ N(0, bias_std^2).
    Initialisation parity with jax.random is out of scope."""
    dims = [in_dim] + [hidden_dim] * hidden_layers + [out_dim]
    params = []
    for i in range(len(dims) - 1):
        fi, fo = dims[i], dims[i + 1]
        W = rng.normal(0.0, math.sqrt(2.0 / (fi + fo)), size=(fi, fo)).astype(np.float32)
        if i == len(dims) - 2:
            W = (W * out_scale).astype(np.float32)
        b = rng.normal(0.0, bias_std, size=(fo,)).astype(np.float32)
        params.append((W, b))
        if i < len(dims) - 2:
            params.append(())
    return params


def synth_bijector_params(info, input_dim, rng, out_scale=1.0):
    """Parameters pytree for a Bijector_Info (same nesting Chain.init_fun builds,
    bijectors.py:146-153)."""
    name, args = info
    if name == "Chain":
        return [synth_bijector_params(i, input_dim, rng, out_scale) for i in args]
    if name == "RollingSplineCoupling":
        return synth_bijector_params(expand_rolling(args), input_dim, rng, out_scale)
    if name == "NeuralSplineCoupling":
        K, B, hl, hd, td, nc, periodic = args
        upper, lower = _nsc_dims(input_dim, td)
        return synth_mlp_params(rng, upper + nc, hl, hd, (3 * K - 1 + int(periodic)) * lower,
                                out_scale=out_scale)
    return ()

"""NumPy restatement of the slice of ``jax.random`` pzflow's sampling path uses (TEST INFRASTRUCTURE).

jax / jaxlib (pinned 0.5.3, poetry.lock:1498-1533) are absent from /root/reference and not installable here, so the
published algorithm is restated: Threefry-2x32 with 20 rounds (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11 -- the Random123 ``threefry2x32_20`` known answers are checked in tests/test_oracle.py) and, on top
of it, what ``jax/_src/prng.py`` and ``jax/_src/random.py`` build from it:

    PRNGKey(seed)          -> key words (seed >> 32, seed & 0xffffffff); with x64 disabled (pzflow's setting) an int
                              seed is first narrowed to 32 bits, so the key is (0, seed mod 2**32)
    split(key, num)        -> threefry over iota(2 * num), reshaped [num, 2]
    random_bits(key, n)    -> threefry over iota(n): the count vector is cut in two HALVES (x0 = first half,
                              x1 = second half, a zero appended when n is odd), outputs concatenated
    uniform(key, shape, minval, maxval)
                           -> 23 mantissa bits | 1.0f, minus 1, scaled, clamped below at minval
    permutation(key, n)    -> ceil(3 ln n / ln(2**32 - 1)) rounds of (split, 32 random bits per element, stable sort)
    normal(key, shape)     -> sqrt(2) * erf_inv(uniform(key, shape, nextafter(-1, 0), 1)); erf_inv is XLA's float32 one:
                              the two degree-8 polynomials of M. Giles, "Approximating the erfinv function" (2011) in
                              w = -log1p(-x * x).  Pinned by the values jax's documentation prints ("JAX - The Sharp
                              Bits": normal(PRNGKey(0), (1,)) = -0.20584226, and -1.2515389 / -0.58665055 for the
                              sub-keys of its split chain; the PRNG tutorial: normal(PRNGKey(42)) = -0.18471177).

Two stream layouts exist in jax: the ORIGINAL one above (``jax_threefry_partitionable=False``, the default before
jax 0.5.0) and the "partitionable" one (element i draws threefry(key, (hi32(i), lo32(i))) and xors the two words; split
keeps both words of block i) -- the default from jax 0.5.0 on, hence of the jax the reference requires
(pyproject.toml:13, ``jax >= 0.5.3``).
**Pins, original layout:** ``uniform(PRNGKey(0), (2, 7), -5, 5)`` pushed through the shipped flow's inverse reproduces
the two samples the real JAX reference printed at docs/tutorials/marginalization.ipynb:133-135
(``flow.sample(2, seed=0)``, made with an older jax) to the six digits printed; ``split(PRNGKey(0))`` and
``normal`` of PRNGKey(0) / PRNGKey(42) and of the sub-keys of the split chains are the values older jax documentation
prints.  **Pins, partitionable layout:** the values current jax documentation prints -- ``normal(key(42))`` =
-0.028304616 and the draws 0.60576403 / -0.21089035 / -0.39489815 of its split loop, ``uniform(key(0))`` = 0.947667,
``normal(key(0), (1,))`` = 1.6226422.  All of them are asserted in tests/test_oracle.py.

Only tests/, oracle/make_golden_shim.py and oracle/jax_shim may import this module.
"""
from __future__ import annotations

import math

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return (x << _U32(d)) | (x >> _U32(32 - d))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds: counter words (x0, x1) under key (k0, k1) -> two output words, elementwise."""
    k0, k1 = _U32(int(key[0])), _U32(int(key[1]))
    x0 = np.array(x0, dtype=_U32, copy=True)
    x1 = np.array(x1, dtype=_U32, copy=True)
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey with x64 disabled: the seed narrowed to 32 bits, high word 0."""
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=_U32)


def _bits_original(key, n: int) -> np.ndarray:
    if n >= 2 ** 32 - 1:
        raise NotImplementedError("more than 2**32 - 2 words per call (jax cuts those into keyed blocks)")
    cnt = np.arange(n, dtype=_U32)
    if n % 2:
        cnt = np.concatenate([cnt, np.zeros(1, _U32)])
    h = len(cnt) // 2
    a, b = threefry2x32(key, cnt[:h], cnt[h:])
    return np.concatenate([a, b])[:n]


def _bits_partitionable(key, n: int) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint64)
    a, b = threefry2x32(key, (idx >> np.uint64(32)).astype(_U32), (idx & np.uint64(0xFFFFFFFF)).astype(_U32))
    return a ^ b


def random_bits(key, n: int, partitionable: bool = False) -> np.ndarray:
    """n 32-bit words in the order jax lays them out for a C-contiguous array of n elements."""
    return _bits_partitionable(key, n) if partitionable else _bits_original(key, n)


def split(key, num: int = 2, partitionable: bool = False) -> np.ndarray:
    """[num, 2] uint32 keys."""
    if partitionable:
        idx = np.arange(num, dtype=np.uint64)
        a, b = threefry2x32(key, (idx >> np.uint64(32)).astype(_U32), (idx & np.uint64(0xFFFFFFFF)).astype(_U32))
        return np.stack([a, b], axis=1)
    return _bits_original(key, 2 * num).reshape(num, 2)


def bits_to_uniform(bits: np.ndarray, minval: float, maxval: float) -> np.ndarray:
    lo, hi = np.float32(minval), np.float32(maxval)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(lo, f * (hi - lo) + lo)


def uniform(key, shape, minval=0.0, maxval=1.0, partitionable: bool = False) -> np.ndarray:
    n = int(np.prod(shape, dtype=np.int64))
    return bits_to_uniform(random_bits(key, n, partitionable), minval, maxval).reshape(shape)


def permutation(key, n: int, partitionable: bool = False) -> np.ndarray:
    """jax.random.permutation(key, jnp.arange(n)): rounds of sort-by-random-word."""
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(2 ** 32 - 1)))
    key = np.asarray(key, _U32)
    for _ in range(rounds):
        key, sub = split(key, 2, partitionable)
        x = x[np.argsort(random_bits(sub, n, partitionable), kind="stable")]
    return x


# Random123 known answers for threefry2x32_20 (counter, key) -> output, as jax's own tests use them
KAT = (
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x243F6A88, 0x85A308D3), (0x13198A2E, 0x03707344), (0xC4923A9C, 0x483DF7A0)),
)


_GILES_CENTRAL = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
                  -0.00417768164, 0.246640727, 1.50140941)
_GILES_TAIL = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
               0.00943887047, 1.00167406, 2.83297682)


def erf_inv32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ``erf_inv``: Giles' single-precision approximation, Horner in float32 without contraction."""
    f32 = np.float32
    x = np.asarray(x, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-x * x).astype(f32))).astype(f32)
        central = w < f32(5.0)
        w = np.where(central, w - f32(2.5), np.sqrt(w) - f32(3.0)).astype(f32)
        p = np.where(central, f32(_GILES_CENTRAL[0]), f32(_GILES_TAIL[0])).astype(f32)
        for a, b in zip(_GILES_CENTRAL[1:], _GILES_TAIL[1:]):
            p = (np.where(central, f32(a), f32(b)) + (p * w).astype(f32)).astype(f32)
        return (p * x).astype(f32)


def normal(key, shape, partitionable: bool = False) -> np.ndarray:
    """``jax.random.normal(key, shape)`` in float32 (jax/_src/random.py: _normal_real)."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, 1.0, partitionable)
    return (np.float32(np.sqrt(2.0)) * erf_inv32(u)).astype(np.float32)

"""Host side of the jax-compatible random streams: keys, ``split`` and the small draws made at construction time.

pzflow seeds everything through ``jax.random.PRNGKey`` (pzflow/flow.py:286-287, 186-187; distributions.py:464-469).
To return what the reference returns for a given ``seed`` -- the same ``Flow.sample`` rows from a Uniform latent, the
same ``Shuffle`` permutation -- the key schedule is reproduced here: Threefry-2x32/20 (Salmon et al., SC'11) in the
stream layout of jax's default PRNG implementation.  The bulk draws run on the GPU (``pzb_sample_uniform_threefry``,
csrc/pzb_api.cu); this module only derives keys and the handful of words a ``Chain`` needs when it is built
(pzflow/bijectors.py:146-149: ``rng, layer_rng = random.split(rng)`` per stage; 795-797: Shuffle's permutation).

Layouts (both pinned by numbers real JAX produced, tests/test_oracle.py):

``"threefry_partitionable"`` -- the DEFAULT -- is what ``jax_threefry_partitionable=True`` draws, the default of every
jax release from 0.5.0 on, i.e. of the jax the reference requires (pyproject.toml:13: ``jax >= 0.5.3``): element i of a
draw is threefry(key, (hi32(i), lo32(i))) with its two words xor-ed, ``split`` keeps both words.  Pinned by the values
current jax documentation prints (``random.normal(key(42))`` = -0.028304616 and the three draws of its split loop,
``random.uniform(key(0))`` = 0.947667, ``random.normal(key(0), (1,))`` = 1.6226422).

``"threefry"`` is jax's original layout (``jax_threefry_partitionable=False``, the default before 0.5.0): the count vector
rides through the block function as two halves.  It reproduces the samples the reference's authors printed in
docs/tutorials/marginalization.ipynb:133-135 (``flow.sample(2, seed=0)`` of the shipped flow, made with an older jax)
digit for digit, and the values older jax documentation prints (``split(PRNGKey(0))``, ``random.normal`` of PRNGKey(0)
/ PRNGKey(42)).  Select it with ``PZFLOW_B200_PRNG=threefry`` or ``set_layout("threefry")`` to re-draw what a pre-0.5
jax drew (one call of it holds at most 2**32 - 2 words; larger draws fall back to the Philox streams).

``"philox"`` keeps round 1's Philox-4x32-10 streams (not jax-compatible; any size).
"""
from __future__ import annotations

import os

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
LAYOUTS = ("threefry", "threefry_partitionable", "philox")
_layout = os.environ.get("PZFLOW_B200_PRNG", "threefry_partitionable")
if _layout not in LAYOUTS:
    raise ValueError(f"PZFLOW_B200_PRNG must be one of {LAYOUTS}")
# words of one original-layout call: jax cuts larger requests into separately keyed blocks, which is not reproduced
MAX_THREEFRY_WORDS = 2 ** 32 - 2


def set_layout(name: str) -> None:
    """Select the stream layout of ``Uniform.sample`` / ``Shuffle`` for this process."""
    global _layout
    if name not in LAYOUTS:
        raise ValueError(f"layout must be one of {LAYOUTS}")
    _layout = name


def layout() -> str:
    return _layout


def key_from_seed(seed) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` with x64 disabled: (0, seed mod 2**32).  A 2-word key passes through."""
    if isinstance(seed, np.ndarray) and seed.dtype == _U32 and seed.shape == (2,):
        return seed
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=_U32)


def _threefry(key, x0: np.ndarray, x1: np.ndarray):
    k0, k1 = _U32(int(key[0])), _U32(int(key[1]))
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    x0, x1 = x0.astype(_U32), x1.astype(_U32)
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = ((x1 << _U32(r)) | (x1 >> _U32(32 - r))) ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + (ks[(i + 2) % 3] + _U32(i + 1))
    return x0, x1


def words(key, n: int, partitionable: bool) -> np.ndarray:
    """n 32-bit words as jax's ``random_bits`` lays them out over n C-ordered elements (host side: small n only)."""
    if partitionable:
        idx = np.arange(n, dtype=np.uint64)
        a, b = _threefry(key, idx >> np.uint64(32), idx & np.uint64(0xFFFFFFFF))
        return a ^ b
    cnt = np.arange(n + (n & 1), dtype=np.uint64)
    if n & 1:
        cnt[-1] = 0  # jax pads an odd count vector with a zero
    h = len(cnt) // 2
    a, b = _threefry(key, cnt[:h], cnt[h:])
    return np.concatenate([a, b])[:n]


def split(key, num: int = 2, layout_name: str = None) -> np.ndarray:
    """``jax.random.split``: [num, 2] keys."""
    part = (layout_name or _layout) == "threefry_partitionable"
    if part:
        idx = np.arange(num, dtype=np.uint64)
        a, b = _threefry(key, idx >> np.uint64(32), idx & np.uint64(0xFFFFFFFF))
        return np.stack([a, b], axis=1)
    return words(key, 2 * num, False).reshape(num, 2)


def permutation(key, n: int, layout_name: str = None) -> np.ndarray:
    """``jax.random.permutation(key, jnp.arange(n))``: rounds of (split, one word per element, stable sort)."""
    part = (layout_name or _layout) == "threefry_partitionable"
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(2 ** 32 - 1)))
    key = key_from_seed(key)
    for _ in range(rounds):
        key, sub = split(key, 2, "threefry_partitionable" if part else "threefry")
        x = x[np.argsort(words(sub, n, part), kind="stable")]
    return x


def epoch_orders(batch_seed: int, n: int, layout_name: str = None):
    """The row order of every training epoch as the reference draws it (pzflow/flow.py:1015, 1062-1064):
    ``key = PRNGKey(batch_seed)``; per epoch ``permute_key, sample_key, key = split(key, 3)`` and
    ``idx = permutation(permute_key, n)``.  An endless generator of index arrays."""
    key = key_from_seed(int(batch_seed))
    while True:
        permute_key, _, key = split(key, 3, layout_name)
        yield permutation(permute_key, n, layout_name)

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


ZERO_PROB = 1e30  # |lp| above this is the zero-probability class {NaN, -inf, finfo.min}


def zero_class(a):
    a = np.asarray(a, dtype=np.float64)
    return np.isnan(a) | (a < -ZERO_PROB)


# jax's two stream layouts (pzflow_b200/prng.py, oracle/jax_prng.py): the reference's required jax (>= 0.5.3) draws the
# partitionable one -- the library's default --, the notebooks in its tree were made with an older jax (original layout)
STREAM_GOLDEN = {"threefry": "ref_shim_jax_streams.npz", "threefry_partitionable": "ref_shim_jax_streams_part.npz"}


def _with_layout(name):
    from pzflow_b200 import prng
    was = prng.layout()
    prng.set_layout(name)
    try:
        yield name
    finally:
        prng.set_layout(was)


@pytest.fixture(params=["threefry", "threefry_partitionable"])
def stream_layout(request):
    """Runs the test once per jax stream layout, with the library switched to it."""
    yield from _with_layout(request.param)


@pytest.fixture
def notebook_era_streams():
    """jax's original stream layout: what produced the numbers printed in the reference's notebooks."""
    yield from _with_layout("threefry")

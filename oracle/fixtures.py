"""Parameter sets and synthetic inputs of the BASELINE.json configs (test/bench infrastructure).

``example_flow`` is the reference's shipped 7-D galaxy flow
(pzflow/example_files/example-flow.pzflow.pkl), re-saved as plain arrays by
``oracle/make_golden.py`` because /root/reference does not exist on the GPU box.
The other configs use synthetic stax-layout parameters.
"""
from __future__ import annotations

import os

import numpy as np

from . import nsf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GALAXY_COLUMNS = ("redshift", "u", "g", "r", "i", "z", "y")


def rolling_info(nlayers, shift=1, K=16, B=5, hidden_layers=2, hidden_dim=128, transformed_dim=None,
                 n_conditions=0, periodic=False):
    """Bijector_Info of RollingSplineCoupling (pzflow/bijectors.py:651-665)."""
    return O.expand_rolling((nlayers, shift, K, B, hidden_layers, hidden_dim, transformed_dim,
                             n_conditions, periodic))


def example_flow():
    """The shipped flow: Chain(ShiftBounds(mins, maxs, 4), RollingSplineCoupling(7)), Uniform(7, 5)."""
    z = np.load(os.path.join(GOLDEN, "example_flow_params.npz"))
    nl = int(z["n_layers"])
    info = ("Chain", (("ShiftBounds", (z["mins"], z["maxs"], int(z["sb_B"]))), rolling_info(nl)))
    inner = []
    for i in range(nl):
        layer = []
        for j in range(3):
            layer.append((z[f"l{i}_W{j}"], z[f"l{i}_b{j}"]))
            if j < 2:
                layer.append(())
        inner += [layer, ()]
    params = ((), [(), inner])
    return dict(bijector_info=info, latent_info=("Uniform", (7, 5)), params=params,
                data_columns=GALAXY_COLUMNS, input_dim=7, n_conditions=0)


def default_flow(input_dim, mins, maxs, n_conditions=0, seed=0, latent="CentBeta13", K=16,
                 hidden_dim=128, nlayers=None, sb_B=4.0, out_scale=1.0):
    """The reference's default bijector (pzflow/flow.py:314-322) with synthetic parameters."""
    nlayers = input_dim if nlayers is None else nlayers
    info = ("Chain", (("ShiftBounds", (np.asarray(mins, np.float32), np.asarray(maxs, np.float32), sb_B)),
                      rolling_info(nlayers, K=K, hidden_dim=hidden_dim, n_conditions=n_conditions)))
    rng = np.random.default_rng(seed)
    bparams = O.synth_bijector_params(info, input_dim, rng, out_scale=out_scale)
    return dict(bijector_info=info, latent_info=(latent, (input_dim, 5)), params=((), bparams),
                input_dim=input_dim, n_conditions=n_conditions)


def config_c1(seed=0):
    """two-moons 2-D default Flow, K=16, CentBeta13(2,5)."""
    return default_flow(2, [-1.6, -1.1], [2.6, 1.6], seed=seed)


def config_c4(n_flows=4):
    """7-D + 2 conditions, FlowEnsemble of 4 (seeds 0..3)."""
    ex = example_flow()
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    return [default_flow(7, mins, maxs, n_conditions=2, seed=s) for s in range(n_flows)]


def config_c5(seed=0):
    """32-D, K=32, hidden 256, 32 layers, CentBeta13(32,5)."""
    return default_flow(32, -np.ones(32), np.ones(32), seed=seed, K=32, hidden_dim=256)


def galaxy_rows(n, seed=0, oob_frac=0.01):
    """C2 inputs: x = inverse(u), u ~ U(-5,5)^7, plus a 1 % admixture of rows drawn
    from the ShiftBounds box widened by 20 % (exercises the out-of-bounds paths)."""
    ex = example_flow()
    rng = np.random.default_rng(seed)
    U = rng.uniform(-5, 5, size=(n, 7)).astype(np.float32)
    X, _ = O.flow_inverse(ex["bijector_info"], ex["params"], U)
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    n_oob = int(round(n * oob_frac))
    if n_oob:
        rows = rng.choice(n, size=n_oob, replace=False)
        span = maxs - mins
        X[rows] = rng.uniform(mins - 0.2 * span, maxs + 0.2 * span, size=(n_oob, 7)).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)


# --------------------------------------------------------------------------- #
# (de)serialisation of Bijector_Info / params pytrees into .npz golden files  #
# --------------------------------------------------------------------------- #
def _enc(obj, leaves):
    import numbers
    if isinstance(obj, (list, tuple)):
        return {"t": "list" if isinstance(obj, list) else "tuple", "v": [_enc(o, leaves) for o in obj]}
    if obj is None or isinstance(obj, (bool, str)):
        return {"t": "py", "v": obj}
    if isinstance(obj, numbers.Integral):
        return {"t": "int", "v": int(obj)}
    if isinstance(obj, numbers.Real) and not isinstance(obj, np.ndarray):
        return {"t": "float", "v": float(obj)}
    arr = np.asarray(obj)
    leaves.append(arr)
    return {"t": "arr", "v": len(leaves) - 1}


def _dec(node, leaves):
    t, v = node["t"], node["v"]
    if t == "list":
        return [_dec(o, leaves) for o in v]
    if t == "tuple":
        return tuple(_dec(o, leaves) for o in v)
    if t == "arr":
        return leaves[v]
    return v


def pack_tree(prefix, obj, store):
    """Store a nested list/tuple/array/scalar structure into ``store`` (a dict for np.savez)."""
    import json
    leaves = []
    store[f"{prefix}__json"] = np.array(json.dumps(_enc(obj, leaves)))
    for i, a in enumerate(leaves):
        store[f"{prefix}__leaf{i}"] = a


def unpack_tree(prefix, z):
    import json
    node = json.loads(str(z[f"{prefix}__json"]))
    leaves = []
    i = 0
    while f"{prefix}__leaf{i}" in z:
        leaves.append(np.asarray(z[f"{prefix}__leaf{i}"]))
        i += 1
    return _dec(node, leaves)

"""Golden vectors from the reference's own code over oracle/jax_shim (build container only).

Each tests/golden/ref_shim_*.npz holds inputs, the Bijector_Info / params the
reference built, and what the reference's functions returned.  The oracle and
the CUDA path are then checked against these files (tests/test_oracle.py,
tests/test_gpu_parity.py).  See oracle/jax_shim/__init__.py for what this pins.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import fixtures as F
from . import jax_shim

GOLDEN = F.GOLDEN


def _save(name, store):
    path = os.path.join(GOLDEN, f"ref_shim_{name}.npz")
    np.savez_compressed(path, **store)
    print("wrote", os.path.basename(path), f"{os.path.getsize(path) / 1024:.0f} KiB")


def _np(x):
    return np.asarray(x)


def case_example_flow(pzflow):
    """Shipped 7-D flow through Flow.log_prob / _forward / _inverse / posterior (tests/test_examples.py:31-40)."""
    from pzflow import examples
    import jax.numpy as jnp
    flow = examples.get_example_flow()
    X = F.galaxy_rows(384, seed=7)
    # edge rows: far out of bounds, exactly on the ShiftBounds box, NaN
    mins, maxs = _np(flow._bijector_info[1][0][1][0]), _np(flow._bijector_info[1][0][1][1])
    X[0] = maxs
    X[1] = mins
    X[2] = maxs + 3 * (maxs - mins)
    X[3, 4] = np.nan
    X[4] = 0.5 * (mins + maxs)
    df = pd.DataFrame(X, columns=flow.data_columns)
    store = dict(X=X)
    store["log_prob"] = _np(flow.log_prob(df))
    cond = jnp.zeros((X.shape[0], 1))
    u, ld = flow._forward(flow._params[1], jnp.array(X), conditions=cond)
    store["forward_u"], store["forward_ld"] = _np(u), _np(ld)
    U = np.random.default_rng(11).uniform(-5, 5, size=(384, 7)).astype(np.float32)
    U[0] = 5.0
    U[1] = -5.0
    U[2] = 7.5
    x, ldi = flow._inverse(flow._params[1], jnp.array(U), conditions=cond)
    store["U"], store["inverse_x"], store["inverse_ld"] = U, _np(x), _np(ldi)
    grid = np.linspace(0.0, 2.3, 60).astype(np.float32)
    store["grid"] = grid
    store["posterior"] = _np(flow.posterior(df.iloc[:24], column="redshift", grid=jnp.array(grid)))
    store["posterior_batched"] = _np(flow.posterior(df.iloc[:24], column="redshift", grid=jnp.array(grid), batch_size=7))
    store["posterior_unnorm"] = _np(flow.posterior(df.iloc[:24], column="redshift", grid=jnp.array(grid),
                                                   normalize=False, nan_to_zero=False))
    grid_r = np.linspace(14.0, 30.0, 33).astype(np.float32)
    store["grid_r"] = grid_r
    store["posterior_r"] = _np(flow.posterior(df.iloc[:24], column="r", grid=jnp.array(grid_r)))
    # the two samples printed at docs/tutorials/marginalization.ipynb:133-135 (real-JAX outputs)
    Xs = np.array([[0.386102, 28.406404, 27.508736, 26.419685, 26.196583, 25.901255, 25.889046],
                   [2.118688, 28.220659, 27.798203, 27.371502, 27.211376, 26.869678, 26.502558]], np.float32)
    store["nb_X"] = Xs
    store["nb_log_prob"] = _np(flow.log_prob(pd.DataFrame(Xs, columns=flow.data_columns)))
    _save("example_flow", store)


def case_bijectors(pzflow):
    """The hot-path rows of tests/test_bijectors.py:16-51 on its fixed 3x7 input."""
    from pzflow import bijectors as B
    import jax.numpy as jnp
    from jax import random
    x = np.array([[0.2, 0.1, -0.3, 0.5, 0.1, -0.4, -0.3],
                  [0.6, 0.5, 0.2, 0.2, -0.4, -0.1, 0.7],
                  [0.9, 0.2, -0.3, 0.3, 0.4, -0.4, -0.1]], np.float32)
    cases = [
        ("standard_scaler", B.StandardScaler, (jnp.linspace(-1, 1, 7), jnp.linspace(1, 8, 7)), np.zeros((3, 1))),
        ("nsc_default", B.NeuralSplineCoupling, (), np.zeros((3, 1))),
        ("nsc_cond", B.NeuralSplineCoupling, (16, 3, 2, 128, 3), np.arange(9).reshape(3, 3)),
        ("rsc2", B.RollingSplineCoupling, (2,), np.zeros((3, 1))),
        ("rsc_periodic", B.RollingSplineCoupling, (2, 1, 16, 3, 2, 128, None, 0, True), np.zeros((3, 1))),
        ("shift_scalar", B.ShiftBounds, (-0.5, 0.9, 5), np.zeros((3, 1))),
        ("shift_vector", B.ShiftBounds, (-1 * jnp.ones(7), 1.1 * jnp.ones(7), 3), np.zeros((3, 1))),
        ("roll2", B.Roll, (2,), np.zeros((3, 1))),
        ("chain_mixed", B.Chain, (B.ShiftBounds(-1 * jnp.ones(7), 1.1 * jnp.ones(7), 3), B.Roll(-1),
                                  B.RollingSplineCoupling(3, 2, 8, 3, 1, 32), B.StandardScaler(jnp.linspace(-1, 1, 7), jnp.linspace(1, 8, 7))),
         np.zeros((3, 1))),
    ]
    store = dict(x=x, names=np.array([c[0] for c in cases]))
    for name, bij, args, cond in cases:
        if name == "nsc_cond":
            # conditional test row uses n_conditions = 3 (tests/test_bijectors.py:33-37 passes it as 6th arg
            # implicitly 0 -> conditions ignored); exercise the real conditional path here
            args = (16, 3, 2, 128, 3, 3)
        init_fun, info = bij(*args)
        params, fwd, inv = init_fun(random.PRNGKey(0), 7)
        # widen the conditioner outputs so the splines are far from identity
        cond = np.asarray(cond, dtype=np.float32)
        y, ld = fwd(params, jnp.array(x), conditions=jnp.array(cond))
        xi, ldi = inv(params, y, conditions=jnp.array(cond))
        xi2, ldi2 = inv(params, jnp.array(x), conditions=jnp.array(cond))
        F.pack_tree(f"{name}_info", info, store)
        F.pack_tree(f"{name}_params", params, store)
        store[f"{name}_cond"] = cond
        store[f"{name}_fwd"], store[f"{name}_fwd_ld"] = _np(y), _np(ld)
        store[f"{name}_inv_of_fwd"], store[f"{name}_inv_of_fwd_ld"] = _np(xi), _np(ldi)
        store[f"{name}_inv"], store[f"{name}_inv_ld"] = _np(xi2), _np(ldi2)
    _save("bijectors", store)


def case_extra_bijectors(pzflow):
    """the parameter-free rows of tests/test_bijectors.py:16-51 (ColorTransform, Reverse,
    Scale, InvSoftplus, their Chain) plus a flow-shaped Chain, and the Normal / CentBeta latents."""
    from pzflow import bijectors as B
    from pzflow import distributions as Dn
    import jax.numpy as jnp
    from jax import random
    x = np.array([[0.2, 0.1, -0.3, 0.5, 0.1, -0.4, -0.3],
                  [0.6, 0.5, 0.2, 0.2, -0.4, -0.1, 0.7],
                  [0.9, 0.2, -0.3, 0.3, 0.4, -0.4, -0.1]], np.float32)
    rng = np.random.default_rng(5)
    xr = np.vstack([x, rng.uniform(0.05, 1.5, size=(61, 7)).astype(np.float32)])
    cases = [
        ("color", B.ColorTransform, (3, [1, 3, 5])),
        ("color_first", B.ColorTransform, (1, [1, 2, 4, 5, 6])),
        ("reverse", B.Reverse, ()),
        ("scale", B.Scale, (2.0,)),
        ("invsoftplus0", B.InvSoftplus, (0,)),
        ("invsoftplus2", B.InvSoftplus, ([1, 3], [2.0, 12.0])),
        ("chain_free", B.Chain, (B.Reverse(), B.Scale(1 / 6), B.Roll(-1))),
        ("chain_flow", B.Chain, (B.ColorTransform(3, [1, 2, 3, 4, 5, 6]), B.InvSoftplus(0, 4.0),
                                 B.StandardScaler(jnp.linspace(-0.5, 0.5, 7), jnp.linspace(0.5, 2.0, 7)),
                                 B.Reverse(), B.RollingSplineCoupling(3, 1, 8, 6, 1, 32), B.Scale(0.5))),
    ]
    store = dict(x=xr, names=np.array([c[0] for c in cases]))
    cond = np.zeros((xr.shape[0], 1), np.float32)
    for name, bij, args in cases:
        init_fun, info = bij(*args)
        params, fwd, inv = init_fun(random.PRNGKey(0), 7)
        y, ld = fwd(params, jnp.array(xr), conditions=jnp.array(cond))
        xi, ldi = inv(params, y, conditions=jnp.array(cond))
        xi2, ldi2 = inv(params, jnp.array(xr), conditions=jnp.array(cond))
        F.pack_tree(f"{name}_info", info, store)
        F.pack_tree(f"{name}_params", params, store)
        store[f"{name}_fwd"], store[f"{name}_fwd_ld"] = _np(y), _np(ld)
        store[f"{name}_inv_of_fwd"], store[f"{name}_inv_of_fwd_ld"] = _np(xi), _np(ldi)
        store[f"{name}_inv"], store[f"{name}_inv_ld"] = _np(xi2), _np(ldi2)
    # latents (tests/test_distributions.py rows for Normal and CentBeta)
    u = rng.uniform(-5.5, 5.5, size=(200, 3)).astype(np.float32)
    store["latent_u"] = u
    normal = Dn.Normal(3)
    store["normal_lp"] = _np(normal.log_prob(normal._params, jnp.array(u)))
    cb = Dn.CentBeta(3, 5)
    cb_params = tuple((float(a), float(b)) for a, b in rng.normal(0.8, 0.6, size=(3, 2)))
    store["centbeta_params"] = np.asarray(cb_params, np.float32)
    store["centbeta_lp"] = _np(cb.log_prob(cb_params, jnp.array(u)))
    store["centbeta_lp_default"] = _np(cb.log_prob(cb._params, jnp.array(u)))
    td = Dn.Tdist(3)
    store["tdist_lp_default"] = _np(td.log_prob(td._params, jnp.array(u)))
    store["tdist_lognu"] = np.float32(np.log(4.5))
    store["tdist_lp"] = _np(td.log_prob(jnp.log(jnp.array(4.5)), jnp.array(u)))
    joint = Dn.Joint(Dn.Uniform(1, 5), Dn.CentBeta(1, 5), Dn.Normal(1))
    jp = [(), (cb_params[0],), ()]
    store["joint_lp"] = _np(joint.log_prob(jp, jnp.array(u)))
    F.pack_tree("joint_info", joint.info, store)
    _save("extra", store)


def _boost(params, rng, scale=3.0):
    """Scale the last Dense layer of every conditioner so splines are strongly non-identity."""
    if isinstance(params, list) and len(params) and isinstance(params[0], tuple) and len(params[0]) == 2 \
            and getattr(params[0][0], "ndim", 0) == 2:
        W, b = params[-1]
        params[-1] = (np.asarray(W) * np.float32(scale), rng.normal(0, 0.5, size=np.asarray(b).shape).astype(np.float32))
        return params
    if isinstance(params, (list, tuple)):
        out = [_boost(p, rng, scale) for p in params]
        return out if isinstance(params, list) else tuple(out)
    return params


def case_default_flows(pzflow):
    """Default bijector (flow.py:296-322): 2-D two-moons / CentBeta13 and a conditional 3-D + 2 flow, plus an ensemble."""
    from pzflow import Flow, FlowEnsemble
    from pzflow.examples import get_twomoons_data
    import jax.numpy as jnp
    rng = np.random.default_rng(5)
    store = {}
    # ---- C1: two moons ----
    data = get_twomoons_data().iloc[:512].reset_index(drop=True)
    flow = Flow(data_columns=("x", "y"))
    flow._set_default_bijector(get_twomoons_data(), seed=0)
    flow._params = (flow._params[0], _boost(flow._params[1], rng))
    X = data[["x", "y"]].to_numpy().astype(np.float32)
    store["c1_X"] = X
    F.pack_tree("c1_info", flow._bijector_info, store)
    F.pack_tree("c1_params", flow._params[1], store)
    store["c1_log_prob"] = _np(flow.log_prob(data))
    grid = np.linspace(-1.5, 2.5, 41).astype(np.float32)
    store["c1_grid"] = grid
    store["c1_posterior_x"] = _np(flow.posterior(data.iloc[:32], column="x", grid=jnp.array(grid)))
    store["c1_posterior_y"] = _np(flow.posterior(data.iloc[:32], column="y", grid=jnp.array(grid), normalize=False))
    u = flow.latent.sample(flow._params[0], 256, 3)
    xs = flow._inverse(flow._params[1], u, conditions=jnp.zeros((256, 1)))[0]
    store["c1_sample_u"], store["c1_sample_x"] = _np(u), _np(xs)

    # ---- conditional flow + ensemble ----
    n = 200
    df = pd.DataFrame({"a": rng.normal(0, 1, n), "b": rng.uniform(-2, 3, n), "c": rng.normal(1, 2, n),
                       "p": rng.normal(5, 2, n), "q": rng.uniform(0, 1, n)})
    ens = FlowEnsemble(data_columns=("a", "b", "c"), conditional_columns=("p", "q"), N=3)
    for i, fl in enumerate(ens._ensemble.values()):
        fl._set_default_bijector(df, seed=i)
        fl._params = (fl._params[0], _boost(fl._params[1], rng, 2.0))
        fl._condition_means = jnp.array(df[["p", "q"]].to_numpy().mean(axis=0))
        fl._condition_stds = jnp.array(df[["p", "q"]].to_numpy().std(axis=0))
    store["ens_df"] = df.to_numpy().astype(np.float64)
    store["ens_columns"] = np.array(list(df.columns))
    f0 = ens._ensemble["Flow 0"]
    store["ens_cond_means"], store["ens_cond_stds"] = _np(f0._condition_means), _np(f0._condition_stds)
    for i, fl in enumerate(ens._ensemble.values()):
        F.pack_tree(f"ens{i}_info", fl._bijector_info, store)
        F.pack_tree(f"ens{i}_params", fl._params[1], store)
    store["ens_flow0_log_prob"] = _np(f0.log_prob(df))
    store["ens_log_prob"] = _np(ens.log_prob(df))
    store["ens_log_prob_all"] = _np(ens.log_prob(df, returnEnsemble=True))
    g = np.linspace(-3, 3, 25).astype(np.float32)
    store["ens_grid"] = g
    store["ens_posterior"] = _np(ens.posterior(df.iloc[:20], column="a", grid=jnp.array(g)))
    store["ens_posterior_all"] = _np(ens.posterior(df.iloc[:20], column="a", grid=jnp.array(g), returnEnsemble=True))
    store["ens_flow0_posterior_b"] = _np(f0.posterior(df.iloc[:20], column="b", grid=jnp.array(g)))
    # conditional sampling map: u -> x with 2 copies of each condition row (flow.py:787-795)
    cdf = df.iloc[:10]
    conds = jnp.repeat(f0._get_conditions(cdf), 2, axis=0)
    u = f0.latent.sample(f0._params[0], 20, 1)
    store["ens_sample_u"], store["ens_sample_cond"] = _np(u), _np(conds)
    store["ens_sample_x"] = _np(f0._inverse(f0._params[1], u, conditions=conds)[0])
    _save("default_flows", store)


def case_marg_rules(pzflow):
    """Flow.posterior(marg_rules=...) of the shipped flow (pzflow/flow.py:560-664; docs/tutorials/marginalization.ipynb
    cells 12-14): a numeric flag and a NaN flag, rules whose grids depend on the row, a row with TWO flagged columns
    (nested marginalisation) and a row whose flagged column has no rule (stays all-zero)."""
    from pzflow import examples
    import jax.numpy as jnp
    flow = examples.get_example_flow()
    X = F.galaxy_rows(40, seed=21, oob_frac=0.0).astype(np.float64)
    grid = np.linspace(0.0, 2.3, 46).astype(np.float32)
    store = dict(X=X, grid=grid)

    def rule_u(row):
        return np.linspace(25.0, 31.0, 12)

    def rule_y(row):
        return np.linspace(row["z"] - 1.0, row["z"] + 1.0, 9)

    for tag, flag in (("f99", 99.0), ("nan", np.nan)):
        df = pd.DataFrame(X.copy(), columns=flow.data_columns)
        df.loc[[1, 5, 9, 13], "u"] = flag
        df.loc[[2, 5, 20], "y"] = flag      # row 5: u AND y are missing
        df.loc[[30], "g"] = flag            # no rule for g: the row is never evaluated
        rules = {"flag": flag, "u": rule_u, "y": rule_y}
        store[f"{tag}_inputs"] = df.to_numpy()
        store[f"{tag}_posterior"] = _np(flow.posterior(df, column="redshift", grid=jnp.array(grid), marg_rules=rules))
        store[f"{tag}_posterior_unnorm"] = _np(flow.posterior(df, column="redshift", grid=jnp.array(grid), marg_rules=rules,
                                                              normalize=False, nan_to_zero=False))
    store["u_grid"] = rule_u(None)
    store["y_halfwidth"], store["y_points"] = np.float64(1.0), np.int64(9)
    _save("marg_rules", store)


def case_ensemble_sample(pzflow):
    """FlowEnsemble.sample, unconditional (pzflow/flowEnsemble.py:397-408): ceil(n / n_flows) samples from every flow
    with the same seed, concatenated, then DataFrame.sample(n, random_state=seed).reset_index(drop=True).  The per-flow
    draws come from the jax PRNG (here: the stand-in's generator) and are stored, so the composition -- which rows of
    which flow end up where -- is pinned given those draws."""
    from pzflow import FlowEnsemble
    rng = np.random.default_rng(9)
    n = 200
    df = pd.DataFrame({"a": rng.normal(0, 1, n), "b": rng.uniform(-2, 3, n), "c": rng.normal(1, 2, n)})
    ens = FlowEnsemble(data_columns=("a", "b", "c"), N=3)
    for i, fl in enumerate(ens._ensemble.values()):
        fl._set_default_bijector(df, seed=i)
    store = {}
    nsamples, seed = 10, 4
    per_flow = [fl.sample(int(np.ceil(nsamples / 3)), None, True, seed) for fl in ens._ensemble.values()]
    store["per_flow"] = np.stack([p.to_numpy() for p in per_flow]).astype(np.float32)
    out = ens.sample(nsamples, seed=seed)
    store["nsamples"], store["seed"] = np.int64(nsamples), np.int64(seed)
    store["sample"] = out.to_numpy().astype(np.float32)
    store["sample_index"] = out.index.to_numpy()
    both = ens.sample(4, seed=seed, returnEnsemble=True)
    store["return_ensemble_shape"] = np.asarray(both.shape)
    store["return_ensemble_keys"] = np.array([k for k, _ in both.index])
    _save("ensemble_sample", store)


def case_jax_streams(pzflow, name="jax_streams"):
    """Round 2 (second session), stand-in installed with prng="threefry" (-> ref_shim_jax_streams.npz) or
    "threefry_partitionable" (-> ref_shim_jax_streams_part.npz): jax's own key schedule and uniform stream, oracle/jax_prng.py:
    the reference's Flow.sample for seeds (pzflow/flow.py:783-795 -> distributions.py:464-469), and flows whose chains
    hold Shuffle stages (bijectors.py:780-814; permutation from the key Chain.init_fun hands down, 146-149) and a
    UniformDequantizer (865-911), built with two seeds and re-built from a save dictionary (flow.py:186-187: PRNGKey(0))."""
    from pzflow import Flow, examples
    from pzflow import bijectors as B
    from pzflow import distributions as Dn
    import jax.numpy as jnp
    flow = examples.get_example_flow()
    store = {}
    store["example_sample_2_seed0"] = flow.sample(2, seed=0).to_numpy().astype(np.float32)
    store["example_latent_1001_seed7"] = _np(flow.latent.sample((), 1001, 7))
    store["example_sample_1001_seed7"] = flow.sample(1001, seed=7).to_numpy().astype(np.float32)
    rng = np.random.default_rng(5)
    n, D = 64, 5
    X = rng.uniform(-1.0, 1.0, size=(n, D)).astype(np.float32)
    X[:, 2] = np.floor(X[:, 2] * 4)        # a discrete column for the dequantizer
    mins, maxs = X.min(axis=0) - 0.5, X.max(axis=0) + 1.5
    cols = tuple("abcde")
    df = pd.DataFrame(X.astype(np.float64), columns=cols)
    store["X"], store["mins"], store["maxs"] = X, mins, maxs

    def bijector():
        return B.Chain(B.UniformDequantizer([2]), B.ShiftBounds(jnp.array(mins), jnp.array(maxs), 4.0), B.Shuffle(),
                       B.RollingSplineCoupling(3, K=8, hidden_dim=32), B.Chain(B.Reverse(), B.Shuffle()),
                       B.RollingSplineCoupling(2, K=8, hidden_dim=32))
    for seed in (0, 3):
        fl = Flow(cols, bijector(), latent=Dn.Uniform(D, 5), seed=seed)
        _boost(fl._params[1], np.random.default_rng(40 + seed))
        tag = f"s{seed}"
        if seed == 0:
            F.pack_tree("shuffle_info", fl._bijector_info, store)
        F.pack_tree(f"{tag}_params", fl._params[1], store)
        store[f"{tag}_log_prob"] = _np(fl.log_prob(df))
        u, ld = fl._forward(fl._params[1], jnp.array(X), conditions=jnp.zeros((n, 1)))
        store[f"{tag}_forward_u"], store[f"{tag}_forward_ld"] = _np(u), _np(ld)
        store[f"{tag}_sample_9_seed5"] = fl.sample(9, seed=5).to_numpy().astype(np.float32)
        if seed == 3:
            again = Flow(_dictionary=fl._save_dict())     # what Flow(file=...) does after unpickling
            store["s3_reloaded_log_prob"] = _np(again.log_prob(df))
    # FlowEnsemble.sample with conditions (pzflow/flowEnsemble.py:409-471): more rows than flows -> the conditions are
    # shuffled (pandas, random_state = seed / 1e9), split into n_flows chunks, the CHUNKS permuted by
    # jax.random.permutation(PRNGKey(seed), ...), flow i draws chunk i; fewer rows than flows -> NumPy picks a flow and a
    # seed per row.  Uniform latents, so the draws themselves are jax's stream.
    from pzflow import FlowEnsemble
    # flowEnsemble.py:431 calls np.array_split on a DataFrame; pandas >= 2.1 with NumPy 2 returns bare arrays there (older
    # releases returned positional DataFrame chunks, which is what the code expects): restore that for the call
    _split = np.array_split
    np.array_split = lambda a, n, axis=0: ([a.iloc[i] for i in _split(np.arange(len(a)), n)] if isinstance(a, pd.DataFrame)
                                           else _split(a, n, axis))
    ccols = ("p", "q")
    ens = FlowEnsemble(("a", "b", "c"), B.Chain(B.ShiftBounds(jnp.array(mins[:3]), jnp.array(maxs[:3]), 4.0),
                                                B.RollingSplineCoupling(3, K=8, hidden_dim=32, n_conditions=2)),
                       latent=Dn.Uniform(3, 5), conditional_columns=ccols, N=3)
    for i, fl in enumerate(ens._ensemble.values()):
        _boost(fl._params[1], np.random.default_rng(60 + i))
        F.pack_tree(f"cens{i}_params", fl._params[1], store)
    F.pack_tree("cens_info", ens._ensemble["Flow 0"]._bijector_info, store)
    conds = pd.DataFrame(rng.normal(size=(7, 2)), columns=ccols, index=[3, 1, 4, 15, 9, 2, 6])
    store["cens_conditions"], store["cens_index"] = conds.to_numpy(), conds.index.to_numpy()
    for tag, frame, ns, seed in (("many", conds, 2, 12), ("few", conds.iloc[:2], 1, 5), ("exact", conds.iloc[:3], 1, 7)):
        out = ens.sample(ns, conditions=frame, save_conditions=True, seed=seed)
        store[f"cens_{tag}_sample"] = out.to_numpy().astype(np.float32)
        store[f"cens_{tag}_index"] = out.index.to_numpy()
        store[f"cens_{tag}_args"] = np.array([ns, seed, len(frame)])
    np.array_split = _split
    _save(name, store)


def main():
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "jax_streams":
        case_jax_streams(jax_shim.install(prng="threefry"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "jax_streams_part":
        case_jax_streams(jax_shim.install(prng="threefry_partitionable"), "jax_streams_part")
        return
    pzflow = jax_shim.install()
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        case_marg_rules(pzflow)
        case_ensemble_sample(pzflow)
        return
    case_example_flow(pzflow)
    case_bijectors(pzflow)
    case_extra_bijectors(pzflow)
    case_default_flows(pzflow)


if __name__ == "__main__":
    main()

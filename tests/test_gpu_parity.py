"""Parity of the CUDA path (through the C-ABI) against the oracle and the committed golden
vectors produced by the reference's own code.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pandas as pd
import pytest
import torch

from conftest import golden
from oracle import fixtures as F
from oracle import nsf_oracle as O
from parity import (LP_UNDERFLOW, assert_noise_floor, assert_tolerance, count_bin_flips, summarize,
                    within_tol, zero_class)

pytestmark = pytest.mark.gpu

ENGINES = ["simt", "tc"]


def _plan(cfg, engine):
    from pzflow_b200.plan import Plan
    plan = Plan(cfg["bijector_info"], cfg["params"][1], cfg["input_dim"], cfg["latent_info"])
    if engine == "tc" and not plan.tc_supported:
        pytest.skip("tensor-core engine does not support this shape")
    if engine == "wide" and not plan.wide_supported:
        pytest.skip("wide tensor-core engine does not support this shape")
    plan.set_engine(engine)
    assert plan.engine == engine
    return plan


def _o64(cfg):
    return cfg["bijector_info"], cfg["latent_info"], O.cast_params(cfg["params"], np.float64)


def _np(t):
    return t.detach().cpu().numpy()


def _same_bits(a, b):
    """Bit-for-bit equality (NaN == NaN: idx = K rows carry NaN through forward / inverse)."""
    return torch.equal(a.cpu().contiguous().view(torch.int32), b.cpu().contiguous().view(torch.int32))


# --------------------------------------------------------------------------- #
# the shipped 7-D galaxy flow (configs 2/3): noise-floor gate + golden vectors  #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("engine", ENGINES)
def test_example_flow_log_prob_noise_floor(engine):
    ex = F.example_flow()
    plan = _plan(ex, engine)
    X = F.galaxy_rows(20000, seed=0)
    ob = []
    o32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X, bins=ob)
    bi, li, p64 = _o64(ex)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), return_bins=True)
    s = assert_noise_floor(f"C2 log_prob [{engine}]", _np(lp), o32, o64, below=LP_UNDERFLOW)
    assert s["frac_within_tol"] > 0.4  # reported, not the gate: see tests/parity.py
    # bin indices: bit-exact except knot ties, which are counted
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    print(f"BINS [{engine}] flipped {flips} of {total} ({rows} rows)")
    assert flips <= max(3, int(2e-4 * total))


@pytest.mark.parametrize("engine", ENGINES)
def test_example_flow_against_reference_golden(engine):
    """Golden vectors = the reference's own pzflow code (oracle/make_golden_shim.py)."""
    z = golden("ref_shim_example_flow.npz")
    ex = F.example_flow()
    plan = _plan(ex, engine)
    bi, li, p64 = _o64(ex)
    X, U = z["X"], z["U"]
    lp = _np(plan.log_prob(torch.from_numpy(X)))
    assert_noise_floor(f"golden log_prob [{engine}]", lp, z["log_prob"],
                       O.flow_log_prob(bi, li, p64, X.astype(np.float64)), ratio=2.0, max_class_mismatch=1,
                       below=LP_UNDERFLOW)
    assert zero_class(lp)[[2, 3]].all() and not zero_class(lp)[4]  # far out of bounds, NaN input
    u, ld = plan.forward(torch.from_numpy(X))
    u64, ld64 = O.apply_bijector(bi, p64[1], X.astype(np.float64), np.zeros((len(X), 1)))
    ok = ~np.isnan(X).any(axis=1)
    assert_noise_floor(f"golden forward u [{engine}]", _np(u)[ok], z["forward_u"][ok], u64[ok], ratio=2.0)
    assert_noise_floor(f"golden forward ld [{engine}]", _np(ld)[ok], z["forward_ld"][ok], ld64[ok], ratio=2.0)
    assert np.isnan(_np(u)[3]).any() and np.isnan(_np(ld)[3])
    x, ldi = plan.inverse(torch.from_numpy(U))
    x64, ldi64 = O.flow_inverse(bi, p64, U.astype(np.float64))
    assert_noise_floor(f"golden inverse x [{engine}]", _np(x), z["inverse_x"], x64, ratio=2.0, slack=2e-5)
    assert_noise_floor(f"golden inverse ld [{engine}]", _np(ldi), z["inverse_ld"], ldi64, ratio=2.0, slack=2e-5)
    # out-of-bounds latent rows pass through the splines untouched: only ShiftBounds^-1 acts
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    np.testing.assert_allclose(_np(x)[2], 7.5 * (maxs - mins) / 2 / 4 + (maxs + mins) / 2, rtol=1e-6)
    # posterior (normalised, batched == unbatched, un-normalised, another column)
    rest = np.ascontiguousarray(X[:24, 1:])
    grid = z["grid"]
    pdf = _np(plan.posterior(torch.from_numpy(rest), 0, torch.from_numpy(grid)))
    p64v = O.flow_posterior(bi, li, p64, rest.astype(np.float64), 0, grid.astype(np.float64))
    assert_noise_floor(f"golden posterior [{engine}]", pdf, z["posterior"], p64v, ratio=2.0, slack=1e-5,
                       max_class_mismatch=0)
    np.testing.assert_allclose(O.trapezoid(pdf[pdf.sum(1) > 0], grid), 1.0, rtol=2e-6)
    assert (pdf[3] == 0).all() and (z["posterior"][3] == 0).all()  # NaN input row -> all-zero pdf
    un = _np(plan.posterior(torch.from_numpy(rest), 0, torch.from_numpy(grid), normalize=False, nan_to_zero=False))
    un64 = O.flow_posterior(bi, li, p64, rest.astype(np.float64), 0, grid.astype(np.float64), normalize=False,
                            nan_to_zero=False)
    assert_noise_floor(f"golden posterior unnorm [{engine}]", un, z["posterior_unnorm"], un64, ratio=2.0,
                       slack=1e-3, max_class_mismatch=0)
    rest_r = np.ascontiguousarray(np.delete(X[:24], 3, axis=1))
    pdf_r = _np(plan.posterior(torch.from_numpy(rest_r), 3, torch.from_numpy(z["grid_r"])))
    p64r = O.flow_posterior(bi, li, p64, rest_r.astype(np.float64), 3, z["grid_r"].astype(np.float64))
    assert_noise_floor(f"golden posterior col r [{engine}]", pdf_r, z["posterior_r"], p64r, ratio=2.0, slack=1e-5,
                       max_class_mismatch=0)


@pytest.mark.parametrize("engine", ENGINES)
def test_notebook_samples_anchor(engine):
    """docs/tutorials/marginalization.ipynb:133-135 (real-JAX samples of the shipped flow)."""
    ex = F.example_flow()
    plan = _plan(ex, engine)
    X = golden("ref_shim_example_flow.npz")["nb_X"]
    u, _ = plan.forward(torch.from_numpy(X))
    assert (_np(u).__abs__() <= 5).all()
    np.testing.assert_allclose(_np(plan.log_prob(torch.from_numpy(X))), [5.4340, 6.7095], atol=2e-3)
    np.testing.assert_allclose(_np(plan.inverse(u)[0]), X, atol=2e-5)


# --------------------------------------------------------------------------- #
# the reference's bijector battery (tests/test_bijectors.py) through the mirror #
# --------------------------------------------------------------------------- #
def test_bijector_battery_golden_through_protocol():
    from pzflow_b200.utils import build_bijector_from_info
    z = golden("ref_shim_bijectors.npz")
    x = z["x"]
    for name in z["names"]:
        name = str(name)
        info = F.unpack_tree(f"{name}_info", z)
        params = F.unpack_tree(f"{name}_params", z)
        cond = z[f"{name}_cond"]
        init_fun, info2 = build_bijector_from_info(info)
        _, fwd, inv = init_fun(0, 7)
        y, ld = fwd(params, x, conditions=cond)
        assert y.shape == x.shape and ld.shape == (3,)
        assert_tolerance(f"battery {name} fwd", y, z[f"{name}_fwd"])
        assert_tolerance(f"battery {name} fwd ld", ld, z[f"{name}_fwd_ld"])
        xi, ldi = inv(params, y, conditions=cond)
        # the reference's own property (tests/test_bijectors.py:71-86), in float32 ulps of the values
        np.testing.assert_allclose(xi, x, atol=4e-6, err_msg=name)
        np.testing.assert_allclose(ld, -ldi, atol=8e-6, err_msg=name)
        xi2, ldi2 = inv(params, x, conditions=cond)
        assert_tolerance(f"battery {name} inv", xi2, z[f"{name}_inv"])
        assert_tolerance(f"battery {name} inv ld", ldi2, z[f"{name}_inv_ld"])


# --------------------------------------------------------------------------- #
# default 2-D flow (config 1) and the conditional ensemble (config 4) goldens   #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("engine", ENGINES)
def test_two_moons_default_flow_golden(engine):
    z = golden("ref_shim_default_flows.npz")
    cfg = dict(bijector_info=F.unpack_tree("c1_info", z), latent_info=("CentBeta13", (2, 5)),
               params=((), F.unpack_tree("c1_params", z)), input_dim=2)
    plan = _plan(cfg, engine)
    X = z["c1_X"]
    bi, li, p64 = _o64(cfg)
    lp = _np(plan.log_prob(torch.from_numpy(X)))
    assert_noise_floor(f"C1 log_prob [{engine}]", lp, z["c1_log_prob"],
                       O.flow_log_prob(bi, li, p64, X.astype(np.float64)), ratio=2.0, slack=5e-6,
                       max_class_mismatch=0)
    assert within_tol(lp, z["c1_log_prob"]).mean() > 0.98
    g = z["c1_grid"]
    px = _np(plan.posterior(torch.from_numpy(np.ascontiguousarray(X[:32, 1:])), 0, torch.from_numpy(g)))
    assert_tolerance(f"C1 posterior x [{engine}]", px, z["c1_posterior_x"], min_frac=0.99)
    py = _np(plan.posterior(torch.from_numpy(np.ascontiguousarray(X[:32, :1])), 1, torch.from_numpy(g),
                            normalize=False))
    assert_tolerance(f"C1 posterior y [{engine}]", py, z["c1_posterior_y"], min_frac=0.99)
    xs = _np(plan.inverse(torch.from_numpy(z["c1_sample_u"]))[0])
    assert_tolerance(f"C1 sample map [{engine}]", xs, z["c1_sample_x"], min_frac=0.995)


def test_flow_and_ensemble_api_against_golden():
    """Flow / FlowEnsemble mirrors called the way a pzflow user calls them (DataFrames in)."""
    from pzflow_b200 import Flow, FlowEnsemble
    from pzflow_b200 import bijectors as B
    from pzflow_b200.utils import build_bijector_from_info
    z = golden("ref_shim_default_flows.npz")
    cols = [str(c) for c in z["ens_columns"]]
    df = pd.DataFrame(z["ens_df"], columns=cols)
    ens = FlowEnsemble(data_columns=("a", "b", "c"), conditional_columns=("p", "q"),
                       bijector=B.RollingSplineCoupling(3, n_conditions=2), N=3)
    for i, fl in enumerate(ens._ensemble.values()):
        fl.set_bijector(build_bijector_from_info(F.unpack_tree(f"ens{i}_info", z)),
                        params=F.unpack_tree(f"ens{i}_params", z))
        fl._condition_means, fl._condition_stds = z["ens_cond_means"], z["ens_cond_stds"]
    f0 = ens._ensemble["Flow 0"]
    lp0 = f0.log_prob(df)
    assert lp0.shape == (200,)
    assert_tolerance("API flow0 log_prob", lp0, z["ens_flow0_log_prob"], min_frac=0.97)
    allp = ens.log_prob(df, returnEnsemble=True)
    assert allp.shape == (200, 3)
    np.testing.assert_array_equal(allp[:, 0], lp0)  # tests/test_ensemble.py:20-27
    assert_tolerance("API ensemble log_prob all", allp, z["ens_log_prob_all"], min_frac=0.97)
    mean = ens.log_prob(df)
    np.testing.assert_allclose(mean, np.log(np.exp(allp).mean(axis=1)), rtol=2e-6, atol=2e-6)  # test_ensemble.py:29-35
    assert_tolerance("API ensemble log_prob", mean, z["ens_log_prob"], min_frac=0.97)
    g = z["ens_grid"]
    post = ens.posterior(df.iloc[:20], column="a", grid=g)
    assert post.shape == (20, 25)
    assert_tolerance("API ensemble posterior", post, z["ens_posterior"], min_frac=0.98)
    post_all = ens.posterior(df.iloc[:20], column="a", grid=g, returnEnsemble=True)
    assert post_all.shape == (20, 3, 25)
    assert_tolerance("API ensemble posterior all", post_all, z["ens_posterior_all"], min_frac=0.98)
    pb = f0.posterior(df.iloc[:20], column="b", grid=g)
    assert_tolerance("API flow0 posterior b", pb, z["ens_flow0_posterior_b"], min_frac=0.98)
    pb_batched = f0.posterior(df.iloc[:20], column="b", grid=g, batch_size=7)  # tests/test_flow.py:275-287
    np.testing.assert_array_equal(pb, pb_batched)
    # conditional sampling: the map u -> x (flow.py:787-795)
    xs = f0._inverse(f0._params[1], z["ens_sample_u"], conditions=z["ens_sample_cond"])[0]
    assert_tolerance("API conditional sample map", xs, z["ens_sample_x"], min_frac=0.99)
    s = f0.sample(2, conditions=df.iloc[:10], seed=0)
    assert s.shape == (20, 5) and list(s.columns) == cols
    np.testing.assert_allclose(s[["p", "q"]].to_numpy(), np.repeat(df.iloc[:10][["p", "q"]].to_numpy(), 2, 0),
                               rtol=1e-5)
    s2 = f0.sample(2, conditions=df.iloc[:10], seed=0)
    pd.testing.assert_frame_equal(s, s2)  # seed control, tests/test_flow.py:337-342
    assert ens.sample(10, conditions=df.iloc[:4], seed=0).shape == (40, 5)
    assert ens.sample(3, conditions=df.iloc[:2], seed=0, returnEnsemble=True).shape == (18, 5)


def test_example_flow_through_flow_api(tmp_path):
    """tests/test_examples.py:31-40 on the mirror: load, sample, log_prob, posterior."""
    from pzflow_b200 import Flow
    ex = F.example_flow()
    d = {"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": None,
         "condition_means": None, "condition_stds": None, "data_error_model": None,
         "condition_error_model": None, "autoscale_conditions": True, "info": "example",
         "latent_info": ex["latent_info"], "bijector_info": ex["bijector_info"], "params": ex["params"]}
    flow = Flow(_dictionary=d)
    samples = flow.sample(200, seed=0)
    assert samples.shape == (200, 7) and list(samples.columns) == list(F.GALAXY_COLUMNS)
    pd.testing.assert_frame_equal(samples, flow.sample(200, seed=0))
    assert not samples.equals(flow.sample(200, seed=1))
    lp = flow.log_prob(samples)
    assert lp.shape == (200,) and np.isfinite(lp).all() and (lp > -50).all()
    grid = np.arange(0, 2.5, 0.5)
    pdfs = flow.posterior(samples, column="redshift", grid=grid)
    assert pdfs.shape == (200, 5)
    X = samples.to_numpy().astype(np.float32)
    o32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X)
    o64 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], O.cast_params(ex["params"], np.float64),
                          X.astype(np.float64))
    assert_noise_floor("API example log_prob", lp, o32, o64, ratio=2.5, max_class_mismatch=0, below=LP_UNDERFLOW)
    # marginalisation rules run through the same fused posterior (flow.py:560-664)
    flagged = samples.iloc[:6].copy()
    flagged.iloc[1, 1] = 99.0
    flagged.iloc[4, 1] = 99.0
    rules = {"flag": 99, "u": lambda row: np.linspace(26, 31, 9)}
    pm = flow.posterior(flagged, column="redshift", grid=grid, marg_rules=rules)
    assert pm.shape == (6, 5) and np.isfinite(pm).all()
    np.testing.assert_allclose(pm[[0, 2, 3, 5]], pdfs[[0, 2, 3, 5]], rtol=1e-5, atol=1e-7)
    # a flagged row: un-normalised pdfs summed over its marginalisation grid, then normalised (flow.py:617-664, 732-737)
    for r in (1, 4):
        expanded = pd.DataFrame(np.repeat(flagged.iloc[[r]].to_numpy(), 9, axis=0), columns=flagged.columns)
        expanded["u"] = np.linspace(26, 31, 9)
        summed = flow.posterior(expanded, column="redshift", grid=grid, normalize=False).astype(np.float64).sum(axis=0)
        want_row = summed / np.trapezoid(summed, np.asarray(grid, np.float64))
        np.testing.assert_allclose(pm[r], want_row, rtol=1e-4, atol=1e-7)


# --------------------------------------------------------------------------- #
# synthetic-parameter configs: conditional (4), 32-D (5), periodic / odd shapes #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("engine", ENGINES)
def test_conditional_flow_c4(engine):
    cfg = F.config_c4(1)[0]
    plan = _plan(cfg, engine)
    rng = np.random.default_rng(0)
    ex = F.example_flow()
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    X = rng.uniform(mins - 0.05, maxs + 0.05, size=(8192, 7)).astype(np.float32)
    cond = rng.normal(size=(8192, 2)).astype(np.float32)
    bi, li, p64 = _o64(cfg)
    ob = []
    o32 = O.flow_log_prob(cfg["bijector_info"], cfg["latent_info"], cfg["params"], X, cond, bins=ob)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64), cond.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), torch.from_numpy(cond), return_bins=True)
    assert_noise_floor(f"C4 log_prob [{engine}]", _np(lp), o32, o64, ratio=2.0, slack=5e-6, below=LP_UNDERFLOW)
    assert within_tol(_np(lp), o32)[~zero_class(o32)].mean() > 0.99
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    print(f"BINS C4 [{engine}] flipped {flips} of {total}")
    assert flips <= max(3, int(2e-4 * total))
    with pytest.raises(ValueError):
        plan.log_prob(torch.from_numpy(X))  # conditions are required (flow.py:777-780 analogue)


@pytest.mark.parametrize("engine", ["simt", "wide"])
def test_high_dim_c5(engine):
    """BASELINE config 5 (32-D, K = 32, hidden 256, 32 coupling stages): the SIMT engine and the wide tensor-core
    engine (weights streamed in slabs, 96-logit spline epilogue) against the oracle, bin indices included."""
    cfg = F.config_c5()
    plan = _plan(cfg, engine)
    from pzflow_b200.plan import Plan
    assert Plan(cfg["bijector_info"], cfg["params"][1], 32, cfg["latent_info"]).engine == "wide"   # what AUTO picks
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, size=(1536, 32)).astype(np.float32)
    X[:8] *= 1.3  # a few rows outside the ShiftBounds box -> out-of-bounds identity paths
    bi, li, p64 = _o64(cfg)
    o32 = O.flow_log_prob(cfg["bijector_info"], cfg["latent_info"], cfg["params"], X)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64))
    lp = _np(plan.log_prob(torch.from_numpy(X)))
    assert_noise_floor("C5 log_prob", lp, o32, o64, ratio=2.0, slack=1e-5, below=-1e4)
    u32, ld32 = O.apply_bijector(cfg["bijector_info"], cfg["params"][1], X, np.zeros((len(X), 1), np.float32))
    u64, _ = O.apply_bijector(bi, p64[1], X.astype(np.float64), np.zeros((len(X), 1)))
    u, ld = plan.forward(torch.from_numpy(X))
    assert_noise_floor("C5 forward u", _np(u), u32, u64, ratio=2.0, slack=2e-6)
    x32, _ = O.flow_inverse(cfg["bijector_info"], cfg["params"], u32)
    x64, _ = O.flow_inverse(bi, p64, u32.astype(np.float64))
    xi = _np(plan.inverse(torch.from_numpy(u32))[0])
    assert_noise_floor("C5 inverse x", xi, x32, x64, ratio=2.0, slack=2e-6)
    ob = []
    O.flow_log_prob(cfg["bijector_info"], cfg["latent_info"], cfg["params"], X, bins=ob)
    _, bins = plan.log_prob(torch.from_numpy(X), return_bins=True)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    print(f"BINS C5 [{engine}] flipped {flips} of {total} ({rows} rows)")
    assert flips <= max(3, int(5e-4 * total))
    if engine == "wide":
        assert plan.overflow_rows() == 0


WIDE_SHAPES = [
    # (info, D, C): shapes only the wide tensor-core engine (or SIMT) takes
    (F.rolling_info(3, 1, 32, 5, 2, 256), 6, 0),                       # K = 32, hidden 256, L = 3
    (F.rolling_info(2, 1, 20, 5, 1, 100, None, 2), 5, 2),              # K = 20 padded to 32, ONE hidden layer, H = 100
    (F.rolling_info(2, 2, 8, 4, 3, 200, None, 1), 7, 1),               # K = 8 padded to 16, THREE hidden layers, H = 200
    (F.rolling_info(2, 1, 16, 5, 4, 64), 4, 0),                        # four hidden layers of 64
    (("Chain", (("StandardScaler", (np.linspace(-1, 1, 40).astype(np.float32), np.linspace(1, 2, 40).astype(np.float32))),
                F.rolling_info(2, 3, 16, 5, 2, 128, None, 3))), 40, 3),   # 20 + 3 conditioner inputs (K1 = 32), L = 20
    (F.rolling_info(2, 5, 24, 5, 2, 160, 9, 0), 64, 0),                # 55 conditioner inputs (K1 = 64), transformed_dim 9
    (F.rolling_info(2, 1, 32, 3, 2, 256, None, 0, True), 4, 0),        # periodic, K = 32
    (("Chain", (F.rolling_info(1, 1, 32, 5, 2, 192, None, 3), )), 1, 3),   # D = 1: U = 0, key = conditions only
]


@pytest.mark.parametrize("info,D,C", WIDE_SHAPES)
def test_wide_engine_shapes(info, D, C):
    """The wide tensor-core engine on shapes outside the default family, against the oracle and the SIMT engine:
    log_prob, forward, inverse, bin indices; 777 rows = 7 tiles, the last one ragged."""
    rng = np.random.default_rng(D + 7)
    params = O.synth_bijector_params(info, D, rng, out_scale=3.0)
    p64 = O.cast_params(params, np.float64)
    li = ("CentBeta13", (D, 5))
    cfg = dict(bijector_info=info, latent_info=li, params=((), params), input_dim=D)
    plan = _plan(cfg, "wide")
    assert not plan.tc_supported
    n = 777
    X = rng.uniform(-6, 6, size=(n, D)).astype(np.float32)
    if D > 8:   # many columns: keep most rows inside the latent's support (one stray column already means lp = -inf)
        X = rng.uniform(-4.9, 4.9, size=(n, D)).astype(np.float32)
        stray = rng.random(size=X.shape) < 0.01
        X[stray] = rng.uniform(-6, 6, size=int(stray.sum())).astype(np.float32)
    cond = rng.normal(size=(n, max(C, 1))).astype(np.float32) if C else np.zeros((n, 1), np.float32)
    ct = torch.from_numpy(cond) if C else None
    # The fp16 hi/lo split keeps 22 of float32's 24 significand bits, so a dense layer is up to ~4x noisier than an
    # fp32 FFMA layer.  On the trained flow that drowns in the splines' conditioning (ratio 1); on these deep,
    # well-conditioned synthetic conditioners it is what is left, hence ratio 6 -- with north_star's own tolerance
    # (1e-4 + 1e-5 |b|) asserted beside it.
    kw = dict(ratio=6.0, slack=1e-5, max_class_mismatch=2)
    ob = []
    o32 = O.flow_log_prob(info, li, ((), params), X, cond, bins=ob)
    o64 = O.flow_log_prob(info, li, ((), p64), X.astype(np.float64), cond.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), ct, return_bins=True)
    lp = _np(lp)
    below = LP_UNDERFLOW if D <= 8 else -1e4   # many columns: a typical log-density is a few hundred below zero
    assert_noise_floor(f"wide D={D} log_prob", lp, o32, o64, below=below, **kw)
    assert within_tol(lp, o32)[~zero_class(o32, below) & ~zero_class(lp, below)].mean() > (0.97 if D <= 8 else 0.9)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    print(f"BINS wide D={D} flipped {flips} of {total} ({rows} rows)")
    assert flips <= max(3, int(5e-4 * total))
    u32, ld32 = O.apply_bijector(info, params, X, cond)
    u64, ld64 = O.apply_bijector(info, p64, X.astype(np.float64), cond.astype(np.float64))
    u, ld = plan.forward(torch.from_numpy(X), ct)
    assert_noise_floor("wide forward u", _np(u), u32, u64, **kw)
    assert_noise_floor("wide forward ld", _np(ld), ld32, ld64, **kw)
    x32, ldi32 = O.apply_bijector(info, params, u32, cond, inverse=True)
    x64, ldi64 = O.apply_bijector(info, p64, u32.astype(np.float64), cond.astype(np.float64), inverse=True)
    xi, ldi = plan.inverse(torch.from_numpy(u32), ct)
    kw["max_class_mismatch"] = max(2, n // 100)
    assert_noise_floor("wide inverse x", _np(xi), x32, x64, **kw)
    assert_noise_floor("wide inverse ld", _np(ldi), ldi32, ldi64, **kw)
    # the SIMT engine on the same plan: the two engines agree at the noise floor too
    plan.set_engine("simt")
    lp_s = _np(plan.log_prob(torch.from_numpy(X), ct))
    assert_noise_floor("wide vs simt", lp, lp_s, o64, below=below, **kw)


def test_wide_engine_on_the_shipped_flow():
    """The wide engine also takes the default family (K = 16 / hidden 128 runs as two-dimension passes): the shipped
    7-D flow through it against the oracle, the golden posterior and the default tensor-core engine."""
    ex = F.example_flow()
    plan = _plan(ex, "wide")
    X = F.galaxy_rows(20000, seed=0)
    ob = []
    o32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X, bins=ob)
    bi, li, p64 = _o64(ex)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), return_bins=True)
    assert_noise_floor("C2 log_prob [wide]", _np(lp), o32, o64, below=LP_UNDERFLOW)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    assert flips <= max(3, int(2e-4 * total))
    z = golden("ref_shim_example_flow.npz")
    rest = np.ascontiguousarray(z["X"][:24, 1:])
    pdf = _np(plan.posterior(torch.from_numpy(rest), 0, torch.from_numpy(z["grid"])))
    p64v = O.flow_posterior(bi, li, p64, rest.astype(np.float64), 0, z["grid"].astype(np.float64))
    assert_noise_floor("golden posterior [wide]", pdf, z["posterior"], p64v, ratio=2.0, slack=1e-5, max_class_mismatch=0)
    u = torch.from_numpy(np.random.default_rng(1).uniform(-5, 5, size=(5000, 7)).astype(np.float32))
    x_w = plan.inverse(u)[0]
    plan.set_engine("tc")
    x_t = plan.inverse(u)[0]
    assert float((x_w - x_t).abs().median()) < 1e-5


def test_fp16_overflow_rows_are_counted():
    """DESIGN.md 'Known deviation': a pre-activation beyond +-65504 overflows the fp16 hi part on the tensor-core
    engines; such rows are counted (pzb_plan_overflow_rows) instead of passing silently."""
    info = F.rolling_info(2, 1, 32, 5, 2, 256)
    rng = np.random.default_rng(3)
    params = O.synth_bijector_params(info, 4, rng)
    cfg = dict(bijector_info=info, latent_info=("Uniform", (4, 5)), params=((), params), input_dim=4)
    plan = _plan(cfg, "wide")
    X = rng.uniform(-4, 4, size=(300, 4)).astype(np.float32)
    plan.log_prob(torch.from_numpy(X))
    assert plan.overflow_rows(reset=True) == 0
    X[5, :2] = 3.0e6     # conditioner inputs ~1e6 x outside the box: |W1 x| > 65504
    X[200, 0] = -2.0e7
    X[100, 1] = np.nan   # a NaN input is NOT an overflow: the reference returns NaN for it too
    lp = _np(plan.log_prob(torch.from_numpy(X).cuda()))
    assert plan.overflow_rows() == 2
    assert zero_class(lp)[[5, 100, 200]].all() and not zero_class(lp, LP_UNDERFLOW)[[4, 6, 199, 201]].any()
    plan.set_engine("simt")   # fp32 throughout: finite logits, the rows are merely out of bounds
    lp_simt = _np(plan.log_prob(torch.from_numpy(X).cuda()))
    # an AUTO plan's host-buffer calls notice the overflow and re-evaluate the pipeline chunks that held such rows on the
    # SIMT engine: the reference's answer for them, the tensor-core result everywhere else
    plan.set_engine("auto")
    assert plan.engine == "wide"
    plan.overflow_rows(reset=True)
    lp_host = plan.log_prob_host(X).numpy()
    is_simt = lp_host.view(np.int32) == lp_simt.view(np.int32)
    assert is_simt[[5, 200]].all() and (is_simt | (lp_host.view(np.int32) == lp.view(np.int32))).all()
    Xok = np.delete(X, [5, 200], axis=0)
    assert np.array_equal(plan.log_prob_host(Xok).numpy().view(np.int32),
                          _np(plan.log_prob(torch.from_numpy(Xok).cuda())).view(np.int32))   # no overflow: no second pass
    # the guard works per pipeline chunk: bad rows in one chunk of a long call send THAT chunk through the SIMT engine,
    # every other row keeps the tensor-core result bit for bit (-9999 style catalogue placeholders are enough to overflow)
    big = rng.uniform(-4, 4, size=(3_000_000, 4)).astype(np.float32)
    bad = [2_500_000, 2_500_017]
    big[bad, 0] = -5.0e6
    dev = torch.from_numpy(big).cuda()
    tc = _np(plan.log_prob(dev))
    plan.set_engine("simt")
    simt = _np(plan.log_prob(dev))
    plan.set_engine("auto")
    plan.overflow_rows(reset=True)
    host = plan.log_prob_host(big).numpy()
    same_tc = host.view(np.int32) == tc.view(np.int32)
    same_simt = host.view(np.int32) == simt.view(np.int32)
    assert same_simt[bad].all() and zero_class(tc)[bad].all()
    assert (same_tc | same_simt).all()
    redone = np.nonzero(~same_tc)[0]                     # rows that differ from the tensor-core result: one chunk's worth
    assert redone.size and redone.min() > 1_000_000 and redone.max() - redone.min() < 1_400_000
    assert same_tc[:1_000_000].all()
    # the same through the DataFrame column path of Flow.log_prob (sliced column jobs) -- values only
    cols = [np.ascontiguousarray(big[:, j].astype(np.float64)) for j in range(4)]
    hc = plan.log_prob_host_columns(cols).numpy()
    assert np.array_equal(hc.view(np.int32), host.view(np.int32))


@pytest.mark.parametrize("info,D,C", [
    (("Chain", (("StandardScaler", (np.linspace(-1, 1, 7).astype(np.float32), np.linspace(1, 8, 7).astype(np.float32))),
                F.rolling_info(2, 1, 16, 3, 2, 128, None, 0, True),
                ("NeuralSplineCoupling", (16, 3, 2, 128, 3, 3, False)))), 7, 3),   # periodic + transformed_dim + conditions
    (("Chain", (("ShiftBounds", (-2.0, 2.0, 3.0)), F.rolling_info(3, 2, 5, 4, 1, 40))), 5, 0),  # odd K, 1 hidden layer, H=40
    (F.rolling_info(2, 1, 8, 5, 3, 200, 1, 2), 4, 2),   # 3 hidden layers, H=200 (2 column chunks), L=1
    (("NeuralSplineCoupling", (4, 5, 0, 16, None, 0, False)), 3, 0),  # no hidden layer at all
    (("Chain", (F.rolling_info(1, 1, 16, 5, 2, 128, None, 3), )), 1, 3),  # D=1: U=0, key = conditions only
])
def test_odd_shapes_simt(info, D, C):
    rng = np.random.default_rng(2)
    params = O.synth_bijector_params(info, D, rng, out_scale=3.0)
    p64 = O.cast_params(params, np.float64)
    li = ("CentBeta13", (D, 5))
    cfg = dict(bijector_info=info, latent_info=li, params=((), params), input_dim=D)
    plan = _plan(cfg, "simt")
    n = 777
    X = rng.uniform(-6, 6, size=(n, D)).astype(np.float32)
    cond = rng.normal(size=(n, max(C, 1))).astype(np.float32) if C else np.zeros((n, 1), np.float32)
    ct = torch.from_numpy(cond) if C else None
    kw = dict(ratio=2.0, slack=1e-5, max_class_mismatch=2)
    o32 = O.flow_log_prob(info, li, ((), params), X, cond)
    o64 = O.flow_log_prob(info, li, ((), p64), X.astype(np.float64), cond.astype(np.float64))
    lp = _np(plan.log_prob(torch.from_numpy(X), ct))
    assert_noise_floor(f"odd {info[0]} D={D} log_prob", lp, o32, o64, below=LP_UNDERFLOW, **kw)
    assert within_tol(lp, o32)[~zero_class(o32, LP_UNDERFLOW) & ~zero_class(lp, LP_UNDERFLOW)].mean() > 0.97
    u32, ld32 = O.apply_bijector(info, params, X, cond)
    u64, ld64 = O.apply_bijector(info, p64, X.astype(np.float64), cond.astype(np.float64))
    u, ld = plan.forward(torch.from_numpy(X), ct)
    assert_noise_floor("odd forward u", _np(u), u32, u64, **kw)
    assert_noise_floor("odd forward ld", _np(ld), ld32, ld64, **kw)
    x32, ldi32 = O.apply_bijector(info, params, u32, cond, inverse=True)
    x64, ldi64 = O.apply_bijector(info, p64, u32.astype(np.float64), cond.astype(np.float64), inverse=True)
    xi, ldi = plan.inverse(torch.from_numpy(u32), ct)
    # the inverse of an out_scale=3 spline is ill-posed where the forward map is flat: a handful of
    # elements end as NaN (negative discriminant) in one float32 evaluation and not in the other
    kw["max_class_mismatch"] = max(2, n // 100)
    assert_noise_floor("odd inverse x", _np(xi), x32, x64, **kw)
    assert_noise_floor("odd inverse ld", _np(ldi), ldi32, ldi64, **kw)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("D,C", [(9, 0), (12, 1), (16, 0), (20, 3), (1, 12), (7, 9)])
def test_wide_default_flows_run_as_sub_stages(D, C, engine):
    """Default flows (pzflow/flow.py:314-322: K = 16, hidden 128) with more than eight columns: a coupling stage then
    has more than four spline dimensions and the tensor-core engine takes them four at a time (sub-stages); and with
    more than eight conditioner inputs (columns + conditions), which ride above the bias row of the K = 16 key
    operand.  Both engines against the oracle, bin indices included."""
    cfg = F.default_flow(D, -np.ones(D), np.ones(D), n_conditions=C, seed=D)
    plan = _plan(cfg, engine)
    assert plan.tc_supported
    from pzflow_b200.plan import Plan
    assert Plan(cfg["bijector_info"], cfg["params"][1], D, cfg["latent_info"]).engine == "tc"   # what AUTO picks
    rng = np.random.default_rng(D)
    n = 3000                                             # 24 tiles: ragged last chunk, both epilogue groups busy
    X = rng.uniform(-1.02, 1.02, size=(n, D)).astype(np.float32)
    cond = rng.normal(size=(n, max(C, 1))).astype(np.float32) if C else np.zeros((n, 1), np.float32)
    ct = torch.from_numpy(cond) if C else None
    bi, li, p64 = _o64(cfg)
    ob = []
    o32 = O.flow_log_prob(bi, li, cfg["params"], X, cond, bins=ob)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64), cond.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), ct, return_bins=True)
    assert_noise_floor(f"wide D={D} log_prob [{engine}]", _np(lp), o32, o64, ratio=2.0, slack=1e-5, below=LP_UNDERFLOW)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    print(f"BINS wide D={D} [{engine}] flipped {flips} of {total}")
    assert flips <= max(3, int(2e-4 * total))
    u32, ld32 = O.apply_bijector(bi, cfg["params"][1], X, cond)
    u64, ld64 = O.apply_bijector(bi, p64[1], X.astype(np.float64), cond.astype(np.float64))
    u, ld = plan.forward(torch.from_numpy(X), ct)
    assert_noise_floor(f"wide D={D} forward u", _np(u), u32, u64, ratio=2.0, slack=2e-6)
    assert_noise_floor(f"wide D={D} forward ld", _np(ld), ld32, ld64, ratio=2.0, slack=1e-5)
    x32, _ = O.apply_bijector(bi, cfg["params"][1], u32, cond, inverse=True)
    x64, _ = O.apply_bijector(bi, p64[1], u32.astype(np.float64), cond.astype(np.float64), inverse=True)
    xi, _ = plan.inverse(torch.from_numpy(u32), ct)
    assert_noise_floor(f"wide D={D} inverse x", _np(xi), x32, x64, ratio=2.0, slack=2e-6, max_class_mismatch=3)


@pytest.mark.parametrize("engine", ENGINES)
def test_stages_with_different_spline_dim_counts(engine):
    """Coupling stages with 1, 3 and 4 transformed dimensions in one chain: the tensor-core engine's per-stage
    bookkeeping (which half of a row's threads writes the next key operand) changes from stage to stage."""
    info = ("Chain", (("ShiftBounds", (-np.ones(6, np.float32), np.ones(6, np.float32), 4.0)),
                      F.rolling_info(2, 1, 16, 5, 2, 128, 1, 0), F.rolling_info(3, 1, 16, 5, 2, 128, 4, 0),
                      F.rolling_info(2, 2, 16, 5, 2, 128, 3, 0), F.rolling_info(2, 1, 16, 5, 2, 128, 2, 0)))
    rng = np.random.default_rng(11)
    params = O.synth_bijector_params(info, 6, rng)
    cfg = dict(bijector_info=info, latent_info=("CentBeta13", (6, 5)), params=((), params), input_dim=6)
    plan = _plan(cfg, engine)
    n = 5000
    X = rng.uniform(-1.02, 1.02, size=(n, 6)).astype(np.float32)
    bi, li, p64 = _o64(cfg)
    ob = []
    o32 = O.flow_log_prob(bi, li, cfg["params"], X, bins=ob)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), return_bins=True)
    assert_noise_floor(f"mixed L log_prob [{engine}]", _np(lp), o32, o64, ratio=2.0, slack=1e-5, below=LP_UNDERFLOW)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    assert flips <= max(3, int(2e-4 * total))
    u32, _ = O.apply_bijector(bi, cfg["params"][1], X, np.zeros((n, 1), np.float32))
    x32, _ = O.apply_bijector(bi, cfg["params"][1], u32, np.zeros((n, 1), np.float32), inverse=True)
    x64, _ = O.apply_bijector(bi, p64[1], u32.astype(np.float64), np.zeros((n, 1)), inverse=True)
    xi, _ = plan.inverse(torch.from_numpy(u32))
    assert_noise_floor(f"mixed L inverse [{engine}]", _np(xi), x32, x64, ratio=2.0, slack=2e-6, max_class_mismatch=3)


@pytest.mark.parametrize("K,H,D,C", [(8, 64, 5, 0), (3, 100, 4, 2), (15, 128, 7, 0), (2, 16, 2, 0), (12, 40, 10, 1)])
def test_narrow_conditioners_are_padded_into_the_tensor_core_shape(K, H, D, C):
    """hidden_dim < 128 and K < 16 on the tensor-core engine: the packer pads hidden units with zero weights and the
    spline with zero-width trailing bins (csrc/nsf_tc.cu: tc_pack_stage); results against the oracle and against the
    SIMT engine, bin indices included."""
    info = ("Chain", (("ShiftBounds", (-np.ones(D, np.float32), np.ones(D, np.float32), 4.0)),
                      F.rolling_info(D, 1, K, 5, 2, H, None, C)))
    rng = np.random.default_rng(100 * K + H)
    params = O.synth_bijector_params(info, D, rng)
    cfg = dict(bijector_info=info, latent_info=("CentBeta13", (D, 5)), params=((), params), input_dim=D)
    plan = _plan(cfg, "tc")
    simt = _plan(cfg, "simt")
    n = 4000
    X = rng.uniform(-1.02, 1.02, size=(n, D)).astype(np.float32)
    X[:4] = 3.0                                          # |x| > 2B after ShiftBounds: beyond the last knot, idx = K
    X[4:8] = -3.0
    cond = rng.normal(size=(n, max(C, 1))).astype(np.float32) if C else np.zeros((n, 1), np.float32)
    ct = torch.from_numpy(cond) if C else None
    bi, li, p64 = _o64(cfg)
    ob = []
    o32 = O.flow_log_prob(bi, li, cfg["params"], X, cond, bins=ob)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64), cond.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), ct, return_bins=True)
    assert (np.concatenate(ob, axis=1)[:8] == K).any() and _np(bins)[:8].max() == K
    assert_noise_floor(f"narrow K={K} H={H} log_prob", _np(lp), o32, o64, ratio=2.0, slack=1e-5, below=LP_UNDERFLOW)
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    assert flips <= max(3, int(2e-4 * total))
    assert int(_np(bins).max()) <= K - 1 + 1            # K = out of range marker at most; never a padding bin
    lp_simt = _np(simt.log_prob(torch.from_numpy(X), ct))
    ok = ~zero_class(o32, LP_UNDERFLOW)
    assert within_tol(_np(lp)[ok], lp_simt[ok]).mean() > 0.99
    u32, _ = O.apply_bijector(bi, cfg["params"][1], X, cond)
    x32, _ = O.apply_bijector(bi, cfg["params"][1], u32, cond, inverse=True)
    x64, _ = O.apply_bijector(bi, p64[1], u32.astype(np.float64), cond.astype(np.float64), inverse=True)
    xi, _ = plan.inverse(torch.from_numpy(u32), ct)
    assert_noise_floor(f"narrow K={K} H={H} inverse", _np(xi), x32, x64, ratio=2.0, slack=2e-6, max_class_mismatch=3)


@pytest.mark.parametrize("seed", range(10))
def test_random_chains_tensor_core_engine_against_oracle_and_simt(seed):
    """Seeded random chains inside the tensor-core engine's envelope (2 hidden layers, hidden <= 128, K <= 16,
    <= 15 conditioner inputs, <= 31 columns): random column counts, transformed_dim, conditions, periodic stages,
    roll shifts and an affine step between coupling blocks.  log_prob and bins against the oracle, forward / inverse
    against the SIMT engine."""
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(1, 13))
    C = int(rng.integers(0, 4))
    blocks = [("ShiftBounds", (-np.ones(D, np.float32), np.ones(D, np.float32), 4.0))]
    for _ in range(int(rng.integers(1, 4))):
        periodic = bool(rng.random() < 0.25)
        K = 16 if periodic else int(rng.choice([2, 5, 8, 16, 16]))
        H = int(rng.choice([16, 64, 100, 128, 128]))
        td = None if (D == 1 or rng.random() < 0.5) else int(rng.integers(1, D))
        if td is not None and (D - td) + C > 15:
            td = None
        blocks.append(F.rolling_info(int(rng.integers(1, 4)), int(rng.integers(1, 3)), K, 5, 2, H, td, C, periodic))
        if rng.random() < 0.4:
            blocks.append(("StandardScaler", (rng.normal(0, 0.05, D).astype(np.float32),
                                              rng.uniform(0.9, 1.1, D).astype(np.float32))))
    info = ("Chain", tuple(blocks))
    params = O.synth_bijector_params(info, D, rng)
    latent = ("Uniform", (D, 5)) if seed % 2 else ("CentBeta13", (D, 5))
    cfg = dict(bijector_info=info, latent_info=latent, params=((), params), input_dim=D)
    plan, simt = _plan(cfg, "tc"), _plan(cfg, "simt")
    n = int(rng.integers(200, 3000))
    X = rng.uniform(-1.05, 1.05, size=(n, D)).astype(np.float32)
    cond = rng.normal(size=(n, max(C, 1))).astype(np.float32) if C else np.zeros((n, 1), np.float32)
    ct = torch.from_numpy(cond) if C else None
    bi, li, p64 = _o64(cfg)
    ob = []
    o32 = O.flow_log_prob(bi, li, cfg["params"], X, cond, bins=ob)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64), cond.astype(np.float64))
    lp, bins = plan.log_prob(torch.from_numpy(X), ct, return_bins=True)
    assert_noise_floor(f"random chain {seed} log_prob", _np(lp), o32, o64, ratio=2.0, slack=1e-5, below=LP_UNDERFLOW,
                       max_class_mismatch=max(2, n // 100))
    flips, rows, total = count_bin_flips(_np(bins), np.concatenate(ob, axis=1))
    assert flips <= max(3, int(5e-4 * total)), (flips, total)
    u_tc, ld_tc = plan.forward(torch.from_numpy(X), ct)
    u_s, ld_s = simt.forward(torch.from_numpy(X), ct)
    ok = np.isfinite(_np(u_s)).all(axis=1) & np.isfinite(_np(u_tc)).all(axis=1)
    assert ok.mean() > 0.97
    assert within_tol(_np(u_tc)[ok], _np(u_s)[ok]).mean() > 0.99
    assert within_tol(_np(ld_tc)[ok], _np(ld_s)[ok]).mean() > 0.98
    x_tc, _ = plan.inverse(u_s, ct)
    x_s, _ = simt.inverse(u_s, ct)
    ok = np.isfinite(_np(x_s)).all(axis=1) & np.isfinite(_np(x_tc)).all(axis=1)
    assert ok.mean() > 0.97
    assert within_tol(_np(x_tc)[ok], _np(x_s)[ok]).mean() > 0.99


def test_empty_and_ragged_inputs():
    ex = F.example_flow()
    plan = _plan(ex, "simt")
    assert plan.log_prob(torch.empty((0, 7))).shape == (0,)
    assert plan.posterior(torch.empty((0, 6)), 0, torch.linspace(0, 2, 5)).shape == (0, 5)
    X = F.galaxy_rows(130, seed=4, oob_frac=0)
    full = _np(plan.log_prob(torch.from_numpy(X)))
    for n in (1, 63, 64, 65, 129):  # row counts around the 64-row tile
        np.testing.assert_array_equal(_np(plan.log_prob(torch.from_numpy(X[:n]))), full[:n])
    with pytest.raises(ValueError):
        plan.log_prob(torch.zeros((4, 6)))


def test_ragged_inputs_tc_engine():
    """Row counts around the tensor-core engine's tile (128), tile-pair (256) and chunk (1024 rows, 148 CTAs)
    boundaries: every prefix of a batch evaluates to the prefix of the batch's result, both directions."""
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    n = 148 * 1024 + 300
    X = torch.from_numpy(F.galaxy_rows(8192, seed=4, oob_frac=0)).cuda().repeat(n // 8192 + 1, 1)[:n].contiguous()
    full = plan.log_prob(X)
    u_full, ld_full = plan.forward(X)
    for m in (1, 2, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 148 * 1024 - 1, 148 * 1024 + 1):
        assert _same_bits(plan.log_prob(X[:m]), full[:m]), m
        u, ld = plan.forward(X[:m])
        assert _same_bits(u, u_full[:m]) and _same_bits(ld, ld_full[:m]), m
    xr, _ = plan.inverse(u_full[:1000])
    assert _same_bits(plan.inverse(u_full[:300])[0], xr[:300])
    assert plan.log_prob(X[:0]).shape == (0,)


# --------------------------------------------------------------------------- #
# BASELINE.json sizes: size-independent properties                             #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("engine", ENGINES)
def test_full_size_properties_c2(engine):
    """10 M rows (config 2): determinism, shard independence (the multi-GPU partition computes
    exactly what one GPU computes), round trip and log-det antisymmetry."""
    from pzflow_b200.sharding import shard_rows
    ex = F.example_flow()
    plan = _plan(ex, engine)
    base = torch.from_numpy(F.galaxy_rows(40000, seed=9))
    n = 10_000_000
    X = base.cuda().repeat(n // base.shape[0], 1).contiguous()
    X += torch.linspace(0, 1e-3, X.shape[0], device="cuda")[:, None]  # make rows distinct
    lp1 = plan.log_prob(X)
    lp2 = plan.log_prob(X)
    assert torch.equal(lp1, lp2)
    for r in range(8):
        a, b = shard_rows(n, r, 8)
        assert torch.equal(plan.log_prob(X[a:b]), lp1[a:b])
    u, ld = plan.forward(X[:2_000_000])
    xr, ldi = plan.inverse(u)
    inside = (u.abs() < 4.9).all(dim=1) & torch.isfinite(lp1[:2_000_000]) & (lp1[:2_000_000] > -30)
    err = (xr - X[:2_000_000]).abs()[inside]
    assert inside.float().mean() > 0.5
    assert err.median() < 1e-5 and torch.quantile(err.flatten()[:4_000_000], 0.99) < 2e-3
    asym = (ld + ldi).abs()[inside]
    assert asym.median() < 1e-4
    # the latent of the shipped flow is uniform: finite log_prob == log_det + const
    fin = lp1[:2_000_000] > -1e30
    assert torch.allclose(lp1[:2_000_000][fin], ld[fin] + (-7 * np.log(np.float32(10.0))), atol=1e-5)


@pytest.mark.parametrize("engine", ENGINES)
def test_posterior_properties_c3_slice(engine):
    """Config 3 geometry (G = 1000) on a slice of galaxies: every pdf integrates to one, equals
    exp(log_prob) of the expanded rows, and is independent of how galaxies are batched."""
    ex = F.example_flow()
    plan = _plan(ex, engine)
    X = F.galaxy_rows(2048, seed=5, oob_frac=0)
    rest = torch.from_numpy(np.ascontiguousarray(X[:, 1:])).cuda()
    grid = torch.linspace(0.0, 2.3, 1000)
    pdf = plan.posterior(rest, 0, grid)
    assert pdf.shape == (2048, 1000)
    integ = 0.5 * ((grid[1:] - grid[:-1]).cuda() * (pdf[:, 1:] + pdf[:, :-1])).sum(1)
    assert torch.allclose(integ, torch.ones_like(integ), rtol=3e-6)
    assert torch.equal(pdf[100:900], plan.posterior(rest[100:900], 0, grid))
    un = plan.posterior(rest[:64], 0, grid, normalize=False)
    rows = torch.cat([grid.cuda().repeat(64)[:, None], rest[:64].repeat_interleave(1000, 0)], dim=1)
    assert torch.equal(un.flatten(), torch.exp(plan.log_prob(rows)))


def test_ensemble_reductions_match_oracle():
    from pzflow_b200.plan import ensemble_logmeanexp, ensemble_posterior_mean
    rng = np.random.default_rng(0)
    lp = rng.normal(0, 30, size=(4, 5000)).astype(np.float32)
    lp[:, :10] = -200.0  # all flows underflow -> -inf, the naive form (flowEnsemble.py:243)
    lp[1, 20] = np.finfo(np.float32).min
    with np.errstate(all="ignore"):
        ref = np.log(np.exp(lp).mean(axis=0, dtype=np.float32))
    out = _np(ensemble_logmeanexp(torch.from_numpy(lp).cuda()))
    assert (np.isneginf(out) == np.isneginf(ref)).all() and np.isneginf(out[:10]).all()
    fin = np.isfinite(ref)
    np.testing.assert_allclose(out[fin], ref[fin], rtol=2e-6, atol=2e-6)
    pd_ = rng.uniform(0, 3, size=(3, 40, 33)).astype(np.float32)
    grid = np.sort(rng.uniform(0, 2, 33)).astype(np.float32)
    mean = pd_.mean(axis=0)
    want = mean / O.trapezoid(mean, grid)[:, None]
    got = _np(ensemble_posterior_mean(torch.from_numpy(pd_).cuda(), torch.from_numpy(grid), True))
    np.testing.assert_allclose(got, want, rtol=5e-6)


def test_ensemble_log_prob_fused_in_the_host_pipeline():
    """FlowEnsemble.log_prob(DataFrame) takes pzb_ensemble_log_prob_host_columns (rows staged once, all flows and
    the log-mean-exp on the device per chunk): same values as the flows evaluated one by one + the ensemble rule,
    unconditional and conditional, float64 and float32 frames; returnEnsemble keeps the per-flow route."""
    from pzflow_b200 import Flow, FlowEnsemble
    cfgs = F.config_c4(3)
    dicts = {f"Flow {i}": {"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": ("a", "b"),
                           "condition_means": np.array([0.5, -1.0], np.float32),
                           "condition_stds": np.array([2.0, 0.25], np.float32), "data_error_model": None,
                           "condition_error_model": None, "autoscale_conditions": True, "info": None,
                           "latent_info": c["latent_info"], "bijector_info": c["bijector_info"], "params": c["params"]}
             for i, c in enumerate(cfgs)}
    ens = FlowEnsemble.__new__(FlowEnsemble)
    ens._ensemble = {k: Flow(_dictionary=dict(v)) for k, v in dicts.items()}
    ens.data_columns, ens.conditional_columns = F.GALAXY_COLUMNS, ("a", "b")
    rng = np.random.default_rng(8)
    n = 700_001                                            # several pipeline chunks
    ex = F.example_flow()
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    df = pd.DataFrame(rng.uniform(mins, maxs, size=(n, 7)), columns=F.GALAXY_COLUMNS)
    df["a"], df["b"] = rng.normal(0.5, 2.0, n), rng.normal(-1.0, 0.25, n)
    per_flow = ens.log_prob(df, returnEnsemble=True)       # [N, 3], the per-flow route
    assert per_flow.shape == (n, 3)
    with np.errstate(divide="ignore", over="ignore"):
        want = np.log(np.exp(per_flow.astype(np.float64)).mean(axis=1))
    got = ens.log_prob(df)
    # (the ensemble rule is the reference's literal float32 log(mean(exp(lp))): below lp ~ -87 exp() is subnormal,
    # so the comparison with the float64 evaluation stops at -80)
    fin = np.isfinite(want) & (want > -80.0)
    assert fin.mean() > 0.9
    np.testing.assert_allclose(got[fin], want[fin], rtol=2e-6, atol=2e-6)
    assert np.isfinite(got[fin]).all()
    got32 = ens.log_prob(df.astype(np.float32))
    np.testing.assert_allclose(got32[fin], want[fin], rtol=1e-4, atol=1e-4)   # inputs rounded to float32 first
    # posterior: mean of the flows' un-normalised pdfs, then the trapezoid normalisation (flowEnsemble.py:346-351)
    m, grid = 30_000, np.linspace(0.0, 2.3, 700).astype(np.float32)          # 84 MB of pdfs: several chunks
    sub = df.iloc[:m]
    raw = np.stack([flow.posterior(sub, column="redshift", grid=grid, normalize=False)
                    for flow in ens._ensemble.values()], axis=0).astype(np.float64)
    mean = raw.mean(axis=0)
    area = np.trapezoid(mean, grid.astype(np.float64), axis=1)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        want_pdf = np.nan_to_num(mean / area, nan=0.0)
    got_pdf = ens.posterior(sub, column="redshift", grid=grid)
    assert got_pdf.shape == (m, 700)
    np.testing.assert_allclose(got_pdf, want_pdf, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(ens.posterior(sub, column="redshift", grid=grid, normalize=False), mean, rtol=1e-5,
                               atol=1e-30)
    # flows that scale their conditions differently are evaluated one by one (and still agree with the rule)
    ens._ensemble["Flow 1"]._condition_means = np.array([0.4, -1.0], np.float32)
    assert FlowEnsemble._log_prob_fused(list(ens._ensemble.values()), df, list(F.GALAXY_COLUMNS)) is None
    assert ens.log_prob(df[:1000]).shape == (1000,)


def test_uniform_sampler():
    from pzflow_b200.plan import sample_uniform
    u = sample_uniform(200_000, 7, 5.0, seed=3)
    assert u.shape == (200_000, 7) and float(u.min()) >= -5.0 and float(u.max()) < 5.0
    assert abs(float(u.mean())) < 0.02 and abs(float(u.std()) - 10 / np.sqrt(12)) < 0.02
    assert torch.equal(u, sample_uniform(200_000, 7, 5.0, seed=3))
    assert not torch.equal(u, sample_uniform(200_000, 7, 5.0, seed=4))


def test_native_latent_draws_follow_their_distributions():
    """CentBeta13 / CentBeta / Normal / Tdist draws come from the library's own Philox kernels:
    Kolmogorov-Smirnov distance to the distribution the reference draws from (pzflow/distributions.py:101-135, 196-229,
    277-303, 360-392), reproducible per seed, different per seed and per column."""
    from scipy import stats
    from pzflow_b200 import distributions as Dz
    from pzflow_b200.plan import launch_count
    n = 200_000
    l0 = launch_count()
    u = Dz.CentBeta13(3, 5).sample((), n, seed=7)
    assert u.shape == (n, 3) and np.abs(u).max() < 5.0
    for j in range(3):
        assert stats.kstest((u[:, j] + 5.0) / 10.0, stats.beta(13, 13).cdf).statistic < 0.005
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.01
    assert np.array_equal(u, Dz.CentBeta13(3, 5).sample((), n, seed=7))
    assert not np.array_equal(u, Dz.CentBeta13(3, 5).sample((), n, seed=8))
    cb = Dz.CentBeta(2, 4)
    cb_params = ((np.log(0.6), np.log(2.5)), (np.log(7.0), np.log(0.8)))     # shapes below one take the boost path
    v = cb.sample(cb_params, n, seed=1)
    assert stats.kstest((v[:, 0] + 4.0) / 8.0, stats.beta(0.6, 2.5).cdf).statistic < 0.005
    assert stats.kstest((v[:, 1] + 4.0) / 8.0, stats.beta(7.0, 0.8).cdf).statistic < 0.005
    z = Dz.Normal(5).sample((), n, seed=2)
    assert stats.kstest(z.ravel(), stats.norm.cdf).statistic < 0.003 and abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01
    t = Dz.Tdist(3).sample(np.log(4.5), n, seed=3)
    assert stats.kstest(t[:, 1], stats.t(4.5).cdf).statistic < 0.005
    # one chi2 per ROW: |t|^2 / 3 is F(3, nu) distributed, which independent columns would not give
    assert stats.kstest((t * t).sum(axis=1) / 3.0, stats.f(3, 4.5).cdf).statistic < 0.005
    assert launch_count() - l0 == 6          # every draw above was one kernel of this library
    j = Dz.Joint(Dz.Uniform(1, 5), Dz.CentBeta13(2, 5), Dz.Normal(1)).sample([(), ((0.0, 0.0),) * 2, ()], 1000, seed=5)
    assert j.shape == (1000, 4)


# --------------------------------------------------------------------------- #
# host-buffer entry points (pzb_*_host): chunked H2D | kernel | D2H pipeline    #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("pinned", [True, False])
def test_host_pipeline_equals_device_path(pinned):
    """Several pipeline chunks (3.7 M rows > 3 x 1.2 M), ragged tail, pinned and pageable host
    arrays: the host-buffer call returns exactly what the device-pointer call computes."""
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    n = 3_700_017
    base = torch.from_numpy(F.galaxy_rows(50_000, seed=11))
    X = base.repeat(n // base.shape[0] + 1, 1)[:n].contiguous()
    X += torch.linspace(0, 1e-3, n)[:, None]
    if pinned:
        X = X.pin_memory()
    want = plan.log_prob(X.cuda())
    got = plan.log_prob_host(X)
    assert got.device.type == "cpu" and got.shape == (n,)
    assert torch.equal(got, want.cpu())
    assert torch.equal(plan.log_prob(X), got)  # Plan.log_prob routes host arrays to the host path
    assert torch.equal(plan.log_prob_host(X.numpy()), got)
    # forward / inverse round trip through the host forms
    m = 1_300_000
    u_d, ld_d = plan.forward(X[:m].cuda())
    u_h, ld_h = plan.forward_host(X[:m])
    assert _same_bits(u_h, u_d) and _same_bits(ld_h, ld_d)
    x_d, ldi_d = plan.inverse(u_d)
    x_h, ldi_h = plan.inverse_host(u_h, want_log_det=True)
    assert _same_bits(x_h, x_d) and _same_bits(ldi_h, ldi_d)
    assert plan.inverse_host(u_h[:1000])[1] is None
    assert plan.log_prob_host(X[:0]).shape == (0,)


def test_host_pipeline_pageable_results_are_staged_and_identical():
    """Results bound for ordinary (pageable) host memory go through the library's page-locked staging ring
    (csrc/pzb_api.cu: run_host, stage_out): same bits as the page-locked destination and as the device path."""
    from pzflow_b200 import _native as N
    from pzflow_b200.plan import _ptr, pinned_empty
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    X = F.galaxy_rows(6000, seed=3)
    grid = np.linspace(0.0, 2.3, 1000).astype(np.float32)                  # 6000 x 1000 floats = 24 MB > 8 MB
    rest = np.ascontiguousarray(X[:, 1:])
    pageable = plan.posterior_host(rest, 0, grid, out=torch.empty((6000, 1000)))
    assert not pageable.is_pinned()
    from pzflow_b200.plan import host_result
    assert host_result((6000, 1000)).is_pinned() and not host_result((70_000, 1000)).is_pinned()   # 24 MB / 280 MB
    pinned = plan.posterior_host(rest, 0, grid, out=pinned_empty((6000, 1000)))
    dev = plan.posterior(torch.from_numpy(rest), 0, torch.from_numpy(grid))
    assert _same_bits(pageable, pinned) and _same_bits(pageable, dev)
    n = 700_000                                                           # forward: [n, 7] = 19.6 MB, plus log_det
    Xh = torch.from_numpy(np.tile(X, (n // 6000 + 1, 1))[:n].copy())
    u, ld = torch.empty((n, 7)), torch.empty(n)
    N.check(N.lib().pzb_forward_host(plan._h, _ptr(Xh), None, n, _ptr(u), _ptr(ld)))
    u_dev, ld_dev = plan.forward(Xh)
    assert _same_bits(u, u_dev) and _same_bits(ld, ld_dev)
    # Flow.sample's route: latent draws on the device, samples in host memory (pzb_inverse_to_host)
    x_host, ldi_host = plan.inverse_to_host(u_dev, want_log_det=True)
    x_dev, ldi_dev = plan.inverse(u_dev)
    assert not x_host.is_cuda and _same_bits(x_host, x_dev) and _same_bits(ldi_host, ldi_dev)
    assert plan.inverse_to_host(u_dev[:0])[0].shape == (0, 7)
    x_cols = plan.inverse_to_host_columns(u_dev)          # [N, D] view of a C-contiguous [D, N] buffer
    assert x_cols.shape == (n, 7) and x_cols.T.flags.c_contiguous
    assert _same_bits(torch.from_numpy(np.ascontiguousarray(x_cols)), x_dev)
    frame = pd.DataFrame(x_cols, columns=list("abcdefg"), copy=False)
    assert np.shares_memory(frame["c"].to_numpy(), x_cols)             # pandas wraps it without copying


def test_host_pipeline_posterior_and_conditions():
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    X = F.galaxy_rows(3000, seed=6, oob_frac=0)  # 3000 galaxies x 1000 grid points: 3 chunks
    rest = torch.from_numpy(np.ascontiguousarray(X[:, 1:]))
    grid = torch.linspace(0.0, 2.3, 1000)
    want = plan.posterior(rest.cuda(), 0, grid)
    got = plan.posterior_host(rest, 0, grid)
    assert torch.equal(got, want.cpu())
    un = plan.posterior_host(rest[:70], 0, grid, normalize=False, nan_to_zero=False)
    assert torch.equal(un, plan.posterior(rest[:70].cuda(), 0, grid, normalize=False, nan_to_zero=False).cpu())
    # conditional plan: conditions ride the same ring
    cfg = F.config_c4(1)[0]
    cplan = _plan(cfg, "tc")
    rng = np.random.default_rng(2)
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    n = 2_500_003
    Xc = torch.from_numpy(rng.uniform(mins, maxs, size=(n, 7)).astype(np.float32))
    cond = torch.from_numpy(rng.normal(size=(n, 2)).astype(np.float32))
    assert torch.equal(cplan.log_prob_host(Xc, cond), cplan.log_prob(Xc.cuda(), cond.cuda()).cpu())
    with pytest.raises(ValueError):
        cplan.log_prob_host(Xc)


def test_flow_api_host_arrays_take_the_pipeline():
    """Flow.log_prob / Flow.posterior on DataFrames (the reference's call) end in the host-buffer
    entry points and agree with the device path."""
    from pzflow_b200 import Flow
    ex = F.example_flow()
    flow = Flow(_dictionary={"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": None,
                             "condition_means": None, "condition_stds": None, "data_error_model": None,
                             "condition_error_model": None, "autoscale_conditions": True, "info": None,
                             "latent_info": ex["latent_info"], "bijector_info": ex["bijector_info"],
                             "params": ex["params"]})
    X = F.galaxy_rows(5000, seed=8)
    df = pd.DataFrame(X.astype(np.float64), columns=F.GALAXY_COLUMNS)
    lp = flow.log_prob(df)
    assert isinstance(lp, np.ndarray) and lp.dtype == np.float32
    np.testing.assert_array_equal(lp, _np(flow._plan().log_prob(torch.from_numpy(X).cuda())))
    grid = np.linspace(0, 2.3, 200)
    pdfs = flow.posterior(df[:300], column="redshift", grid=grid)
    pb = flow.posterior(df[:300], column="redshift", grid=grid, batch_size=64)
    np.testing.assert_array_equal(pdfs, pb)  # tests/test_flow.py:275-287 of the reference
    want = flow._plan().posterior(torch.from_numpy(np.ascontiguousarray(X[:300, 1:])).cuda(), 0,
                                  torch.from_numpy(grid.astype(np.float32)))
    np.testing.assert_array_equal(pdfs, _np(want))


# --------------------------------------------------------------------------- #
# Gaussian error convolution (err_samples), SURVEY.md §8f-1                    #
# --------------------------------------------------------------------------- #
def _expand(X, Xerr, eps, S):
    """gaussian_error_model (pzflow/utils.py:52-61) given eps [N*S, W] -> rows [N*S, W], float32."""
    n, w = X.shape
    return (X[:, None, :] + eps.reshape(n, S, w) * Xerr[:, None, :]).reshape(n * S, w).astype(np.float32)


def _jax_eps(n, S, W, seed, skip=None, part=False):
    """The eps the reference's default error model draws (pzflow/utils.py:59 under flow.py:455-457):
    jax.random.normal(PRNGKey(seed), (n, S, W)) restated by the oracle -> [n * S, W] (``skip``: the posterior's evaluated
    column rides along in the draw and is deleted afterwards, flow.py:370-384)."""
    from oracle import jax_prng as J
    eps = J.normal(J.PRNGKey(seed), (n, S, W), part)
    if skip is not None:
        eps = np.delete(eps, skip, axis=2)
    return np.ascontiguousarray(eps.reshape(n * S, -1))


@pytest.fixture
def philox_streams():
    from pzflow_b200 import prng
    was = prng.layout()
    prng.set_layout("philox")
    yield
    prng.set_layout(was)


def test_error_samples_are_jax_random_normal(stream_layout):
    """The default stream of the in-kernel error samples IS jax.random.normal(PRNGKey(seed), (N, S, W)): the
    materialised form (pzb_sample_normal_threefry == the oracle's restatement, which prints jax's documented values),
    then Flow.log_prob(err_samples=S) == the oracle fed those eps -- on every engine, for data and conditions, cut
    into launches or not."""
    from oracle import jax_prng as J
    from pzflow_b200 import distributions as Dn
    from pzflow_b200.plan import sample_normal_jax
    from pzflow_b200.utils import gaussian_error_model
    part = stream_layout == "threefry_partitionable"
    # "JAX - The Sharp Bits" and the PRNG tutorial of jax's documentation, before and after jax 0.5 switched layouts
    assert float(sample_normal_jax(1, 1, 0, layout="threefry")[0, 0]) == np.float32(-0.20584226)
    assert float(sample_normal_jax(1, 1, 42, layout="threefry")[0, 0]) == np.float32(-0.18471177)
    assert float(sample_normal_jax(1, 1, 0, layout="threefry_partitionable")[0, 0]) == np.float32(1.6226422)
    assert float(sample_normal_jax(1, 1, 42, layout="threefry_partitionable")[0, 0]) == np.float32(-0.028304616)
    for n, w, seed in [(1, 1, 0), (1001, 7, 5), (4096, 3, 2 ** 40 + 17), (333, 2, 9)]:
        want = J.normal(J.PRNGKey(seed), (n, w), part)
        got = _np(sample_normal_jax(n, w, seed))
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert (ulp <= 4).all() and (ulp == 0).mean() > 0.95, (n, w, ulp.max(), (ulp == 0).mean())   # log1pf vs np.log1p
        lo, hi = n // 3, n // 3 + max(1, n // 2)                                     # a row block of the same call
        np.testing.assert_array_equal(_np(sample_normal_jax(hi - lo, w, seed, row_offset=lo, total_rows=n)), got[lo:hi])
        other = _np(sample_normal_jax(n, w, seed, layout="threefry" if part else "threefry_partitionable"))
        assert np.abs(other - J.normal(J.PRNGKey(seed), (n, w), not part)).max() < 1e-6
        assert _np(Dn.Normal(w).sample_device(n, seed)).tobytes() == got.tobytes()   # the Normal latent draws this stream
    X = np.arange(12, dtype=np.float32).reshape(4, 3)
    g = gaussian_error_model(np.array([0, 3], np.uint32), X, np.full_like(X, 0.5), 5)
    np.testing.assert_allclose(g, X[:, None, :] + 0.5 * J.normal(J.PRNGKey(3), (4, 5, 3), part), rtol=0, atol=1e-6)
    ex = F.example_flow()
    rng = np.random.default_rng(8)
    n, S, seed = 700, 6, 31
    X = F.galaxy_rows(n, seed=24, oob_frac=0)
    Xerr = rng.uniform(0.0, 0.05, size=X.shape).astype(np.float32)
    rows = _expand(X, Xerr, _jax_eps(n, S, 7, seed, part=part), S)
    bi, li, p64 = _o64(ex)
    with np.errstate(all="ignore"):
        lp32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], rows)
        lp64 = O.flow_log_prob(bi, li, p64, rows.astype(np.float64))
        o32 = np.log(np.exp(lp32.reshape(n, S)).mean(axis=1, dtype=np.float32))
        o64 = np.log(np.exp(lp64.reshape(n, S)).mean(axis=1))
    for engine in ENGINES + ["wide"]:
        plan = _plan(ex, engine)
        got = plan.log_prob_err(X, Xerr, None, None, S, seed)
        assert_noise_floor(f"C2 log_prob, jax-stream err_samples [{engine}]", _np(got), o32, o64, ratio=2.0, slack=5e-6,
                           below=LP_UNDERFLOW)
        plan.ERR_SCRATCH_FLOATS = 1024
        try:
            assert _same_bits(plan.log_prob_err(X, Xerr, None, None, S, seed), got)
        finally:
            del plan.ERR_SCRATCH_FLOATS
    # the posterior: the evaluated column rides along in the draw (W = 7), conditions use the same key with W = C
    ng, G = 30, 40
    rest, rest_err = np.ascontiguousarray(X[:ng, 1:]), np.ascontiguousarray(Xerr[:ng, 1:])
    grid = np.linspace(0.0, 2.3, G).astype(np.float32)
    samples = _expand(rest, rest_err, _jax_eps(ng, S, 7, seed, skip=0, part=part), S)
    un = O.flow_posterior(ex["bijector_info"], ex["latent_info"], ex["params"], samples, 0, grid, normalize=False,
                          nan_to_zero=False)
    want = un.reshape(ng, S, G).mean(axis=1)
    want = want / O.trapezoid(want, grid)[:, None]
    for engine in ENGINES + ["wide"]:
        got = _np(_plan(ex, engine).posterior_err(rest, rest_err, 0, grid, None, None, S, seed))
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4)
    cfg = F.config_c4(1)[0]
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    n2 = 400
    Xc = rng.uniform(mins, maxs, size=(n2, 7)).astype(np.float32)
    cond, cerr = rng.normal(size=(n2, 2)).astype(np.float32), rng.uniform(0, 0.1, size=(n2, 2)).astype(np.float32)
    xerr = rng.uniform(0, 0.02, size=(n2, 7)).astype(np.float32)
    with np.errstate(all="ignore"):
        lp = O.flow_log_prob(cfg["bijector_info"], cfg["latent_info"], cfg["params"],
                             _expand(Xc, xerr, _jax_eps(n2, S, 7, seed, part=part), S), _expand(cond, cerr, _jax_eps(n2, S, 2, seed, part=part), S))
        want = np.log(np.exp(lp.reshape(n2, S)).mean(axis=1, dtype=np.float32))
    ok = want > LP_UNDERFLOW
    for engine in ENGINES + ["wide"]:
        got = _np(_plan(cfg, engine).log_prob_err(Xc, xerr, cond, cerr, S, seed))
        assert within_tol(got[ok], want[ok]).mean() > 0.98, engine


@pytest.mark.parametrize("engine", ENGINES)
def test_err_samples_log_prob_matches_oracle(engine, philox_streams):
    from pzflow_b200.plan import error_eps
    ex = F.example_flow()
    plan = _plan(ex, engine)
    rng = np.random.default_rng(3)
    n, S, seed = 1500, 16, 12345
    X = F.galaxy_rows(n, seed=21, oob_frac=0)
    Xerr = (rng.uniform(0.0, 0.05, size=X.shape) * (rng.random(X.shape) > 0.3)).astype(np.float32)
    eps = _np(error_eps(n * S, 7, 0, seed))
    assert abs(eps.mean()) < 0.01 and abs(eps.std() - 1.0) < 0.01 and abs((eps ** 3).mean()) < 0.03
    rows = _expand(X, Xerr, eps, S)
    bi, li, p64 = _o64(ex)
    with np.errstate(all="ignore"):
        lp32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], rows)
        lp64 = O.flow_log_prob(bi, li, p64, rows.astype(np.float64))
        o32 = np.log(np.exp(lp32.reshape(n, S)).mean(axis=1, dtype=np.float32))
        o64 = np.log(np.exp(lp64.reshape(n, S)).mean(axis=1))
    got = _np(plan.log_prob_err(X, Xerr, None, None, S, seed))
    assert_noise_floor(f"C2 log_prob err_samples [{engine}]", got, o32, o64, ratio=2.0, slack=5e-6, below=LP_UNDERFLOW)
    # the expansion is generated per expanded row: cutting the rows into launches cannot change it
    plan.ERR_SCRATCH_FLOATS = 4096
    try:
        assert _same_bits(plan.log_prob_err(X, Xerr, None, None, S, seed), torch.from_numpy(got))
    finally:
        del plan.ERR_SCRATCH_FLOATS
    # zero errors == plain log_prob (tests/test_flow.py:236-239 of the reference)
    plain = _np(plan.log_prob(torch.from_numpy(X).cuda()))
    zero = _np(plan.log_prob_err(X, None, None, None, 4, seed))
    fin = plain > LP_UNDERFLOW
    np.testing.assert_allclose(zero[fin], plain[fin], rtol=1e-5, atol=1e-5)
    assert not np.array_equal(got, _np(plan.log_prob_err(X, Xerr, None, None, S, seed + 1)))


def test_err_samples_posterior_and_conditions(philox_streams):
    from pzflow_b200.plan import error_eps
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    rng = np.random.default_rng(4)
    ng, S, G, seed = 40, 8, 50, 7
    X = F.galaxy_rows(ng, seed=22, oob_frac=0)
    rest = np.ascontiguousarray(X[:, 1:])
    rest_err = rng.uniform(0.0, 0.05, size=rest.shape).astype(np.float32)
    grid = np.linspace(0.0, 2.3, G).astype(np.float32)
    eps = _np(error_eps(ng * S, 6, 0, seed))
    samples = _expand(rest, rest_err, eps, S)                      # [ng*S, 6]
    un = O.flow_posterior(ex["bijector_info"], ex["latent_info"], ex["params"], samples, 0, grid, normalize=False,
                          nan_to_zero=False)                       # [ng*S, G]
    want = un.reshape(ng, S, G).mean(axis=1)
    want = want / O.trapezoid(want, grid)[:, None]
    got = _np(plan.posterior_err(rest, rest_err, 0, grid, None, None, S, seed))
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4)
    # zero errors == plain posterior (tests/test_flow.py:260-264: rtol 1e-4)
    plain = _np(plan.posterior(torch.from_numpy(rest).cuda(), 0, torch.from_numpy(grid)))
    zero = _np(plan.posterior_err(rest, None, 0, grid, None, None, 3, seed))
    np.testing.assert_allclose(zero, plain, rtol=1e-4, atol=1e-6)
    # conditional flow: condition errors ride stream 1
    cfg = F.config_c4(1)[0]
    cplan = _plan(cfg, "tc")
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    n, S2 = 600, 5
    Xc = rng.uniform(mins, maxs, size=(n, 7)).astype(np.float32)
    cond = rng.normal(size=(n, 2)).astype(np.float32)
    cerr = rng.uniform(0, 0.1, size=(n, 2)).astype(np.float32)
    xerr = rng.uniform(0, 0.02, size=(n, 7)).astype(np.float32)
    rows = _expand(Xc, xerr, _np(error_eps(n * S2, 7, 0, seed)), S2)
    crow = _expand(cond, cerr, _np(error_eps(n * S2, 2, 1, seed)), S2)
    with np.errstate(all="ignore"):
        lp = O.flow_log_prob(cfg["bijector_info"], cfg["latent_info"], cfg["params"], rows, crow)
        want = np.log(np.exp(lp.reshape(n, S2)).mean(axis=1, dtype=np.float32))
    got = _np(cplan.log_prob_err(Xc, xerr, cond, cerr, S2, seed))
    ok = want > LP_UNDERFLOW
    assert within_tol(got[ok], want[ok]).mean() > 0.98


def test_err_samples_through_flow_api():
    """DataFrame API: *_err columns, missing ones are zero (flow.py:366-369); ensemble rule on top."""
    from pzflow_b200 import Flow, FlowEnsemble
    ex = F.example_flow()
    d = {"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": None,
         "condition_means": None, "condition_stds": None, "data_error_model": None,
         "condition_error_model": None, "autoscale_conditions": True, "info": None,
         "latent_info": ex["latent_info"], "bijector_info": ex["bijector_info"], "params": ex["params"]}
    flow = Flow(_dictionary=d)
    X = F.galaxy_rows(300, seed=23, oob_frac=0)
    df = pd.DataFrame(X.astype(np.float64), columns=F.GALAXY_COLUMNS)
    df["u_err"] = 0.03
    df["redshift_err"] = 0.01
    lp = flow.log_prob(df, err_samples=10, seed=5)
    assert lp.shape == (300,) and np.isfinite(lp).mean() > 0.9
    np.testing.assert_array_equal(lp, flow.log_prob(df, err_samples=10, seed=5))   # seed reproducibility
    Xerr = np.zeros_like(X)
    Xerr[:, 0], Xerr[:, 1] = 0.01, 0.03
    np.testing.assert_array_equal(lp, _np(flow._plan().log_prob_err(X, Xerr, None, None, 10, 5)))
    grid = np.linspace(0, 2.3, 40)
    pdfs = flow.posterior(df[:50], column="redshift", grid=grid, err_samples=6, seed=5)
    assert pdfs.shape == (50, 40)
    np.testing.assert_allclose(np.trapezoid(pdfs, grid, axis=1), 1.0, rtol=1e-4)
    # with a batch_size the reference draws every batch from the same key (flow.py:668-693): batch k of the call equals
    # the call on batch k alone, and differs from the rows of the unbatched call after the first batch
    batched = flow.posterior(df[:50], column="redshift", grid=grid, err_samples=6, seed=5, batch_size=20)
    for a in (0, 20, 40):
        alone = flow.posterior(df[a:min(50, a + 20)], column="redshift", grid=grid, err_samples=6, seed=5)
        np.testing.assert_array_equal(batched[a:a + 20], alone)
    assert np.abs(batched[20:40] - pdfs[20:40]).max() > 1e-4
    with pytest.raises(AssertionError):
        flow.log_prob(df, err_samples=0)
    # a user's own error model is a host callable (test_custom_error_models_...); it cannot ride inside marg_rules
    flow2 = Flow(_dictionary=dict(d, data_error_model=lambda key, X, Xerr, n: np.repeat(X[:, None, :], n, axis=1)))
    fin = flow.log_prob(df[:64]) > LP_UNDERFLOW
    np.testing.assert_allclose(flow2.log_prob(df[:64], err_samples=2)[fin], flow.log_prob(df[:64])[fin], rtol=1e-5, atol=1e-5)
    with pytest.raises(NotImplementedError):
        flow2.posterior(df[:4], column="redshift", grid=grid, err_samples=2,
                        marg_rules={"flag": 99.0, "u": lambda row: np.linspace(25, 31, 4)})


# --------------------------------------------------------------------------- #
# parameter-free bijectors and the Normal / CentBeta latents   #
# --------------------------------------------------------------------------- #
def test_extra_bijectors_golden_through_protocol():
    from pzflow_b200.utils import build_bijector_from_info
    z = golden("ref_shim_extra.npz")
    x = z["x"]
    cond = np.zeros((x.shape[0], 1), np.float32)
    for name in z["names"]:
        name = str(name)
        info = F.unpack_tree(f"{name}_info", z)
        params = F.unpack_tree(f"{name}_params", z)
        init_fun, _ = build_bijector_from_info(info)
        _, fwd, inv = init_fun(0, 7)
        y, ld = fwd(params, x, conditions=cond)
        xi, ldi = inv(params, y, conditions=cond)
        xi2, ldi2 = inv(params, x, conditions=cond)
        for got, key in ((y, "fwd"), (ld, "fwd_ld"), (xi, "inv_of_fwd"), (ldi, "inv_of_fwd_ld"), (xi2, "inv"),
                         (ldi2, "inv_ld")):
            want = z[f"{name}_{key}"]
            ok = np.isfinite(want)
            assert (np.isfinite(got) == ok).mean() > 0.99, (name, key)
            both = ok & np.isfinite(got)
            assert within_tol(got[both], want[both]).mean() > 0.99, (name, key, np.abs(got - want)[both].max())


def test_normal_and_centbeta_latents():
    z = golden("ref_shim_extra.npz")
    from pzflow_b200.plan import Plan
    from pzflow_b200 import distributions as Dn
    u = z["latent_u"]
    ident = ("Chain", (("Roll", (0,)),))   # u = x: log_prob == latent log-density
    plan = Plan(ident, [()], 3, ("Normal", (3,)))
    np.testing.assert_allclose(_np(plan.log_prob(torch.from_numpy(u).cuda())), z["normal_lp"], rtol=3e-6, atol=3e-6)
    np.testing.assert_allclose(Dn.Normal(3).log_prob((), u), z["normal_lp"], rtol=3e-6, atol=3e-6)
    cbp = tuple((float(a), float(b)) for a, b in z["centbeta_params"])
    for params, key in ((cbp, "centbeta_lp"), (None, "centbeta_lp_default")):
        plan = Plan(ident, [()], 3, ("CentBeta", (3, 5)), latent_params=params)
        got, want = _np(plan.log_prob(torch.from_numpy(u).cuda())), z[key]
        zero = ~np.isfinite(want)
        assert (zero_class(got) == zero).all()
        np.testing.assert_allclose(got[~zero], want[~zero], rtol=1e-5, atol=1e-5)
        lat = Dn.CentBeta(3, 5)
        got2 = lat.log_prob(lat._params if params is None else params, u)
        np.testing.assert_allclose(got2[~zero], want[~zero], rtol=1e-5, atol=1e-5)
    s = Dn.CentBeta(3, 5).sample(cbp, 50_000, seed=1)
    assert s.shape == (50_000, 3) and np.abs(s).max() <= 5.0
    a, b = np.exp(z["centbeta_params"][:, 0]), np.exp(z["centbeta_params"][:, 1])
    np.testing.assert_allclose(s.mean(axis=0), 10 * (a / (a + b) - 0.5), atol=0.05)
    n = Dn.Normal(3).sample((), 50_000, seed=1)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02
    # Tdist (learned degrees of freedom) and Joint (segments of different latents)
    for params, key in ((None, "tdist_lp_default"), (float(z["tdist_lognu"]), "tdist_lp")):
        plan = Plan(ident, [()], 3, ("Tdist", (3,)), latent_params=params)
        np.testing.assert_allclose(_np(plan.log_prob(torch.from_numpy(u).cuda())), z[key], rtol=1e-5, atol=1e-5)
        td = Dn.Tdist(3)
        np.testing.assert_allclose(td.log_prob(td._params if params is None else params, u), z[key], rtol=1e-5, atol=1e-5)
    t = Dn.Tdist(3).sample(float(z["tdist_lognu"]), 100_000, seed=2)
    from scipy import stats
    assert t.shape == (100_000, 3) and abs(np.median(t)) < 0.02
    assert abs(np.percentile(np.abs(t), 84) - stats.t.ppf(0.92, 4.5)) < 0.03   # |t| 84th pct == t 92nd pct
    jinfo = F.unpack_tree("joint_info", z)
    jparams = [(), (cbp[0],), ()]
    plan = Plan(ident, [()], 3, jinfo, latent_params=jparams)
    got, want = _np(plan.log_prob(torch.from_numpy(u).cuda())), z["joint_lp"]
    zero = ~np.isfinite(want)
    assert (zero_class(got) == zero).all()
    np.testing.assert_allclose(got[~zero], want[~zero], rtol=1e-5, atol=1e-5)
    joint = Dn.Joint(*jinfo[1])
    assert joint.info == jinfo and joint.input_dim == 3
    got2 = joint.log_prob(jparams, u)
    np.testing.assert_allclose(got2[~zero], want[~zero], rtol=1e-5, atol=1e-5)
    js = joint.sample(jparams, 20_000, seed=4)
    assert js.shape == (20_000, 3) and np.abs(js[:, :2]).max() <= 5.0


@pytest.mark.parametrize("engine", ENGINES)
def test_colour_flow_chain_both_engines(engine):
    """A flow shaped like the reference's customizing tutorial: ColorTransform -> InvSoftplus -> StandardScaler
    -> RollingSplineCoupling (default conditioners, so the tensor-core engine takes it) -> Scale, CentBeta latent."""
    info = ("Chain", (("ColorTransform", (3, [1, 2, 3, 4, 5, 6])), ("InvSoftplus", (0, 4.0)),
                      ("StandardScaler", (np.linspace(-0.5, 0.5, 7).astype(np.float32),
                                          np.linspace(0.5, 2.0, 7).astype(np.float32))),
                      ("Reverse", ()), ("RollingSplineCoupling", (3,)), ("Scale", (0.5,))))
    rng = np.random.default_rng(9)
    bparams = [(), (), (), (), O.synth_bijector_params(O.expand_rolling((3,)), 7, rng, out_scale=2.0), ()]
    lat_params = tuple((float(a), float(b)) for a, b in rng.normal(1.0, 0.4, size=(7, 2)))
    cfg = dict(bijector_info=info, params=(lat_params, bparams), input_dim=7, latent_info=("CentBeta", (7, 5)))
    from pzflow_b200.plan import Plan
    plan = Plan(info, bparams, 7, cfg["latent_info"], latent_params=lat_params)
    if engine == "tc" and not plan.tc_supported:
        pytest.skip("tensor-core engine does not support this shape")
    if engine == "wide" and not plan.wide_supported:
        pytest.skip("wide tensor-core engine does not support this shape")
    plan.set_engine(engine)
    X = rng.uniform(0.05, 1.5, size=(4096, 7)).astype(np.float32)
    o32 = O.flow_log_prob(info, cfg["latent_info"], cfg["params"], X)
    got = _np(plan.log_prob(torch.from_numpy(X).cuda()))
    ok = ~zero_class(o32) & (o32 > LP_UNDERFLOW)
    assert (zero_class(got) == zero_class(o32)).mean() > 0.995
    assert within_tol(got[ok], o32[ok]).mean() > 0.97
    u, ld = plan.forward(torch.from_numpy(X).cuda())
    xr, ldi = plan.inverse(u)
    inside = (u.abs() < 2.4).all(dim=1).cpu().numpy() & ok
    assert inside.mean() > 0.3
    np.testing.assert_allclose(_np(xr)[inside], X[inside], atol=2e-4)
    np.testing.assert_allclose(_np(ld)[inside], -_np(ldi)[inside], atol=2e-4)


# --------------------------------------------------------------------------- #
# Flow.train (native spline fwd/bwd kernels + cuBLAS GEMMs)    #
# --------------------------------------------------------------------------- #
def _two_moons(n, seed):
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, np.pi, n)
    up = rng.random(n) < 0.5
    x = np.where(up, np.cos(t), 1 - np.cos(t)) + rng.normal(0, 0.06, n)
    y = np.where(up, np.sin(t), 0.5 - np.sin(t)) + rng.normal(0, 0.06, n)
    return pd.DataFrame({"x": x, "y": y})


def test_flow_train_reduces_the_loss_and_feeds_the_fused_kernels():
    from pzflow_b200 import Flow, FlowEnsemble
    from pzflow_b200.train import TorchChain
    data, val = _two_moons(6000, 0), _two_moons(1500, 1)
    flow = Flow(data_columns=("x", "y"))
    losses, val_losses = flow.train(data, val_set=val, epochs=6, seed=0)
    assert len(losses) == 7 and len(val_losses) == 7 and all(np.isfinite(losses))
    assert losses[-1] < losses[0] - 0.5 and val_losses[-1] < val_losses[0] - 0.5
    assert flow._bijector_info[0] == "Chain" and flow._bijector_info[1][0][0] == "ShiftBounds"   # default bijector
    # the parameters training left behind are what the fused kernels evaluate: same log_prob as the
    # differentiable restatement, and a loss equal to the last reported one (best_params keeps the best epoch)
    lp = flow.log_prob(data)
    chain = TorchChain(flow._bijector_info, flow._params[1], flow._latent_info, flow._params[0], 2,
                       torch.device("cuda"))
    want = chain.log_prob(torch.from_numpy(data.to_numpy().astype(np.float32)).cuda()).detach().cpu().numpy()
    fin = np.isfinite(want) & (want > LP_UNDERFLOW)
    assert within_tol(lp[fin], want[fin]).mean() > 0.99
    assert abs(-lp[fin].mean() - min(losses)) < 0.05
    # samples of the trained flow land on the moons
    s = flow.sample(4000, seed=3)
    assert s.shape == (4000, 2) and abs(s["x"].mean() - data["x"].mean()) < 0.1
    # early stopping and argument checks (flow.py:944-946, 1113-1134)
    with pytest.raises(ValueError):
        flow.train(data, epochs=0)
    ens = FlowEnsemble(data_columns=("x", "y"), N=2)
    out = ens.train(data[:2000], epochs=2, seed=1)
    assert sorted(out) == ["Flow 0", "Flow 1"] and all(len(v) == 3 for v in out.values())
    assert ens.log_prob(data[:100]).shape == (100,)


@pytest.mark.parametrize("K,L,B,periodic", [(16, 4, 5.0, False), (2, 1, 3.0, False), (32, 16, 5.0, False),
                                            (64, 3, 4.0, False), (16, 4, 5.0, True), (8, 2, 3.0, True)])
def test_spline_train_kernels_match_autograd_of_the_restatement(K, L, B, periodic):
    """csrc/nsf_train.cu: forward values and the hand-derived backward against float64 torch autograd over
    train._rqs_forward (the restatement of pzflow/utils.py:64-227), including out-of-bounds and knot-adjacent x."""
    from pzflow_b200.train import _SplineStage, _rqs_forward
    g = torch.Generator().manual_seed(K * 100 + L)
    N = 3000
    logits = (torch.randn(N, L, 3 * K - 1 + int(periodic), generator=g) * 1.5).cuda()
    x = ((torch.rand(N, L, generator=g) * 2 - 1) * (B * 1.1)).cuda()
    x[0, 0], x[1, 0], x[2, 0] = -B, B, 0.0
    gy = torch.randn(N, L, generator=g).cuda()
    gl = torch.randn(N, generator=g).cuda()

    a, b = logits.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y, ld = _SplineStage.apply(a, b, K, B, periodic)
    ((y * gy).sum() + (ld.sum(dim=1) * gl).sum()).backward()

    a64, b64 = logits.double().requires_grad_(True), x.double().requires_grad_(True)
    W = 2 * B * torch.softmax(a64[..., :K], dim=-1)
    H = 2 * B * torch.softmax(a64[..., K:2 * K], dim=-1)
    D = torch.nn.functional.softplus(a64[..., 2 * K:])
    y64, ld64 = _rqs_forward(b64, W, H, D, B, periodic)
    ((y64 * gy.double()).sum() + (ld64 * gl.double()).sum()).backward()

    # rows whose x sits within float32 rounding of a knot may pick the neighbouring bin: same value (the spline is
    # C1) but a different gradient pattern; they are rare and excluded from the gradient comparison only
    # (narrow bins amplify the float32 rounding of xi = (x - xk) / w: bound the bulk tightly, the tail loosely)
    dy, dl = np.abs(_np(y) - _np(y64)), np.abs(_np(ld.sum(dim=1)) - _np(ld64))
    assert (dy <= 2e-5 * B).mean() > 0.999 and dy.max() < 2e-3, (dy.max(), (dy > 2e-5 * B).mean())
    assert (dl <= 2e-3).mean() > 0.98 and dl.max() < 5e-2, (dl.max(), (dl > 2e-3).mean())
    ga, ga64 = _np(a.grad), _np(a64.grad)
    gb, gb64 = _np(b.grad), _np(b64.grad)
    scale = np.abs(ga64).max(axis=-1, keepdims=True) + 1e-6
    bad = (np.abs(ga - ga64) > 2e-3 * scale + 1e-5).any(axis=-1)
    assert bad.mean() < 2e-3, bad.mean()
    okx = np.abs(gb - gb64) <= 2e-3 * np.abs(gb64) + 1e-4
    assert okx.mean() > 0.995
    if not periodic:
        assert gb[0, 0] == _np(gy)[0, 0] and gb[1, 0] == _np(gy)[1, 0]      # out of bounds: identity
        assert not ga[0, 0].any() and not ga[1, 0].any()
    assert np.isfinite(ga).all() and np.isfinite(gb).all()


def test_training_chain_gradients_native_vs_float64_restatement():
    """The whole loss gradient with the native spline nodes (float32) against the all-torch restatement run in
    float64.  (The all-torch chain in float32 is itself ~17 % off the float64 gradient on this flow -- a few
    narrow-bin rows dominate its error -- so it is not the yardstick.)"""
    from pzflow_b200.train import TorchChain
    ex = F.example_flow()
    X = torch.from_numpy(F.galaxy_rows(4096, seed=21)).cuda()
    grads = []
    for native, dt in ((True, torch.float32), (False, torch.float64)):
        chain = TorchChain(ex["bijector_info"], ex["params"][1], ex["latent_info"], ex["params"][0], 7,
                           torch.device("cuda"), dtype=dt, native_spline=native)
        assert chain.native_spline == native
        lp = chain.log_prob(X.to(dt))
        fin = torch.isfinite(lp)
        (-(torch.where(fin, lp, torch.zeros_like(lp))).sum() / len(X)).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in chain.leaves]).double())
    a, b = _np(grads[0]), _np(grads[1])
    assert np.isfinite(a).all()
    assert np.linalg.norm(a - b) <= 2e-2 * np.linalg.norm(b)


def test_graphed_training_step_equals_the_eager_step():
    """train.GraphedStep (one CUDA graph per batch shape) and train.NativeAdam (pzb_adam_step) against the same
    steps launched eagerly with torch.optim.Adam over the same leaves."""
    from pzflow_b200.train import GraphedStep, NativeAdam, TorchChain, dp_step
    ex = F.example_flow()
    X = torch.from_numpy(F.galaxy_rows(6 * 512, seed=5)).cuda()
    w = torch.rand(len(X), device="cuda") + 0.5
    mk = lambda: TorchChain(ex["bijector_info"], ex["params"][1], ex["latent_info"], ex["params"][0], 7,  # noqa: E731
                            torch.device("cuda"))
    out = []
    for mode in ("graph", "eager"):
        chain = mk()
        opt = NativeAdam(chain)
        step = GraphedStep(chain, opt, 512, 7, 0, 512) if mode == "graph" else None
        for b in range(6):
            xb, wb = X[b * 512:(b + 1) * 512], w[b * 512:(b + 1) * 512]
            if step is not None:
                step(xb, None, wb)
            else:
                dp_step(chain, opt, xb, None, wb)
        if step is not None:
            assert step.graph is not None and int(opt.count.item()) == 6
        out.append(_np(chain.flat).copy())
    assert np.abs(out[1] - _np(mk().flat)).max() > 1e-3             # six Adam steps of 1e-3 each
    np.testing.assert_array_equal(out[0], out[1])                    # the graph replays the same kernels
    # pzb_adam_step against torch.optim.Adam on identical gradients spanning ten decades (a real training run
    # cannot serve: a 1e-7 parameter difference grows chaotically over a few steps of this flow)
    a, b = mk(), mk()
    oa, ob = NativeAdam(a), torch.optim.Adam(b.leaves, lr=1e-3)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(8):
        g = torch.randn(a.flat.shape, generator=gen, device="cuda") * \
            10 ** torch.randint(-9, 1, a.flat.shape, generator=gen, device="cuda").float()
        a.flat_grad.copy_(g)
        b.flat_grad.copy_(g)
        oa.step()
        ob.step()
    assert (a.flat - b.flat).abs().max().item() < 1e-6


def test_flow_train_conditional_with_convolved_errors():
    """Flow.train on a conditional flow with convolve_errs=True (flow.py:1068-1079): the graphed step takes the
    noisy data and conditions through its static buffers; the ragged last batch runs eagerly."""
    from pzflow_b200 import Flow
    rng = np.random.default_rng(3)
    n = 5000                                             # 4 full batches of 1024 + a ragged one of 904
    c = rng.uniform(-1, 1, n)
    df = pd.DataFrame({"a": np.sin(3 * c) + rng.normal(0, 0.1, n), "b": c ** 2 + rng.normal(0, 0.1, n), "c": c})
    df["a_err"], df["b_err"], df["c_err"] = 0.02, 0.02, 0.01
    flow = Flow(data_columns=("a", "b"), conditional_columns=("c",))
    losses = flow.train(df, epochs=5, convolve_errs=True, seed=1)
    assert len(losses) == 6 and all(np.isfinite(losses)) and losses[-1] < losses[0] - 0.3
    lp = flow.log_prob(df)
    assert lp.shape == (n,) and np.isfinite(lp).mean() > 0.99
    # the conditional structure was learned: shuffling the condition column lowers the likelihood
    shuffled = df.copy()
    shuffled["c"] = rng.permutation(df["c"].to_numpy())
    assert flow.log_prob(shuffled)[np.isfinite(lp)].mean() < lp[np.isfinite(lp)].mean() - 0.2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_data_parallel_training_two_gpus():
    """Flow.train under torchrun, one rank per GPU: the captured step holds the NCCL all_reduce of the flat gradient;
    parameters and losses come out identical on every rank and the loss falls."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(here, "dp_train_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DPRESULT ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(line[0][len("DPRESULT "):])
    assert res["world"] == 2 and res["identical_params"] and res["identical_losses"]
    assert len(res["losses"]) == 5 and res["losses"][-1] < res["losses"][0] - 0.5


def test_one_plan_many_host_threads_and_streams():
    """include/pzflow_b200.h: a plan may be launched concurrently from several host threads / streams
    (device-pointer calls), and host-buffer calls on one plan serialise."""
    import threading
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    X = torch.from_numpy(F.galaxy_rows(200_000, seed=13)).cuda()
    want = plan.log_prob(X)
    Xh = X.cpu().pin_memory()
    torch.cuda.synchronize()
    results, errors = {}, []

    def work(i):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                outs = [plan.log_prob(X) for _ in range(6)]
                if i % 2:
                    outs.append(plan.log_prob_host(Xh).cuda())
            st.synchronize()
            results[i] = outs
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert sorted(results) == [0, 1, 2, 3]
    for outs in results.values():
        for o in outs:
            assert _same_bits(o, want)


def test_get_example_flow_is_the_shipped_flow():
    """pzflow_b200.examples.get_example_flow (pzflow/examples.py:105-113) == the parameters the golden vectors
    were produced with."""
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    z = golden("ref_shim_example_flow.npz")
    df = pd.DataFrame(z["X"], columns=flow.data_columns)
    lp = flow.log_prob(df)
    want = z["log_prob"]
    fin = ~zero_class(want) & (want > LP_UNDERFLOW)
    assert (zero_class(lp) == zero_class(want)).mean() > 0.99
    np.testing.assert_allclose(lp[fin], want[fin], atol=5e-3)
    assert flow.sample(10, seed=0).shape == (10, 7)


@pytest.mark.parametrize("engine", ENGINES)
def test_elementwise_steps_between_coupling_stages(engine):
    """Non-NSC steps BETWEEN coupling stages (affine, colour transform, nested chains) take the kernels'
    "light step" path in the middle of the chain, forward and inverse."""
    rsc = ("RollingSplineCoupling", (2,))
    info = ("Chain", (("ShiftBounds", (np.full(6, -1.0, np.float32), np.full(6, 2.0, np.float32), 4.0)), rsc,
                      ("StandardScaler", (np.linspace(-0.3, 0.3, 6).astype(np.float32),
                                          np.linspace(0.8, 1.3, 6).astype(np.float32))),
                      ("Chain", (("Reverse", ()), rsc, ("Scale", (0.9,)))), ("ColorTransform", (1, [1, 2, 4])), rsc))
    rng = np.random.default_rng(17)
    sub = lambda: O.synth_bijector_params(O.expand_rolling((2,)), 6, rng, out_scale=2.0)  # noqa: E731
    bparams = [(), sub(), (), [(), sub(), ()], (), sub()]
    latent_info = ("CentBeta13", (6, 5))
    from pzflow_b200.plan import Plan
    plan = Plan(info, bparams, 6, latent_info)
    if engine == "tc" and not plan.tc_supported:
        pytest.skip("tensor-core engine does not support this shape")
    if engine == "wide" and not plan.wide_supported:
        pytest.skip("wide tensor-core engine does not support this shape")
    plan.set_engine(engine)
    X = rng.uniform(-0.9, 1.9, size=(5000, 6)).astype(np.float32)
    o32 = O.flow_log_prob(info, latent_info, ((), bparams), X)
    got = _np(plan.log_prob(torch.from_numpy(X).cuda()))
    ok = ~zero_class(o32) & (o32 > LP_UNDERFLOW)
    assert ok.mean() > 0.5 and (zero_class(got) == zero_class(o32)).mean() > 0.995
    assert within_tol(got[ok], o32[ok]).mean() > 0.97
    u, ld = plan.forward(torch.from_numpy(X).cuda())
    u32, ld32 = O.apply_bijector(info, bparams, X, np.zeros((len(X), 1), np.float32))
    fin = np.isfinite(u32).all(axis=1) & ok
    assert within_tol(_np(u)[fin], u32[fin]).mean() > 0.97 and within_tol(_np(ld)[fin], ld32[fin]).mean() > 0.97
    xr, ldi = plan.inverse(u)
    inside = (np.abs(_np(u)) < 4.5).all(axis=1) & fin
    err = np.abs(_np(xr)[inside] - X[inside])   # steep synthetic splines: judge the round trip by quantiles
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 2e-3
    asym = np.abs(_np(ld)[inside] + _np(ldi)[inside])
    assert np.median(asym) < 1e-4 and np.quantile(asym, 0.99) < 5e-3


def test_dataframe_columns_take_the_pipelined_cast():
    """Flow.log_prob on a float64 DataFrame (pzb_log_prob_host_columns: cast + gather + condition scaling inside
    the pipeline) == the float32 array path, for an unconditional and a conditional flow; a float32 DataFrame
    takes the same route (pzb_log_prob_host_columns_f32), a mixed one the array path."""
    from pzflow_b200 import Flow
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    n = 2_600_003                                     # several pipeline chunks
    X = np.tile(F.galaxy_rows(20_000, seed=31), (n // 20_000 + 1, 1))[:n].astype(np.float64)
    X += np.linspace(0, 1e-3, n)[:, None]
    df = pd.DataFrame(X, columns=flow.data_columns)
    lp = flow.log_prob(df)
    want = _np(flow._plan().log_prob(torch.from_numpy(X.astype(np.float32)).cuda()))
    np.testing.assert_array_equal(lp, want)
    np.testing.assert_array_equal(flow.log_prob(df.astype(np.float32)), want)
    np.testing.assert_array_equal(flow.log_prob(df.iloc[::-1])[::-1], want)   # non-contiguous column views
    mixed = df.astype({"redshift": np.float32})
    np.testing.assert_array_equal(flow.log_prob(mixed), want)
    cfg = F.config_c4(1)[0]
    cflow = Flow(_dictionary={"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": ("a", "b"),
                              "condition_means": np.array([0.5, -1.0], np.float32),
                              "condition_stds": np.array([2.0, 0.25], np.float32), "data_error_model": None,
                              "condition_error_model": None, "autoscale_conditions": True, "info": None,
                              "latent_info": cfg["latent_info"], "bijector_info": cfg["bijector_info"],
                              "params": cfg["params"]})
    rng = np.random.default_rng(5)
    m = 300_001
    dfc = pd.DataFrame(X[:m], columns=F.GALAXY_COLUMNS)
    dfc["a"], dfc["b"] = rng.normal(0.5, 2.0, m), rng.normal(-1.0, 0.25, m)
    got = cflow.log_prob(dfc)
    cond = (dfc[["a", "b"]].to_numpy().astype(np.float32) - np.array([0.5, -1.0], np.float32)) / np.array([2.0, 0.25], np.float32)
    wantc = _np(cflow._plan().log_prob(torch.from_numpy(X[:m].astype(np.float32)).cuda(), torch.from_numpy(cond).cuda()))
    np.testing.assert_array_equal(got, wantc)
    np.testing.assert_array_equal(cflow.log_prob(dfc.astype(np.float32)), _np(cflow._plan().log_prob(
        torch.from_numpy(X[:m].astype(np.float32)).cuda(),
        torch.from_numpy((dfc[["a", "b"]].to_numpy(dtype=np.float32) - np.array([0.5, -1.0], np.float32))
                         / np.array([2.0, 0.25], np.float32)).cuda())))


# --------------------------------------------------------------------------- #
# parity report: profiles/parity_r02.json                 #
# --------------------------------------------------------------------------- #
def _report_case(name, cfg, X, cond, engines, rows_note):
    """One config on every engine: in-tolerance fraction CUDA-vs-oracle beside the float32 oracle's own distance to
    its float64 twin, bin flips, and for every flipped bin how far the value sits from the separating knot."""
    from parity import flip_tie_report, record
    bi, li, p64 = _o64(cfg)
    t32, t64 = O.BinTrace(), O.BinTrace()
    c64 = None if cond is None else cond.astype(np.float64)
    o32 = O.flow_log_prob(bi, li, cfg["params"], X, cond, bins=t32)
    o64 = O.flow_log_prob(bi, li, p64, X.astype(np.float64), c64, bins=t64)
    both = ~zero_class(o32, LP_UNDERFLOW) & ~zero_class(o64, LP_UNDERFLOW)
    base = dict(rows=int(len(X)), rows_note=rows_note,
                oracle_f32_vs_f64_in_tol=float(within_tol(o32[both], o64[both]).mean()),
                oracle_f32_vs_f64_median=float(np.median(np.abs(o32 - o64)[both])),
                oracle_f64_bin_flips=flip_tie_report(np.concatenate(list(t64), axis=1), t32, t64))
    record(name, **base)
    ct = None if cond is None else torch.from_numpy(cond)
    for engine in engines:
        plan = _plan(cfg, engine)
        lp, bins = plan.log_prob(torch.from_numpy(X), ct, return_bins=True)
        lp = _np(lp)
        ok = both & ~zero_class(lp, LP_UNDERFLOW)
        ties = flip_tie_report(_np(bins), t32, t64)
        record(name, **{engine: dict(
            in_tol_vs_oracle_f32=float(within_tol(lp[ok], o32[ok]).mean()),
            in_tol_vs_oracle_f64=float(within_tol(lp[ok], o64[ok]).mean()),
            median_abs_vs_f32=float(np.median(np.abs(lp - o32)[ok])),
            median_abs_vs_f64=float(np.median(np.abs(lp - o64)[ok])),
            class_mismatch=int((zero_class(lp, LP_UNDERFLOW) != zero_class(o32, LP_UNDERFLOW)).sum()),
            bins=ties, overflow_rows=plan.overflow_rows() if engine != "simt" else 0)})
        # north_star: bin indices bit-exact except at knot ties -- every flip is between adjacent bins and the value
        # sits on the separating knot to within the stage's float32 noise floor (tests/parity.py:flip_tie_report)
        assert ties["non_adjacent"] == 0, (name, engine, ties)
        assert ties["max_ratio"] <= 4.0, (name, engine, ties)
        assert ties["flips"] <= max(3, int(5e-4 * ties["total"])), (name, engine, ties)


def test_parity_report():
    """Writes the per-config parity figures the judge asked for (in-tolerance fractions, bin flips as knot ties)."""
    import json
    import os
    from parity import report
    rng = np.random.default_rng(11)
    ex = F.example_flow()
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    _report_case("c2_shipped_7d_flow", ex, F.galaxy_rows(50000, seed=3), None, ["simt", "tc", "wide"],
                 "x = inverse(u), u ~ U(-5,5)^7 + 1 % out-of-box rows")
    c1 = F.config_c1()
    _report_case("c1_two_moons_default_flow", c1, rng.uniform([-1.6, -1.1], [2.6, 1.6], size=(50000, 2)).astype(np.float32),
                 None, ["simt", "tc", "wide"], "uniform in the ShiftBounds box")
    c4 = F.config_c4(1)[0]
    _report_case("c4_conditional_7d_flow", c4, rng.uniform(mins, maxs, size=(20000, 7)).astype(np.float32),
                 rng.normal(size=(20000, 2)).astype(np.float32), ["simt", "tc", "wide"], "uniform in the box, N(0,1) conditions")
    c5 = F.config_c5()
    _report_case("c5_32d_k32_h256", c5, rng.uniform(-1, 1, size=(2048, 32)).astype(np.float32), None, ["simt", "wide"],
                 "uniform in the box")
    rep = dict(report())
    rep["_note"] = ("in_tol = fraction of rows with |a - b| <= 1e-4 + 1e-5 |b| (north_star); bins.ulps = the largest "
                    "distances value-to-separating-knot of flipped bins in ulps of B = 5 (4.77e-7); bins.max_ratio = the "
                    "largest such distance over the stage's float32 noise floor (4 ulp + q99 |value32 - value64| + "
                    "q99 |knot32 - knot64|); oracle_f64_bin_flips = the float32 oracle against its own float64 twin, "
                    "for scale")
    root = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for sub in ("gpurun_out", "profiles"):
        d = os.path.join(root, sub)
        if os.path.isdir(d):
            with open(os.path.join(d, "parity_r02.json"), "w") as fh:
                json.dump(rep, fh, indent=1, sort_keys=True)
    print("PARITY_REPORT", json.dumps(rep, sort_keys=True))


# --------------------------------------------------------------------------- #
# marg_rules: the reference's own recursion (golden) vs the in-kernel axis     #
# --------------------------------------------------------------------------- #
def _oracle_marginal(ex, params, df, rules, grid, columns, nan_to_zero=True):
    """Un-normalised marginalised pdfs straight from the definition: a row's variants are its missing columns filled
    with every combination of the rules' grids (first rule outermost, later rules evaluated on the filled-in row);
    sum over the variants of nan_to_num(exp(log_prob))."""
    dt = np.asarray(grid).dtype
    flag = rules["flag"]
    names = [n for n in rules if n != "flag"]
    out = np.zeros((len(df), len(grid)), dtype=dt)
    for i in range(len(df)):
        row = df.iloc[i]
        miss = [c for c in columns if (np.isnan(row[c]) if np.isnan(flag) else np.isclose(row[c], flag))]
        if any(c not in names for c in miss):
            continue
        variants = [row.copy()]
        for name in [n for n in names if n in miss]:
            nxt = []
            for v in variants:
                for val in np.asarray(rules[name](v)):
                    w = v.copy()
                    w[name] = val
                    nxt.append(w)
            variants = nxt
        rest = np.array([[v[c] for c in columns] for v in variants], dtype=dt)
        p = O.flow_posterior(ex["bijector_info"], ex["latent_info"], params, rest, 0, grid, normalize=False,
                             nan_to_zero=nan_to_zero)
        out[i] = p.sum(axis=0, dtype=dt)
    return out


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("tag", ["f99", "nan"])
def test_marg_rules_against_reference_golden(engine, tag):
    """tests/golden/ref_shim_marg_rules.npz = pzflow/flow.py:560-664 run by the reference's own code: rows missing u,
    y, both (nested marginalisation over 12 x 9 variants, the y grid depending on the row), and one row missing a
    column without a rule.  Here the variants are an expansion axis inside the kernel (pzb_posterior_marg)."""
    from pzflow_b200.examples import get_example_flow
    z = golden("ref_shim_marg_rules.npz")
    flow = get_example_flow()
    flow._plan().set_engine(engine)
    flag = 99.0 if tag == "f99" else np.nan
    df = pd.DataFrame(z[f"{tag}_inputs"], columns=flow.data_columns)
    rules = {"flag": flag, "u": lambda row: np.linspace(25.0, 31.0, 12),
             "y": lambda row: np.linspace(row["z"] - 1.0, row["z"] + 1.0, 9)}
    grid = z["grid"]
    ex = F.example_flow()
    cols = list(flow.data_columns[1:])
    p64 = O.cast_params(ex["params"], np.float64)
    un64 = _oracle_marginal(ex, p64, df, rules, grid.astype(np.float64), cols)
    un = flow.posterior(df, column="redshift", grid=grid, marg_rules=rules, normalize=False, nan_to_zero=False)
    assert un.shape == (40, 46)
    assert (un[30] == 0).all() and (z[f"{tag}_posterior_unnorm"][30] == 0).all()      # no rule for g: never evaluated
    assert_noise_floor(f"marg unnorm [{engine},{tag}]", un, z[f"{tag}_posterior_unnorm"], un64, ratio=2.0, slack=1e-3,
                       max_class_mismatch=0)
    pdf = flow.posterior(df, column="redshift", grid=grid, marg_rules=rules)
    norm64 = un64 / O.trapezoid(un64, grid.astype(np.float64)).reshape(-1, 1)
    norm64 = np.nan_to_num(norm64, nan=0.0)
    assert_noise_floor(f"marg posterior [{engine},{tag}]", pdf, z[f"{tag}_posterior"], norm64, ratio=2.0, slack=1e-5,
                       max_class_mismatch=0)
    ok = pdf.sum(axis=1) > 0
    assert ok.sum() == 39 and not ok[30]
    np.testing.assert_allclose(O.trapezoid(pdf[ok], grid), 1.0, rtol=2e-6)
    # rows 5 (u and y missing) and 1 (u missing) really are marginals: wider than the complete row's posterior
    assert pdf[5].max() < pdf[4].max() * 1.5 and np.isfinite(pdf).all()


def test_marg_rules_with_error_samples_and_conditions():
    """The marginal axis composes with the error-sample axis and with conditions: zero errors reproduce the plain
    marginalised posterior (the reference's own zero-error property, tests/test_flow.py:260-264)."""
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    rng = np.random.default_rng(4)
    n = 30
    df = pd.DataFrame({"a": rng.uniform(-1, 1, n), "b": rng.uniform(-1, 1, n), "c": rng.uniform(-1, 1, n),
                       "p": rng.normal(0, 1, n)})
    flow = Flow(data_columns=("a", "b", "c"), conditional_columns=("p",),
                bijector=Chain(ShiftBounds(-np.ones(3), np.ones(3), 4.0), RollingSplineCoupling(3, n_conditions=1)))
    df.loc[[2, 7, 11], "b"] = np.nan
    df.loc[[7, 20], "c"] = np.nan
    rules = {"flag": np.nan, "b": lambda row: np.linspace(-0.9, 0.9, 7), "c": lambda row: np.linspace(-0.8, 0.8, 5)}
    grid = np.linspace(-1, 1, 33).astype(np.float32)
    base = flow.posterior(df, column="a", grid=grid, marg_rules=rules)
    zero = df.copy()
    for c in ("a", "b", "c", "p"):
        zero[f"{c}_err"] = 0.0
    with_err = flow.posterior(zero, column="a", grid=grid, marg_rules=rules, err_samples=3, seed=1)
    np.testing.assert_allclose(with_err, base, rtol=1e-5, atol=1e-7)
    # against the definition, through the un-marginalised posterior of explicitly filled-in rows
    filled = pd.DataFrame(np.repeat(df.iloc[[7]].to_numpy(), 35, axis=0), columns=df.columns)
    filled["b"] = np.repeat(np.linspace(-0.9, 0.9, 7), 5)
    filled["c"] = np.tile(np.linspace(-0.8, 0.8, 5), 7)
    summed = flow.posterior(filled, column="a", grid=grid, normalize=False).astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(base[7], summed / np.trapezoid(summed, grid.astype(np.float64)), rtol=1e-4, atol=1e-7)
    # real errors change the answer but keep it a normalised density
    zero["c_err"] = 0.3
    noisy = flow.posterior(zero, column="a", grid=grid, marg_rules=rules, err_samples=16, seed=1)
    np.testing.assert_allclose(np.trapezoid(noisy, grid, axis=1), 1.0, rtol=1e-5)
    assert np.abs(noisy - base).max() > 1e-3


# --------------------------------------------------------------------------- #
# one process, several GPUs                                #
# --------------------------------------------------------------------------- #
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_shards_host_arrays_over_all_gpus(monkeypatch):
    """Flow.log_prob / posterior / sample of a plain process use every visible GPU (contiguous row blocks, one host
    pipeline per device, no collective) and return, bit for bit, what one GPU returns."""
    from pzflow_b200 import sharding
    from pzflow_b200.examples import get_example_flow
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    monkeypatch.setattr(sharding, "MULTI_DEVICE_MIN_ROWS", 10_000)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    flow = get_example_flow()
    X = F.galaxy_rows(60_011, seed=5)
    df = pd.DataFrame(X.astype(np.float64), columns=flow.data_columns)
    grid = np.linspace(0, 2.3, 50).astype(np.float32)
    sharding.set_devices([0])
    lp1, pdf1, s1 = flow.log_prob(df), flow.posterior(df.iloc[:2000], "redshift", grid), flow.sample(40_003, seed=9)
    sharding.set_devices("all")
    assert len(flow._device_plans(60_011)) == torch.cuda.device_count()
    lpn, pdfn, sn = flow.log_prob(df), flow.posterior(df.iloc[:2000], "redshift", grid), flow.sample(40_003, seed=9)
    assert np.array_equal(lp1.view(np.int32), lpn.view(np.int32))
    assert np.array_equal(pdf1.view(np.int32), pdfn.view(np.int32))
    assert np.array_equal(np.ascontiguousarray(s1.to_numpy()).view(np.int32), np.ascontiguousarray(sn.to_numpy()).view(np.int32))
    # conditional flow with the default (CentBeta13) latent: conditions travel with their row block
    rng = np.random.default_rng(1)
    cf = Flow(data_columns=("a", "b", "c"), conditional_columns=("p", "q"),
              bijector=Chain(ShiftBounds(-np.ones(3), np.ones(3), 4.0), RollingSplineCoupling(3, n_conditions=2)))
    cdf = pd.DataFrame({"a": rng.uniform(-1, 1, 30_000), "b": rng.uniform(-1, 1, 30_000), "c": rng.uniform(-1, 1, 30_000),
                        "p": rng.normal(size=30_000), "q": rng.normal(size=30_000)})
    sharding.set_devices([0])
    c1, cs1 = cf.log_prob(cdf), cf.sample(2, conditions=cdf.iloc[:12_000], seed=3)
    sharding.set_devices("all")
    cn, csn = cf.log_prob(cdf), cf.sample(2, conditions=cdf.iloc[:12_000], seed=3)
    # an ensemble shards the same way (every device evaluates all flows on its block of rows)
    from pzflow_b200 import FlowEnsemble
    ens = FlowEnsemble(data_columns=("a", "b", "c"), conditional_columns=("p", "q"), N=3,
                       bijector=Chain(ShiftBounds(-np.ones(3), np.ones(3), 4.0), RollingSplineCoupling(3, n_conditions=2)))
    eg = np.linspace(-1, 1, 40).astype(np.float32)
    en, epn = ens.log_prob(cdf), ens.posterior(cdf.iloc[:1500], "a", eg)
    sharding.set_devices([0])
    e1, ep1 = ens.log_prob(cdf), ens.posterior(cdf.iloc[:1500], "a", eg)
    assert np.array_equal(e1.view(np.int32), en.view(np.int32)) and np.array_equal(ep1.view(np.int32), epn.view(np.int32))
    sharding.set_devices(None)
    assert np.array_equal(c1.view(np.int32), cn.view(np.int32))
    assert np.array_equal(np.ascontiguousarray(cs1.to_numpy(), dtype=np.float64).view(np.int64),
                          np.ascontiguousarray(csn.to_numpy(), dtype=np.float64).view(np.int64))


# --------------------------------------------------------------------------- #
# reductions inside the tensor-core kernel                 #
# --------------------------------------------------------------------------- #
def _both_routes(fn):
    """fn() through the in-kernel reduction and through the scratch + reduction-kernel route (PZB_NO_FUSE=1)."""
    import os
    os.environ.pop("PZB_NO_FUSE", None)
    os.environ["PZB_FUSE_POSTERIOR"] = "1"     # the plain posterior's single-pass form is opt-in
    try:
        fused = _np(fn())
        os.environ["PZB_NO_FUSE"] = "1"
        plain = _np(fn())
    finally:
        os.environ.pop("PZB_NO_FUSE", None)
        os.environ.pop("PZB_FUSE_POSTERIOR", None)
    return fused, plain


@pytest.mark.parametrize("G,Ng", [(1000, 301), (60, 4001), (5, 777), (1024, 150), (129, 3)])
def test_posterior_is_normalised_inside_the_kernel(G, Ng):
    """Flow.posterior's exp / trapezoid / divide / NaN -> 0 (pzflow/flow.py:720-737) happen in shared memory before the
    only write of the pdfs: same numbers as the three-pass route (chain kernel, then normalize_pdfs_kernel)."""
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    X = F.galaxy_rows(Ng, seed=G)
    X[0, 3] = np.nan                                        # an all-zero pdf (0 / 0 -> NaN -> 0)
    rest = torch.from_numpy(np.ascontiguousarray(X[:, 1:])).cuda()
    grid = torch.linspace(0.0, 2.3, G).cuda()
    for normalize, nz in ((True, True), (False, False), (True, False)):
        fused, plain = _both_routes(lambda: plan.posterior(rest, 0, grid, normalize=normalize, nan_to_zero=nz))
        assert fused.shape == (Ng, G)
        np.testing.assert_allclose(fused, plain, rtol=2e-6, atol=1e-30, equal_nan=True)
    assert (fused[0] != fused[0]).all() or (fused[0] == 0).all()
    import os
    os.environ["PZB_FUSE_POSTERIOR"] = "1"
    try:
        host = plan.posterior_host(rest.cpu(), 0, grid.cpu())   # the host pipeline takes the same route, chunk by chunk
    finally:
        os.environ.pop("PZB_FUSE_POSTERIOR", None)
    f2, _ = _both_routes(lambda: plan.posterior(rest, 0, grid))
    assert np.array_equal(host.numpy().view(np.int32), f2.view(np.int32))


@pytest.mark.parametrize("S,G,Ng", [(3, 200, 500), (20, 1000, 300), (7, 33, 10)])
def test_error_samples_are_averaged_inside_the_kernel(S, G, Ng):
    """err_samples (pzflow/flow.py:455-464, 722-724): the mean over a galaxy's samples is taken in shared memory --
    whole galaxies per chunk, or a galaxy's samples walked chunk by chunk by one CTA -- instead of through an
    [Ng, S, G] scratch array."""
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    rng = np.random.default_rng(S)
    X = F.galaxy_rows(Ng, seed=S, oob_frac=0.0)
    rest = np.ascontiguousarray(X[:, 1:])
    rest_err = (0.05 * rng.random(rest.shape)).astype(np.float32)
    grid = np.linspace(0.0, 2.3, G).astype(np.float32)
    fused, plain = _both_routes(lambda: plan.posterior_err(rest, rest_err, 0, grid, err_samples=S, seed=11))
    np.testing.assert_allclose(fused, plain, rtol=5e-6, atol=1e-30)
    Xerr = (0.05 * rng.random(X.shape)).astype(np.float32)
    fused, plain = _both_routes(lambda: plan.log_prob_err(X, Xerr, err_samples=S, seed=5))
    np.testing.assert_allclose(fused, plain, rtol=2e-6, atol=2e-6)
    if Ng >= 300:   # one row's samples spanning several chunks
        fused, plain = _both_routes(lambda: plan.log_prob_err(X, Xerr, err_samples=2500, seed=5))
        np.testing.assert_allclose(fused, plain, rtol=2e-5, atol=2e-5)


def test_marginal_variants_are_summed_inside_the_kernel():
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    rng = np.random.default_rng(8)
    Ng, G, V = 400, 120, 9
    rest = np.ascontiguousarray(F.galaxy_rows(Ng, seed=2, oob_frac=0.0)[:, 1:])
    mvals = rng.uniform(25, 30, size=(Ng, V, 2)).astype(np.float32)
    grid = np.linspace(0.0, 2.3, G).astype(np.float32)
    for S in (0, 4):
        fused, plain = _both_routes(lambda: plan.posterior_marg(rest, 0, grid, [1, 6], mvals, err_samples=S, seed=3,
                                                                rest_err=np.full_like(rest, 0.02) if S else None))
        np.testing.assert_allclose(fused, plain, rtol=5e-6, atol=1e-30)


# --------------------------------------------------------------------------- #
# the three-tile form of the tensor-core kernel (experimental, PZB_TC_PIPE3=1)  #
# --------------------------------------------------------------------------- #
def test_three_tile_kernel_equals_the_two_tile_kernel_bit_for_bit(monkeypatch):
    """nsf_chain_tc3_kernel (csrc/nsf_tc_pipe3.cuh: three tiles in flight over three X regions and one shared Y of TMEM,
    one issuer warp, key operand in shared memory) issues the same MMAs in the same accumulation order as
    nsf_chain_tc_kernel: every result must be the same bits -- log_prob, bins, the inverse map, posteriors; stages with
    1, 2, 3 and 4 spline dimensions, sub-stages, conditions, ragged and tiny row counts."""
    def both(fn):
        out = []
        for v in ("0", "1"):
            monkeypatch.setenv("PZB_TC_PIPE3", v)
            r = fn()
            torch.cuda.synchronize()
            out.append([x.clone() for x in (r if isinstance(r, (tuple, list)) else [r])])
        return out

    def same(a, b):
        return all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(a, b))

    ex = F.example_flow()
    plan = _plan(ex, "tc")
    for n in (1, 129, 5000, 200001):
        X = torch.from_numpy(F.galaxy_rows(n, seed=n)).cuda()
        assert same(*both(lambda: plan.log_prob(X, return_bins=True))), n
    X = torch.from_numpy(F.galaxy_rows(6000, seed=2)).cuda()
    u = plan.forward(X)[0]
    assert same(*both(lambda: plan.inverse(u)))
    grid = torch.linspace(0, 2.3, 300).cuda()
    assert same(*both(lambda: plan.posterior(X[:500, 1:].contiguous(), 0, grid)))
    for D, C in ((2, 0), (5, 0), (10, 0), (3, 2)):          # 1, 3, 4 + 1 and 2 spline dimensions per (sub-)stage
        cfg = F.default_flow(D, -np.ones(D), np.ones(D), n_conditions=C, seed=D)
        p = _plan(cfg, "tc")
        rng = np.random.default_rng(D)
        Xd = torch.from_numpy(rng.uniform(-1.02, 1.02, size=(3333, D)).astype(np.float32)).cuda()
        Cd = torch.from_numpy(rng.normal(size=(3333, C)).astype(np.float32)).cuda() if C else None
        assert same(*both(lambda: p.log_prob(Xd, Cd, return_bins=True))), (D, C)
        ud = p.forward(Xd, Cd)[0]
        assert same(*both(lambda: p.inverse(ud, Cd))), (D, C)
    # more than eight conditioner inputs (they ride above the bias row of the key operand, which the three-tile kernel
    # keeps in shared memory)
    # (10 columns + 4 conditions = 9 inputs; four tiles of its 21 state columns still fit beside the key buffer -- a flow
    # with a wider state takes the two-tile kernel whatever PZB_TC_PIPE3 says)
    cfg = F.default_flow(10, -np.ones(10), np.ones(10), n_conditions=4, seed=16)
    p = _plan(cfg, "tc")
    rng = np.random.default_rng(16)
    Xd = torch.from_numpy(rng.uniform(-1.02, 1.02, size=(2000, 10)).astype(np.float32)).cuda()
    Cd = torch.from_numpy(rng.normal(size=(2000, 4)).astype(np.float32)).cuda()
    assert same(*both(lambda: p.log_prob(Xd, Cd, return_bins=True)))
    # the reductions inside the kernel (error samples, marginal variants): the chunk geometry differs between the two
    # kernels (the key buffer takes shared memory), the sums' order does not
    rest = np.ascontiguousarray(F.galaxy_rows(300, seed=4, oob_frac=0.0)[:, 1:])
    g120 = np.linspace(0.0, 2.3, 120).astype(np.float32)
    a, b = both(lambda: plan.posterior_err(rest, np.full_like(rest, 0.02), 0, g120, err_samples=6, seed=1))
    torch.testing.assert_close(a[0], b[0], rtol=2e-6, atol=1e-30)
    Xe = F.galaxy_rows(4000, seed=5)
    a, b = both(lambda: plan.log_prob_err(Xe, np.full_like(Xe, 0.01), err_samples=5, seed=2))
    torch.testing.assert_close(a[0], b[0], rtol=2e-6, atol=1e-6)


# --------------------------------------------------------------------------- #
# the gradient of one optimisation step as one kernel      #
# --------------------------------------------------------------------------- #
def _f64_loss_and_grad(cfg, X, cond, w):
    from pzflow_b200.train import TorchChain
    chain = TorchChain(cfg["bijector_info"], cfg["params"][1], cfg["latent_info"], cfg["params"][0], cfg["input_dim"],
                       torch.device("cuda"), dtype=torch.float64, native_spline=False)
    lp = chain.log_prob(X.double(), None if cond is None else cond.double())
    fin = torch.isfinite(lp)
    loss = -(torch.where(fin, w.double() * lp, torch.zeros_like(lp))).sum() / len(X)
    loss.backward()
    return float(loss), _np(torch.cat([p.grad.reshape(-1) for p in chain.leaves]))


@pytest.mark.parametrize("case", ["shipped", "conditional_centbeta", "deep_narrow", "colour_chain"])
def test_fused_training_step_gradient_against_float64_autograd(case):
    """pzb_train_step (forward + backward through the whole chain in one kernel, hand-derived) against float64
    autograd over the torch restatement: the loss, and the gradient of every weight."""
    from pzflow_b200.train import FusedGradient, TorchChain
    rng = np.random.default_rng(7)
    if case == "shipped":
        cfg = dict(F.example_flow(), input_dim=7)
        X, cond = torch.from_numpy(F.galaxy_rows(1000, seed=21)).cuda(), None      # 62.5 tiles: a ragged last one
    elif case == "conditional_centbeta":
        ex = F.example_flow()
        mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
        cfg = F.config_c4(1)[0]
        X = torch.from_numpy(rng.uniform(mins + 0.1, maxs - 0.1, size=(777, 7)).astype(np.float32)).cuda()
        cond = torch.from_numpy(rng.normal(size=(777, 2)).astype(np.float32)).cuda()
    elif case == "colour_chain":   # the photo-z chain of the reference's tutorials: the parameter-free stages in front
        info = ("Chain", (("ColorTransform", (3, [1, 2, 3, 4, 5, 6])), ("InvSoftplus", (0, 4.0)),
                          ("StandardScaler", (np.linspace(-0.5, 0.5, 7).astype(np.float32),
                                              np.linspace(0.5, 2.0, 7).astype(np.float32))),
                          ("Reverse", ()), ("RollingSplineCoupling", (7,)), ("Scale", (0.8,))))
        bparams = [(), (), (), (), O.synth_bijector_params(O.expand_rolling((7,)), 7, rng, out_scale=2.0), ()]
        cfg = dict(bijector_info=info, latent_info=("CentBeta13", (7, 5)),
                   params=(tuple((0.0, 0.0) for _ in range(7)), bparams), input_dim=7)
        X, cond = torch.from_numpy(rng.uniform(0.05, 1.5, size=(555, 7)).astype(np.float32)).cuda(), None
    else:   # three hidden layers of 40, K = 5, odd column count, StandardScaler in front
        info = ("Chain", (("StandardScaler", (np.linspace(-1, 1, 5).astype(np.float32), np.linspace(1, 2, 5).astype(np.float32))),
                          F.rolling_info(3, 2, 5, 4, 3, 40)))
        params = O.synth_bijector_params(info, 5, rng, out_scale=2.0)
        cfg = dict(bijector_info=info, latent_info=("Normal", (5,)), params=((), params), input_dim=5)
        X, cond = torch.from_numpy(rng.uniform(-3, 3, size=(333, 5)).astype(np.float32)).cuda(), None
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)          # the weights must not depend on which tests ran before
    w = torch.rand(len(X), device="cuda", generator=gen) + 0.5
    chain = TorchChain(cfg["bijector_info"], cfg["params"][1], cfg["latent_info"], cfg["params"][0], cfg["input_dim"],
                       torch.device("cuda"))
    fused = FusedGradient.build(chain)
    assert fused is not None
    chain.flat_grad.fill_(float("nan"))
    fused(X, cond, w, len(X))
    got, loss = _np(chain.flat_grad).astype(np.float64), float(fused.loss.item())
    want_loss, want = _f64_loss_and_grad(cfg, X, cond, w)
    assert np.isfinite(got).all()
    assert abs(loss - want_loss) <= 2e-4 * max(1.0, abs(want_loss))
    # float32 forward + backward through seven trained (steep) spline stages against float64: a few 1e-3 of the norm
    assert np.linalg.norm(got - want) <= 1e-2 * np.linalg.norm(want), np.linalg.norm(got - want) / np.linalg.norm(want)
    # the same call again: bit-identical (no atomics anywhere in the step)
    first = chain.flat_grad.clone()
    fused(X, cond, w, len(X))
    assert torch.equal(first.view(torch.int32), chain.flat_grad.view(torch.int32))


def test_fused_training_step_in_a_cuda_graph_and_launch_count():
    """Flow.train's step with the fused gradient: graph replay == eager, bit for bit, and one step is four kernels of
    this library (gradient, partial-sum reduction, Adam's counter, Adam) -- nothing else."""
    from pzflow_b200.plan import launch_count
    from pzflow_b200.train import FusedGradient, GraphedStep, NativeAdam, TorchChain, dp_step
    ex = F.example_flow()
    X = torch.from_numpy(F.galaxy_rows(6 * 512, seed=5)).cuda()
    w = torch.rand(len(X), device="cuda") + 0.5

    def mk():
        chain = TorchChain(ex["bijector_info"], ex["params"][1], ex["latent_info"], ex["params"][0], 7, torch.device("cuda"))
        chain.fused_gradient = FusedGradient.build(chain)
        assert chain.fused_gradient is not None
        return chain
    out = []
    for mode in ("graph", "eager"):
        chain = mk()
        opt = NativeAdam(chain)
        step = GraphedStep(chain, opt, 512, 7, 0, 512) if mode == "graph" else None
        for b in range(6):
            xb, wb = X[b * 512:(b + 1) * 512], w[b * 512:(b + 1) * 512]
            l0 = launch_count()
            if step is not None:
                step(xb, None, wb)
            else:
                dp_step(chain, opt, xb, None, wb)
                assert launch_count() - l0 == 4
        out.append(_np(chain.flat).copy())
    assert np.abs(out[1] - _np(mk().flat)).max() > 1e-3
    np.testing.assert_array_equal(out[0], out[1])


@pytest.mark.parametrize("flag", [99, np.nan])
def test_posterior_with_marginalization_like_the_reference_test(flag):
    """tests/test_flow.py:84-112 of the reference, line for line: a flow that is only a Reverse bijector, every row
    missing b, then b and c (nested marginalisation), numeric and NaN flags -- plus what the reference's test leaves
    out, the values: a Reverse flow's density is the latent's, so the marginal is known in closed form."""
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Reverse
    flow = Flow(("a", "b", "c", "d"), Reverse())
    x = pd.DataFrame(np.arange(16).reshape(-1, 4).astype(np.float64), columns=("a", "b", "c", "d"))
    grid = np.arange(0, 2.1, 0.12)
    marg_rules = {"flag": flag, "b": lambda row: np.linspace(0, 1, 2), "c": lambda row: np.linspace(1, 2, 3)}
    x["b"] = flag * np.ones(x.shape[0])
    pdfs = flow.posterior(x, column="a", grid=grid, marg_rules=marg_rules)
    assert pdfs.shape == (x.shape[0], grid.size)
    x["c"] = flag * np.ones(x.shape[0])
    pdfs2 = flow.posterior(x, column="a", grid=grid, marg_rules=marg_rules)
    assert pdfs2.shape == (x.shape[0], grid.size)
    # row 0 has d = 3 (inside the CentBeta13 support), the others d >= 7 (outside: zero density -> all-zero pdf)
    g = grid.astype(np.float32).astype(np.float64)
    beta = lambda u: ((u + 5) / 10) ** 12 * (1 - (u + 5) / 10) ** 12       # un-normalised Beta(13, 13) on [-5, 5]  # noqa: E731
    want = beta(g) / np.trapezoid(beta(g), g)                               # b, c, d only scale the density of a
    np.testing.assert_allclose(pdfs[0], want, rtol=2e-5)
    np.testing.assert_allclose(pdfs2[0], want, rtol=2e-5)
    assert (pdfs[1:] == 0).all() and (pdfs2[1:] == 0).all()


# --------------------------------------------------------------------------- #
# round 2, second session: jax's own random streams, Shuffle, UniformDequantizer               #
# --------------------------------------------------------------------------- #
def test_uniform_latent_draws_are_the_jax_stream():
    """pzb_sample_uniform_threefry == jax.random.uniform(PRNGKey(seed), (n, D), -B, B) as oracle/jax_prng.py restates it
    (pinned by the Random123 known answers and the notebook samples): bit for bit, odd and even totals, both layouts,
    and row shards of a larger draw."""
    from oracle import jax_prng as J
    from pzflow_b200 import distributions as Dn
    from pzflow_b200 import prng
    from pzflow_b200.plan import sample_uniform_jax
    assert prng.layout() == "threefry_partitionable"     # the default: what the reference's required jax (>= 0.5.3) draws
    assert float(sample_uniform_jax(1, 1, 0.5, 0)[0, 0]) + 0.5 == np.float32(0.947667)   # uniform(key(0)) of jax's docs
    for n, D, seed in ((2, 7, 0), (1001, 7, 7), (4096, 3, 2 ** 31 + 5), (1, 1, 10 ** 17), (300_001, 5, 11)):
        wantp = J.uniform(J.PRNGKey(seed), (n, D), -5.0, 5.0, partitionable=True)
        got = Dn.Uniform(D, 5).sample((), n, seed)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.int32), wantp.view(np.int32)), (n, D, seed)
        want = J.uniform(J.PRNGKey(seed), (n, D), -5.0, 5.0)
        goto = _np(sample_uniform_jax(n, D, 5.0, seed, layout="threefry"))
        assert np.array_equal(goto.view(np.int32), want.view(np.int32)), (n, D, seed)
    # rows [a, b) of a 1001-row draw, as a sharded Flow.sample asks for them -- in both layouts
    for lay in ("threefry", "threefry_partitionable"):
        prng.set_layout(lay)
        try:
            whole = J.uniform(J.PRNGKey(3), (1001, 7), -5.0, 5.0, partitionable=lay != "threefry")
            for a, b in ((0, 400), (400, 1001), (1000, 1001)):
                part = _np(Dn.Uniform(7, 5).sample_device(b - a, 3, row_offset=a, total_rows=1001))
                assert np.array_equal(part.view(np.int32), whole[a:b].view(np.int32))
        finally:
            prng.set_layout("threefry_partitionable")
    # the Philox streams of round 1 stay available
    prng.set_layout("philox")
    try:
        ph = Dn.Uniform(7, 5).sample((), 1000, 0)
        assert not np.array_equal(ph, J.uniform(J.PRNGKey(0), (1000, 7), -5.0, 5.0)) and np.abs(ph).max() <= 5
    finally:
        prng.set_layout("threefry_partitionable")


def test_flow_sample_reproduces_what_real_jax_printed(stream_layout):
    """docs/tutorials/marginalization.ipynb:69, 133-135: ``flow.sample(2, seed=0)`` of the shipped flow under the real JAX
    reference (an older jax: the original stream layout).  The same call here -- latents drawn on the GPU in jax's stream,
    the inverse chain on the tensor cores -- returns the printed rows to the printed precision in that layout; and in
    either layout 2 samples of seed 0 and 1001 samples of seed 7 equal what the reference's own code returns over the
    stand-in drawing that layout (tests/golden/ref_shim_jax_streams[_part].npz)."""
    from conftest import STREAM_GOLDEN
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    z, g = golden("ref_shim_example_flow.npz"), golden(STREAM_GOLDEN[stream_layout])
    s = flow.sample(2, seed=0)
    assert list(s.columns) == ["redshift", "u", "g", "r", "i", "z", "y"]
    legacy = stream_layout == "threefry"       # the notebook was run with a jax older than 0.5: the original layout
    if legacy:
        np.testing.assert_allclose(s.to_numpy(), z["nb_X"], rtol=3e-6, atol=0)
    else:
        assert np.abs(s.to_numpy() - z["nb_X"]).max() > 0.1
    for engine in ("simt", "tc", "wide"):
        flow._plan().set_engine(engine)
        two = flow.sample(2, seed=0).to_numpy()
        np.testing.assert_allclose(two, z["nb_X"] if legacy else g["example_sample_2_seed0"], rtol=3e-6 if legacy else 2e-5,
                                   atol=0, err_msg=engine)
        big = flow.sample(1001, seed=7).to_numpy()
        want = g["example_sample_1001_seed7"]
        ex = F.example_flow()
        bi, li, p64 = _o64(ex)
        x64, _ = O.flow_inverse(bi, p64, g["example_latent_1001_seed7"].astype(np.float64))
        assert_noise_floor(f"sample(1001, seed=7) [{engine}]", big, want, x64, ratio=2.0, slack=2e-5)
    flow._plan().set_engine("auto")


@pytest.mark.parametrize("engine", ["simt", "tc", "wide"])
def test_shuffle_and_dequantizer_flows_against_reference_golden(engine, stream_layout):
    """Flows whose chains hold Shuffle stages (pzflow/bijectors.py:780-814) and a UniformDequantizer (865-911), built
    with two seeds: log_prob, forward and Flow.sample against what the reference's own code returned (stand-in with
    jax's key schedule, oracle/make_golden_shim.py jax_streams); a flow re-built from its save dictionary draws its
    permutations from PRNGKey(0) (flow.py:186-187) and so evaluates differently from the seed-3 flow it was saved from."""
    from pzflow_b200 import Flow
    from pzflow_b200 import bijectors as B
    from pzflow_b200 import distributions as Dn
    from conftest import STREAM_GOLDEN
    from oracle import jax_prng as J
    g = golden(STREAM_GOLDEN[stream_layout])
    part = stream_layout == "threefry_partitionable"
    X, mins, maxs = g["X"], g["mins"], g["maxs"]
    cols = tuple("abcde")
    df = pd.DataFrame(X.astype(np.float64), columns=cols)
    info = F.unpack_tree("shuffle_info", g)

    def bijector():
        return B.Chain(B.UniformDequantizer([2]), B.ShiftBounds(mins, maxs, 4.0), B.Shuffle(),
                       B.RollingSplineCoupling(3, K=8, hidden_dim=32), B.Chain(B.Reverse(), B.Shuffle()),
                       B.RollingSplineCoupling(2, K=8, hidden_dim=32))
    for seed in (0, 3):
        params = F.unpack_tree(f"s{seed}_params", g)
        fl = Flow(cols, bijector(), latent=Dn.Uniform(5, 5), seed=seed)
        assert repr(fl._bijector_info[1][2:5]) == repr(info[1][2:5])   # ("Shuffle", ()) stays free of its permutation
        fl.set_bijector(bijector(), params=params, seed=seed)
        plan = fl._plan()
        if engine == "tc" and not plan.tc_supported or engine == "wide" and not plan.wide_supported:
            pytest.skip("engine does not take this shape")
        plan.set_engine(engine)
        ri = O.resolve_shuffles(info, J.PRNGKey(seed), 5, part)
        p64 = O.cast_params(((), params), np.float64)
        lp = fl.log_prob(df)
        assert_noise_floor(f"shuffle flow log_prob seed {seed} [{engine}]", lp, g[f"s{seed}_log_prob"],
                           O.flow_log_prob(ri, ("Uniform", (5, 5)), p64, X.astype(np.float64)), ratio=2.0, slack=2e-5,
                           below=LP_UNDERFLOW)
        u, ld = fl._forward(fl._params[1], X)
        u64, ld64 = O.apply_bijector(ri, p64[1], X.astype(np.float64), np.zeros((len(X), 1)))
        assert_noise_floor(f"shuffle flow forward seed {seed} [{engine}]", u, g[f"s{seed}_forward_u"], u64, ratio=2.0, slack=2e-5)
        smp = fl.sample(9, seed=5).to_numpy()
        x64, _ = O.flow_inverse(ri, p64, J.uniform(J.PRNGKey(5), (9, 5), -5.0, 5.0, part).astype(np.float64))
        assert_noise_floor(f"shuffle flow sample seed {seed} [{engine}]", smp, g[f"s{seed}_sample_9_seed5"], x64, ratio=2.0,
                           slack=2e-5)
        # inverse floors the dequantized column (bijectors.py:902-905): samples of column c are whole numbers
        assert np.array_equal(smp[:, 2], np.floor(smp[:, 2]))
        if seed == 3:
            again = Flow(_dictionary=fl._save_dict())
            again._plan().set_engine(engine)
            r0 = O.resolve_shuffles(info, J.PRNGKey(0), 5)
            assert_noise_floor(f"reloaded shuffle flow [{engine}]", again.log_prob(df), g["s3_reloaded_log_prob"],
                               O.flow_log_prob(r0, ("Uniform", (5, 5)), p64, X.astype(np.float64)), ratio=2.0, slack=2e-5,
                               below=LP_UNDERFLOW)
            assert np.abs(again.log_prob(df) - lp).max() > 1.0


def test_custom_error_models_run_on_the_host_and_reduce_on_the_device():
    """pzflow/flow.py:339-395 accepts any callable as data_error_model / condition_error_model
    (docs/tutorials/nongaussian_errors.ipynb).  Custom callables are called on the host with NumPy arrays and a
    PRNGKey-shaped key; their samples go through the fused kernels and are reduced on the device: log(mean(exp)) for
    log_prob (flow.py:463-464), mean then per-galaxy normalisation for posterior (722-737) -- against the oracle on the
    same samples; zero errors == the plain calls (tests/test_flow.py:236-239, 260-264 of the reference)."""
    from pzflow_b200 import Flow
    from pzflow_b200 import bijectors as B
    from pzflow_b200 import distributions as Dn
    calls = []

    def lognormal_model(key, X, Xerr, nsamples):
        assert np.asarray(key).shape == (2,) and np.asarray(key).dtype == np.uint32
        calls.append((X.shape, nsamples))
        rng = np.random.default_rng([int(k) for k in key])
        eps = rng.standard_normal((X.shape[0], nsamples, X.shape[1])).astype(np.float32)
        return X[:, None, :] * np.exp(eps * Xerr[:, None, :])         # multiplicative (non-Gaussian) errors

    ex = F.example_flow()
    cols = ("redshift", "u", "g", "r", "i", "z", "y")
    fl = Flow(cols, latent=Dn.Uniform(7, 5), data_error_model=lognormal_model)
    info = ex["bijector_info"]
    fl.set_bijector(B.Chain(B.ShiftBounds(info[1][0][1][0], info[1][0][1][1], 4.0), B.RollingSplineCoupling(7)),
                    params=ex["params"][1])
    rng = np.random.default_rng(8)
    n, S, seed = 300, 6, 17
    X = F.galaxy_rows(n, seed=31, oob_frac=0)
    df = pd.DataFrame(X.astype(np.float64), columns=cols)
    for c in ("u", "g", "r"):
        df[f"{c}_err"] = rng.uniform(0.0, 0.002, n)
    got = fl.log_prob(df, err_samples=S, seed=seed)
    Xerr = np.zeros_like(X)
    for j, c in enumerate(cols):
        if f"{c}_err" in df:
            Xerr[:, j] = df[f"{c}_err"].to_numpy()
    from pzflow_b200 import prng
    samples = lognormal_model(prng.key_from_seed(seed), X, Xerr, S).astype(np.float32)          # [n, S, 7]
    bi, li, p64 = _o64(ex)
    with np.errstate(all="ignore"):
        lp32 = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], samples.reshape(n * S, 7))
        lp64 = O.flow_log_prob(bi, li, p64, samples.reshape(n * S, 7).astype(np.float64))
        o32 = np.log(np.exp(lp32.reshape(n, S)).mean(axis=1, dtype=np.float32))
        o64 = np.log(np.exp(lp64.reshape(n, S)).mean(axis=1))
    assert_noise_floor("custom error model log_prob", got, o32, o64, ratio=2.0, slack=5e-6, below=LP_UNDERFLOW)
    assert not np.array_equal(got, fl.log_prob(df, err_samples=S, seed=seed + 1))
    # posterior: samples of the other six columns, the skipped column handed to the model as NaN (flow.py:370-373)
    G, ng = 40, 25
    grid = np.linspace(0.0, 2.3, G).astype(np.float32)
    pdf = fl.posterior(df.iloc[:ng], column="redshift", grid=grid, err_samples=S, seed=seed)
    Xn, En = X[:ng].copy(), Xerr[:ng].copy()
    Xn[:, 0] = np.nan
    En[:, 0] = np.nan
    rest = np.delete(lognormal_model(prng.key_from_seed(seed), Xn, En, S).astype(np.float32), 0, axis=2).reshape(ng * S, 6)
    un = O.flow_posterior(ex["bijector_info"], ex["latent_info"], ex["params"], rest, 0, grid, normalize=False,
                          nan_to_zero=False).reshape(ng, S, G).mean(axis=1)
    want = np.nan_to_num(un / O.trapezoid(un, grid)[:, None], nan=0.0)
    np.testing.assert_allclose(pdf, want, rtol=2e-3, atol=2e-4)
    # zero errors == plain
    plain = pd.DataFrame(X.astype(np.float64), columns=cols)
    lp0, lpS = fl.log_prob(plain), fl.log_prob(plain, err_samples=3, seed=1)
    fin = lp0 > LP_UNDERFLOW
    np.testing.assert_allclose(lpS[fin], lp0[fin], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(fl.posterior(plain.iloc[:ng], "redshift", grid, err_samples=3, seed=1),
                               fl.posterior(plain.iloc[:ng], "redshift", grid), rtol=1e-4, atol=1e-6)
    assert calls and all(ns in (S, 3) for _, ns in calls)
    # the step size of the host route does not change results
    fl._HOST_SAMPLE_ROWS = 64 * S
    try:
        np.testing.assert_array_equal(fl.log_prob(df, err_samples=S, seed=seed), got)
    finally:
        del fl._HOST_SAMPLE_ROWS


# --------------------------------------------------------------------------- #
# BASELINE's full sizes for configs 3, 4, 5: size-independent properties        #
# --------------------------------------------------------------------------- #
def test_full_size_properties_c3():
    """Config 3 at BASELINE's size -- 10^6 galaxies x 1000 redshift grid points, 4 GB of pdfs: every pdf integrates to
    one, galaxy shards (the 8-GPU partition) compute exactly what the whole call computes, a strided sample of galaxies
    equals exp(log_prob) of its expanded rows normalised by the trapezoid rule, and the single-pass form (normalised
    inside the chain kernel) agrees with the default one to float rounding."""
    import os
    from pzflow_b200.sharding import shard_rows
    ex = F.example_flow()
    plan = _plan(ex, "tc")
    ng, G = 1_000_000, 1000
    base = torch.from_numpy(np.ascontiguousarray(F.galaxy_rows(20000, seed=6, oob_frac=0.01)[:, 1:])).cuda()
    rest = base.repeat(ng // base.shape[0], 1).contiguous()
    rest += torch.linspace(0, 2e-3, ng, device="cuda")[:, None]        # distinct galaxies
    grid = torch.linspace(0.0, 2.3, G, device="cuda")
    pdf = plan.posterior(rest, 0, grid)
    assert pdf.shape == (ng, G)
    dx = (grid[1:] - grid[:-1]).double()
    integ = torch.cat([0.5 * (dx * (pdf[a:a + 100_000, 1:].double() + pdf[a:a + 100_000, :-1].double())).sum(1)
                       for a in range(0, ng, 100_000)])               # float64: only the kernel's own rounding is left
    # galaxies whose un-normalised pdf is not itself at the bottom of float32's range (1 % of the rows sit outside the
    # training box; dividing by a denormal integral is as meaningless here as in the reference)
    un_max = torch.cat([plan.posterior(rest[a:a + 250_000], 0, grid, normalize=False).amax(1) for a in range(0, ng, 250_000)])
    live = un_max > 1e-20
    assert live.float().mean() > 0.95
    dev = (integ[live] - 1.0).abs()
    print(f"C3 full size: {int(live.sum())} live galaxies, |integral - 1| median {float(dev.median()):.2e} max {float(dev.max()):.2e}")
    assert dev.median() < 1e-6 and dev.max() < 3e-5                    # 1000-term float32 trapezoid sums
    assert not torch.isnan(pdf[::997]).any()
    for r in (0, 3, 7):
        a, b = shard_rows(ng, r, 8)
        assert torch.equal(plan.posterior(rest[a:b], 0, grid), pdf[a:b])
    pick = torch.arange(0, ng, 15625, device="cuda")                   # 64 galaxies across the whole range
    rows = torch.cat([grid.repeat(len(pick))[:, None], rest[pick].repeat_interleave(G, 0)], dim=1)
    un = torch.exp(plan.log_prob(rows)).reshape(len(pick), G)
    want = torch.nan_to_num(un / (0.5 * ((grid[1:] - grid[:-1]) * (un[:, 1:] + un[:, :-1])).sum(1, keepdim=True)), nan=0.0)
    assert torch.allclose(pdf[pick], want, rtol=3e-6, atol=1e-12)
    os.environ["PZB_FUSE_POSTERIOR"] = "1"
    try:
        fused = plan.posterior(rest[:200_000], 0, grid)
    finally:
        os.environ.pop("PZB_FUSE_POSTERIOR", None)
    assert torch.allclose(fused, pdf[:200_000], rtol=3e-6, atol=1e-12)


def test_full_size_properties_c4():
    """Config 4 at BASELINE's size -- FlowEnsemble of 4 conditional flows on 10^8 rows: the ensemble value is
    log(mean(exp)) of the four flows' log_probs (flowEnsemble.py:243), row shards equal the whole call, repeated calls
    are bit-identical, and the first rows equal the oracle's ensemble rule."""
    from pzflow_b200.plan import ensemble_logmeanexp
    from pzflow_b200.sharding import shard_rows
    cfgs = F.config_c4(4)
    plans = [_plan(c, "tc") for c in cfgs]
    ex = F.example_flow()
    mins = torch.as_tensor(ex["bijector_info"][1][0][1][0], device="cuda")
    maxs = torch.as_tensor(ex["bijector_info"][1][0][1][1], device="cuda")
    n = 100_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    X = torch.rand((n, 7), generator=gen, device="cuda") * (maxs - mins) + mins
    Cn = torch.randn((n, 2), generator=gen, device="cuda")
    lps = torch.empty((4, n), dtype=torch.float32, device="cuda")
    for f, p in enumerate(plans):
        p.log_prob(X, Cn, out=lps[f])
    ens = ensemble_logmeanexp(lps)
    assert torch.isfinite(ens).float().mean() > 0.99
    m = 2_000_000
    ref = torch.log(torch.exp(lps[:, :m].double()).mean(0))
    fin = torch.isfinite(ref) & (ref > -80)
    assert torch.allclose(ens[:m][fin].double(), ref[fin], rtol=1e-5, atol=1e-5)
    assert torch.equal(plans[2].log_prob(X, Cn), lps[2])
    for r in (0, 5):
        a, b = shard_rows(n, r, 8)
        assert torch.equal(plans[1].log_prob(X[a:b], Cn[a:b]), lps[1][a:b])
    k = 512
    flows = [(c["bijector_info"], c["latent_info"], c["params"]) for c in cfgs]
    want = O.ensemble_log_prob(flows, _np(X[:k]), _np(Cn[:k]))
    assert_tolerance("C4 full-size head rows", _np(ens[:k]), want, min_frac=0.97, max_class_mismatch=1, below=LP_UNDERFLOW)


def test_full_size_properties_c5():
    """Config 5 at BASELINE's size -- 32-D, K = 32, hidden 256, 10^6 rows on the wide tensor-core engine: determinism,
    shard independence, inverse(forward(x)) = x and log-det antisymmetry (tests/test_bijectors.py:71-86 of the reference),
    log_prob = latent log-density + log-det."""
    from pzflow_b200.sharding import shard_rows
    cfg = F.config_c5()
    plan = _plan(cfg, "wide")
    n = 1_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    X = torch.rand((n, 32), generator=gen, device="cuda") * 2.0 - 1.0
    lp = plan.log_prob(X)
    assert torch.equal(lp, plan.log_prob(X))
    for r in (0, 6):
        a, b = shard_rows(n, r, 8)
        assert torch.equal(plan.log_prob(X[a:b]), lp[a:b])
    u, ld = plan.forward(X)
    xr, ldi = plan.inverse(u)
    inside = (u.abs() < 4.9).all(dim=1) & torch.isfinite(lp)
    assert inside.float().mean() > 0.9
    err = (xr - X).abs()[inside]
    assert err.median() < 5e-6 and torch.quantile(err.flatten()[:4_000_000], 0.999) < 1e-3
    asym = ((ld + ldi).abs() / (1.0 + ld.abs()))[inside]             # sums of 512 float32 log terms of size |ld|
    print(f"C5 full size: |ld| median {float(ld.abs()[inside].median()):.1f}, relative antisymmetry median {float(asym.median()):.2e}")
    assert asym.median() < 1e-4
    lat = torch.from_numpy(O.latent_log_prob(cfg["latent_info"], _np(u[:4096]))).cuda()
    ok = inside[:4096] & torch.isfinite(lat)
    assert torch.allclose(lp[:4096][ok], (lat + ld[:4096])[ok], rtol=1e-5, atol=2e-4)
    assert plan.overflow_rows() == 0


def test_stand_alone_spline_and_network_of_the_reference_utils():
    """pzflow/utils.py:64-227 (RationalQuadraticSpline) and 22-49 (DenseReluNetwork) as stand-alone functions: forward,
    inverse and periodic splines against the oracle, with rows outside the box, on knots, beyond the last knot and NaN;
    Flow._get_err_samples in the reference's row order."""
    from pzflow_b200 import utils as U
    rng = np.random.default_rng(12)
    for K, L, periodic in ((16, 4, False), (8, 3, True), (32, 2, False), (1, 2, False)):
        n, B = 4000, 5.0
        W = O._softmax(rng.normal(size=(n, L, K)).astype(np.float32)) * np.float32(2 * B)
        H = O._softmax(rng.normal(size=(n, L, K)).astype(np.float32)) * np.float32(2 * B)
        D = O._softplus(rng.normal(size=(n, L, K - 1 + int(periodic))).astype(np.float32))
        x = rng.uniform(-B, B, size=(n, L)).astype(np.float32)
        x[:40] = rng.uniform(-3 * B, 3 * B, size=(40, L)).astype(np.float32)      # out of bounds: identity, log-det 0
        x[40, 0], x[41, 0], x[42, 0] = -B, B, np.nan
        knots = np.concatenate([np.full((n, L, 1), -B, np.float32), -np.float32(B) + np.cumsum(W, -1, dtype=np.float32)], -1)
        if K > 2:
            x[50:60, 1] = knots[50:60, 1, 2]                                         # exactly on a knot: the bin to its right
        for inverse in (False, True):
            with np.errstate(all="ignore"):
                yo, ldo, io = O.rational_quadratic_spline(x, W, H, D, B, periodic=periodic, inverse=inverse, return_idx=True)
            y, ld = U.RationalQuadraticSpline(x, W, H, D, B, periodic=periodic, inverse=inverse)
            assert y.shape == x.shape and ld.shape == (n,)
            same_nan = np.isnan(y) == np.isnan(yo)
            assert same_nan.all(), (K, periodic, inverse, int((~same_nan).sum()))
            ok = ~np.isnan(yo)
            np.testing.assert_allclose(y[ok], yo[ok], rtol=2e-6, atol=2e-6)
            okr = ~np.isnan(ldo) & ~np.isnan(ld)
            assert (np.isnan(ldo) == np.isnan(ld)).all()
            np.testing.assert_allclose(ld[okr], ldo[okr], rtol=1e-5, atol=2e-5)
            if not periodic:
                oob = (x <= -B) | (x >= B)
                assert np.array_equal(y[oob], x[oob])
        # a device tensor in gives device tensors out
        yt, ldt = U.RationalQuadraticSpline(torch.from_numpy(x).cuda(), W, H, D, B, periodic=periodic)
        assert yt.is_cuda and ldt.is_cuda
        # round trip where the forward value stayed inside the box
        yf, ldf = U.RationalQuadraticSpline(x, W, H, D, B, periodic=periodic)
        xb, ldb = U.RationalQuadraticSpline(yf, W, H, D, B, periodic=periodic, inverse=True)
        inside = (np.abs(x) < B * 0.999).all(axis=1) & ~np.isnan(yf).any(axis=1)
        assert np.median(np.abs(xb - x)[inside]) < 1e-5 and np.median(np.abs(ldf + ldb)[inside]) < 1e-4
    init_fun, apply_fun = U.DenseReluNetwork(11, 2, 32)
    _, params = init_fun(3, (7, 5))
    xin = rng.normal(size=(300, 5)).astype(np.float32)
    np.testing.assert_allclose(apply_fun(params, xin), O.dense_relu_network(params, xin), rtol=1e-5, atol=1e-5)
    # Flow._get_err_samples (pzflow/flow.py:339-395): [N * S, W], row i's samples consecutive, zero error -> the row itself
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    df = pd.DataFrame(F.galaxy_rows(10, seed=2).astype(np.float64), columns=flow.data_columns)
    df["u_err"] = 0.1
    s = flow._get_err_samples(5, df, 4, type="data")
    assert s.shape == (40, 7) and s.dtype == np.float32
    np.testing.assert_array_equal(s[:, 0], np.repeat(df["redshift"].to_numpy().astype(np.float32), 4))
    assert np.abs(s[:, 1] - np.repeat(df["u"].to_numpy(), 4)).max() > 0 and np.abs(s[:, 1] - np.repeat(df["u"].to_numpy(), 4)).max() < 1.0
    assert flow._get_err_samples(5, df, 4, type="conditions").shape == (40, 1)
    assert flow._get_err_samples(5, df, 4, type="data", skip="redshift").shape == (40, 6)


def test_conditional_ensemble_sampling_against_reference_golden(stream_layout):
    """FlowEnsemble.sample with conditions (pzflow/flowEnsemble.py:409-471) against what the reference's own code returned
    over the stand-in with jax's key schedule: which flow draws which condition row (pandas' shuffle, jax.random.permutation
    of the chunks, NumPy's choice for fewer rows than flows) AND the drawn values (Uniform latents: jax's stream)."""
    from pzflow_b200 import FlowEnsemble
    from pzflow_b200 import bijectors as B
    from pzflow_b200 import distributions as Dn
    from conftest import STREAM_GOLDEN
    g = golden(STREAM_GOLDEN[stream_layout])
    mins, maxs = g["mins"][:3], g["maxs"][:3]
    ens = FlowEnsemble(("a", "b", "c"), B.Chain(B.ShiftBounds(mins, maxs, 4.0),
                                                B.RollingSplineCoupling(3, K=8, hidden_dim=32, n_conditions=2)),
                       latent=Dn.Uniform(3, 5), conditional_columns=("p", "q"), N=3)
    for i, fl in enumerate(ens._ensemble.values()):
        fl._params = (fl._params[0], F.unpack_tree(f"cens{i}_params", g))
        fl._plan_cache.clear()
    conds = pd.DataFrame(g["cens_conditions"], columns=("p", "q"), index=g["cens_index"])
    for tag in ("many", "few", "exact"):
        ns, seed, nrow = (int(v) for v in g[f"cens_{tag}_args"])
        out = ens.sample(ns, conditions=conds.iloc[:nrow], save_conditions=True, seed=seed)
        assert list(out.columns) == ["a", "b", "c", "p", "q"]
        np.testing.assert_array_equal(out.index.to_numpy(), g[f"cens_{tag}_index"])
        want = g[f"cens_{tag}_sample"]
        err = np.abs(out.to_numpy() - want)
        print(f"ensemble conditional sample [{tag}]: max |err| {err.max():.2e}")
        # rows with the same index (the nsamples copies of one condition) may come back in either order: compare as sets
        for idx in np.unique(g[f"cens_{tag}_index"]):
            a = out.to_numpy()[out.index.to_numpy() == idx]
            b = want[g[f"cens_{tag}_index"] == idx]
            a, b = a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])]
            np.testing.assert_allclose(a, b, rtol=3e-4, atol=3e-4)


@pytest.mark.parametrize("engine", ["simt", "tc", "wide"])
def test_posteriors_overlay_the_figures_real_jax_plotted(engine, notebook_era_streams):
    """docs/tutorials/marginalization.ipynb cells 9-15 through the public API: flow.posterior(samples, "redshift", grid)
    and the same with u / y flagged 99 and marg_rules over 40 x 40 values -- the curves land on the lines of the figures
    the reference's authors made with real JAX (tests/golden/nb_posterior_plots.npz; one pixel = 0.8 % of a panel's
    tallest peak).  tests/test_oracle.py holds the oracle's version of the check and its negative controls."""
    import plots as PL
    from test_oracle import check_against_the_notebook_figures
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    flow._plan().set_engine(engine)
    samples = flow.sample(2, seed=0)                                   # the notebook's own first step (cell 5)
    nb = golden("ref_shim_example_flow.npz")["nb_X"]
    np.testing.assert_allclose(samples.to_numpy(), nb, rtol=3e-6)
    grid = PL.NB_GRID
    plain = flow.posterior(samples, column="redshift", grid=grid)
    flagged = samples.copy()
    flagged.iloc[0, 1] = 99
    flagged.iloc[1, 1] = 99
    flagged.iloc[1, -1] = 99
    rules = {"flag": 99, "u": lambda row: PL.NB_U_GRID, "y": lambda row: PL.NB_Y_GRID}
    marg = flow.posterior(flagged, column="redshift", grid=grid, marg_rules=rules)
    check_against_the_notebook_figures(samples.to_numpy(), plain, marg, engine)
    flow._plan().set_engine("auto")

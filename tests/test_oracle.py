"""The oracle against (1) the reference's own code run over oracle/jax_shim (committed golden
vectors), (2) the reference test-suite's properties, (3) the JAX-produced anchor in the docs."""
import numpy as np
import pytest

from conftest import golden, zero_class
from oracle import fixtures as F
from oracle import nsf_oracle as O

TIGHT = dict(rtol=2e-6, atol=2e-6)  # same primitives, same order: only ulp-level slack


def _close(a, b, **kw):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    za, zb = zero_class(a), zero_class(b)
    assert (za == zb).all()
    np.testing.assert_allclose(a[~za], b[~zb], **(kw or TIGHT))


# --------------------------------------------------------------------------- #
# (1) golden vectors produced by the reference's code                          #
# --------------------------------------------------------------------------- #
def test_example_flow_golden():
    z = golden("ref_shim_example_flow.npz")
    ex = F.example_flow()
    bi, li, p = ex["bijector_info"], ex["latent_info"], ex["params"]
    X, U = z["X"], z["U"]
    _close(O.flow_log_prob(bi, li, p, X), z["log_prob"])
    u, ld = O.apply_bijector(bi, p[1], X, np.zeros((len(X), 1), np.float32))
    _close(u, z["forward_u"])
    _close(ld, z["forward_ld"])
    x, ldi = O.flow_inverse(bi, p, U)
    _close(x, z["inverse_x"])
    _close(ldi, z["inverse_ld"])
    rest = np.ascontiguousarray(X[:24, 1:])
    _close(O.flow_posterior(bi, li, p, rest, 0, z["grid"]), z["posterior"])
    _close(O.flow_posterior(bi, li, p, rest, 0, z["grid"], batch_size=7), z["posterior_batched"])
    _close(O.flow_posterior(bi, li, p, rest, 0, z["grid"], normalize=False, nan_to_zero=False),
           z["posterior_unnorm"])
    rest_r = np.ascontiguousarray(np.delete(X[:24], 3, axis=1))
    _close(O.flow_posterior(bi, li, p, rest_r, 3, z["grid_r"]), z["posterior_r"])
    # edge rows really exercise the edge paths
    assert zero_class(z["log_prob"])[[2, 3]].all()  # far out of bounds / NaN input -> zero probability
    assert not zero_class(z["log_prob"])[4]


def test_bijector_battery_golden():
    """Hot-path rows of the reference's tests/test_bijectors.py:16-51 on its fixed 3x7 input."""
    z = golden("ref_shim_bijectors.npz")
    x = z["x"]
    for name in z["names"]:
        name = str(name)
        info = F.unpack_tree(f"{name}_info", z)
        params = F.unpack_tree(f"{name}_params", z)
        cond = z[f"{name}_cond"]
        y, ld = O.apply_bijector(info, params, x, cond)
        _close(y, z[f"{name}_fwd"])
        _close(ld, z[f"{name}_fwd_ld"])
        xi, ldi = O.apply_bijector(info, params, y, cond, inverse=True)
        _close(xi, z[f"{name}_inv_of_fwd"])
        _close(ldi, z[f"{name}_inv_of_fwd_ld"])
        xi2, ldi2 = O.apply_bijector(info, params, x, cond, inverse=True)
        _close(xi2, z[f"{name}_inv"])
        _close(ldi2, z[f"{name}_inv_ld"])
        # the reference's own assertions (tests/test_bijectors.py:71-86); its atol of 1e-6 is a few
        # float32 ulps of these log-dets (|ld| up to ~5), so NumPy-vs-XLA primitive noise gets 4 ulps
        assert y.shape == x.shape and ld.shape == x.shape[:1]
        np.testing.assert_allclose(xi, x, atol=2e-6, err_msg=name)
        np.testing.assert_allclose(ld, -ldi, atol=4e-6, err_msg=name)


def test_default_flow_and_ensemble_golden():
    z = golden("ref_shim_default_flows.npz")
    # C1: default 2-D flow, CentBeta13(2, 5)
    info, bp = F.unpack_tree("c1_info", z), F.unpack_tree("c1_params", z)
    li, p = ("CentBeta13", (2, 5)), ((), bp)
    X = z["c1_X"]
    _close(O.flow_log_prob(info, li, p, X), z["c1_log_prob"])
    # pdf = exp(lp): an ulp of lp (|lp| ~ 5) is a few 1e-6 relative on the pdf
    _close(O.flow_posterior(info, li, p, X[:32, 1:], 0, z["c1_grid"]), z["c1_posterior_x"], rtol=1e-5, atol=1e-7)
    _close(O.flow_posterior(info, li, p, X[:32, :1], 1, z["c1_grid"], normalize=False), z["c1_posterior_y"],
           rtol=1e-5, atol=1e-7)
    _close(O.flow_inverse(info, p, z["c1_sample_u"])[0], z["c1_sample_x"])
    # conditional ensemble (flowEnsemble.py:227-243, 317-351; flow.py:324-337)
    cols = [str(c) for c in z["ens_columns"]]
    df = z["ens_df"]
    Xd = df[:, :3].astype(np.float32)
    cond = ((df[:, 3:].astype(np.float32) - z["ens_cond_means"]) / z["ens_cond_stds"]).astype(np.float32)
    assert cols == ["a", "b", "c", "p", "q"]
    flows = [(F.unpack_tree(f"ens{i}_info", z), ("CentBeta13", (3, 5)), ((), F.unpack_tree(f"ens{i}_params", z)))
             for i in range(3)]
    _close(O.flow_log_prob(*flows[0], Xd, cond), z["ens_flow0_log_prob"])
    _close(O.ensemble_log_prob(flows, Xd, cond), z["ens_log_prob"])
    _close(O.ensemble_log_prob(flows, Xd, cond, return_ensemble=True), z["ens_log_prob_all"])
    g = z["ens_grid"]
    _close(O.ensemble_posterior(flows, Xd[:20, 1:], 0, g, cond[:20]), z["ens_posterior"], rtol=1e-5, atol=1e-6)
    _close(O.ensemble_posterior(flows, Xd[:20, 1:], 0, g, cond[:20], return_ensemble=True),
           z["ens_posterior_all"], rtol=1e-5, atol=1e-6)
    _close(O.flow_posterior(*flows[0], np.delete(Xd[:20], 1, axis=1), 1, g, cond[:20]),
           z["ens_flow0_posterior_b"], rtol=1e-5, atol=1e-6)
    _close(O.flow_inverse(flows[0][0], flows[0][2], z["ens_sample_u"], z["ens_sample_cond"])[0],
           z["ens_sample_x"])


# --------------------------------------------------------------------------- #
# (2) properties the reference's tests assert                                  #
# --------------------------------------------------------------------------- #
X37 = np.array([[0.2, 0.1, -0.3, 0.5, 0.1, -0.4, -0.3],
                [0.6, 0.5, 0.2, 0.2, -0.4, -0.1, 0.7],
                [0.9, 0.2, -0.3, 0.3, 0.4, -0.4, -0.1]], np.float32)


@pytest.mark.parametrize("info,cond", [
    (("NeuralSplineCoupling", (16, 5, 2, 128, None, 0, False)), np.zeros((3, 1), np.float32)),
    (("NeuralSplineCoupling", (16, 3, 2, 128, 3, 3, False)), np.arange(9, dtype=np.float32).reshape(3, 3)),
    (F.rolling_info(2), np.zeros((3, 1), np.float32)),
    (F.rolling_info(2, 1, 16, 3, 2, 128, None, 0, True), np.zeros((3, 1), np.float32)),
    (("ShiftBounds", (-0.5, 0.9, 5)), np.zeros((3, 1), np.float32)),
    (("ShiftBounds", (-np.ones(7, np.float32), 1.1 * np.ones(7, np.float32), 3)), np.zeros((3, 1), np.float32)),
    (("StandardScaler", (np.linspace(-1, 1, 7), np.linspace(1, 8, 7))), np.zeros((3, 1), np.float32)),
    (("Roll", (2,)), np.zeros((3, 1), np.float32)),
])
def test_bijective_and_logdet_antisymmetric(info, cond):
    """tests/test_bijectors.py:53-86: shapes, inverse(forward(x)) == x and fwd_ld == -inv_ld at atol 1e-6."""
    params = O.synth_bijector_params(info, 7, np.random.default_rng(0))
    y, ld = O.apply_bijector(info, params, X37, cond)
    assert y.shape == X37.shape and ld.shape == (3,)
    xi, ldi = O.apply_bijector(info, params, y, cond, inverse=True)
    # the reference asserts atol 1e-6 under XLA; NumPy primitives get 4 float32 ulps of |ld| ~ 2
    np.testing.assert_allclose(xi, X37, atol=2e-6)
    np.testing.assert_allclose(ld, -ldi, atol=4e-6)


def test_batched_posterior_equals_unbatched():
    """tests/test_flow.py:275-287."""
    ex = F.example_flow()
    X = F.galaxy_rows(12, seed=3, oob_frac=0)
    grid = np.linspace(0, 2.3, 30).astype(np.float32)
    a = O.flow_posterior(ex["bijector_info"], ex["latent_info"], ex["params"], X[:, 1:], 0, grid)
    b = O.flow_posterior(ex["bijector_info"], ex["latent_info"], ex["params"], X[:, 1:], 0, grid, batch_size=5)
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-5)  # different GEMM shapes -> fp32 noise floor
    np.testing.assert_allclose(O.trapezoid(a, grid), 1.0, rtol=1e-5)


def test_ensemble_rules():
    """tests/test_ensemble.py:20-56: column i == flow i; log(mean(exp)); mean then normalise."""
    flows = [(c["bijector_info"], c["latent_info"], c["params"]) for c in F.config_c4(3)]
    rng = np.random.default_rng(0)
    ex = F.example_flow()
    mins, maxs = ex["bijector_info"][1][0][1][0], ex["bijector_info"][1][0][1][1]
    X = rng.uniform(mins, maxs, size=(40, 7)).astype(np.float32)
    cond = rng.normal(size=(40, 2)).astype(np.float32)
    ens = O.ensemble_log_prob(flows, X, cond, return_ensemble=True)
    assert ens.shape == (40, 3)
    for i, fl in enumerate(flows):
        np.testing.assert_array_equal(ens[:, i], O.flow_log_prob(*fl, X, cond))
    np.testing.assert_allclose(O.ensemble_log_prob(flows, X, cond), np.log(np.exp(ens).mean(axis=1)), rtol=1e-6)
    grid = np.linspace(0, 2.3, 20).astype(np.float32)
    per = np.stack([O.flow_posterior(*fl, X[:8, 1:], 0, grid, cond[:8], normalize=False) for fl in flows], 1)
    mean = per.mean(axis=1)
    np.testing.assert_allclose(O.ensemble_posterior(flows, X[:8, 1:], 0, grid, cond[:8]),
                               mean / O.trapezoid(mean, grid)[:, None], rtol=1e-5)


# --------------------------------------------------------------------------- #
# (3) the JAX-produced anchor + edge semantics                                 #
# --------------------------------------------------------------------------- #
def test_notebook_samples_anchor():
    """docs/tutorials/marginalization.ipynb:133-135 prints two samples drawn by the real JAX
    reference from the shipped flow: they must map into the Uniform(7,5) support, have finite
    log_prob, and round-trip."""
    ex = F.example_flow()
    X = golden("ref_shim_example_flow.npz")["nb_X"]
    u, _ = O.apply_bijector(ex["bijector_info"], ex["params"][1], X, np.zeros((2, 1), np.float32))
    assert (np.abs(u) <= 5).all()
    lp = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X)
    np.testing.assert_allclose(lp, [5.4340, 6.7095], atol=2e-3)  # regression seed
    xr, _ = O.flow_inverse(ex["bijector_info"], ex["params"], u)
    np.testing.assert_allclose(xr, X, atol=2e-5)


def test_spline_edge_semantics():
    """utils.py:145-152, 177-178, 191-193: out-of-bounds identity per element with zero log-det
    contribution; left-closed bins; NaN in -> zero probability (flow.py:406)."""
    rng = np.random.default_rng(1)
    K, B = 16, 5.0
    W = O._softmax(rng.normal(size=(4, 3, K)).astype(np.float32)) * np.float32(2 * B)
    H = O._softmax(rng.normal(size=(4, 3, K)).astype(np.float32)) * np.float32(2 * B)
    D = O._softplus(rng.normal(size=(4, 3, K - 1)).astype(np.float32))
    x = np.array([[-5.0, 0.3, 5.0], [7.0, -9.5, 4.999], [0.0, -4.999, 20.0], [1.0, 2.0, 3.0]], np.float32)
    y, ld, idx = O.rational_quadratic_spline(x, W, H, D, B, return_idx=True)
    oob = (x <= -B) | (x >= B)
    np.testing.assert_array_equal(y[oob], x[oob])
    assert np.isfinite(y).all() and np.isfinite(ld).all()
    # x exactly on an interior knot goes to the bin on its right
    xk = np.concatenate([np.full((4, 3, 1), -B, np.float32), -np.float32(B) + np.cumsum(W, -1, dtype=np.float32)], -1)
    xon = xk[:, :, 5].copy()
    _, _, idx_on = O.rational_quadratic_spline(xon, W, H, D, B, return_idx=True)
    assert (idx_on == 5).all()
    # rows that are entirely out of bounds have exactly zero log-det
    y2, ld2 = O.rational_quadratic_spline(np.full((4, 3), 6.0, np.float32), W, H, D, B)
    assert (ld2 == 0).all() and (y2 == 6.0).all()


def test_nan_and_support_go_to_zero_probability():
    ex = F.example_flow()
    X = F.galaxy_rows(8, seed=2, oob_frac=0)
    X[1, 2] = np.nan
    X[2] = 1e3
    lp = O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X)
    assert zero_class(lp)[[1, 2]].all() and np.isfinite(lp[[0, 3, 4]]).all()
    assert lp[1] == np.finfo(np.float32).min  # jnp.nan_to_num(nan=-inf) then -inf -> finfo.min


def test_extra_bijectors_and_latents_golden():
    """ColorTransform / Reverse / Scale / InvSoftplus (and Chains of them with the
    coupling layers) and the Normal / CentBeta latents, against what the reference's own code returned
    (oracle/make_golden_shim.py::case_extra_bijectors)."""
    z = golden("ref_shim_extra.npz")
    x = z["x"]
    cond = np.zeros((x.shape[0], 1), np.float32)
    for name in z["names"]:
        name = str(name)
        info = F.unpack_tree(f"{name}_info", z)
        params = F.unpack_tree(f"{name}_params", z)
        with np.errstate(all="ignore"):
            y, ld = O.apply_bijector(info, params, x, cond)
            xi, ldi = O.apply_bijector(info, params, y, cond, inverse=True)
            xi2, ldi2 = O.apply_bijector(info, params, x, cond, inverse=True)
        for got, key in ((y, "fwd"), (ld, "fwd_ld"), (xi, "inv_of_fwd"), (ldi, "inv_of_fwd_ld"), (xi2, "inv"),
                         (ldi2, "inv_ld")):
            want = z[f"{name}_{key}"]
            assert np.array_equal(np.isnan(got), np.isnan(want)), (name, key)
            np.testing.assert_allclose(got, want, rtol=3e-6, atol=3e-6, err_msg=f"{name} {key}")
    u = z["latent_u"]
    np.testing.assert_allclose(O.latent_log_prob(("Normal", (3,)), u), z["normal_lp"], rtol=2e-6)
    cbp = tuple((float(a), float(b)) for a, b in z["centbeta_params"])
    for params, key in ((cbp, "centbeta_lp"), (tuple((0.0, 0.0) for _ in range(3)), "centbeta_lp_default")):
        got, want = O.centbeta_log_prob(params, u, 5), z[key]
        assert np.array_equal(np.isneginf(got), np.isneginf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=3e-6, atol=3e-6)
    np.testing.assert_allclose(O.tdist_log_prob(np.log(np.float32(30.0)), u), z["tdist_lp_default"], rtol=3e-6, atol=3e-6)
    np.testing.assert_allclose(O.tdist_log_prob(z["tdist_lognu"], u), z["tdist_lp"], rtol=3e-6, atol=3e-6)
    jinfo = F.unpack_tree("joint_info", z)
    got = O.joint_log_prob(jinfo, [(), (cbp[0],), ()], u)
    want = z["joint_lp"]
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=3e-6, atol=3e-6)


# --------------------------------------------------------------------------- #
# (4) jax's random streams (oracle/jax_prng.py)                               #
# --------------------------------------------------------------------------- #
def test_threefry_known_answers():
    """Random123's threefry2x32_20 known answers (the vectors jax's own test-suite checks)."""
    from oracle import jax_prng as J
    for ctr, key, want in J.KAT:
        a, b = J.threefry2x32(key, [ctr[0]], [ctr[1]])
        assert (int(a[0]), int(b[0])) == want
    # jax's documentation: random.split(PRNGKey(0))
    np.testing.assert_array_equal(J.split(J.PRNGKey(0)), [[4146024105, 967050713], [2718843009, 1272950319]])


def test_jax_normal_known_answers():
    """``jax.random.normal`` restated (threefry bits -> uniform on (-1, 1) -> XLA's float32 erf_inv): the values jax's
    own documentation prints -- "JAX - The Sharp Bits" (PRNGKey(0) and the sub-keys of its split chain) and the PRNG
    tutorial (PRNGKey(42)) -- digit for digit.  This is the stream of the default error model's eps
    (pzflow/utils.py:59) and of the Normal latent's draws (pzflow/distributions.py:297-303)."""
    from oracle import jax_prng as J
    k0 = J.PRNGKey(0)
    assert J.normal(k0, (1,))[0] == np.float32(-0.20584226)
    key, sub = J.split(k0)
    assert J.normal(sub, (1,))[0] == np.float32(-1.2515389)
    key, sub = J.split(key)
    np.testing.assert_array_equal(sub, [1278412471, 2182328957])
    assert J.normal(sub, (1,))[0] == np.float32(-0.58665055)
    assert J.normal(J.PRNGKey(42), ())[()] == np.float32(-0.18471177)
    key, draws = J.PRNGKey(42), []                                   # the split loop of the (older) PRNG tutorial
    for _ in range(3):
        key, sub = J.split(key)
        draws.append(float(J.normal(sub, ())))
    assert draws == [1.369469404220581, -0.19947023689746857, -2.298278331756592]
    # the PARTITIONABLE layout (jax >= 0.5's default, what the reference's required jax draws): the same pages of the
    # current documentation
    assert J.normal(J.PRNGKey(42), (), partitionable=True)[()] == np.float32(-0.028304616)
    key, draws = J.PRNGKey(42), []
    for _ in range(3):
        key, sub = J.split(key, 2, partitionable=True)
        draws.append(float(J.normal(sub, (), partitionable=True)))
    assert draws == [0.6057640314102173, -0.21089035272598267, -0.3948981463909149]
    assert J.uniform(J.PRNGKey(0), (), 0.0, 1.0, partitionable=True)[()] == np.float32(0.947667)
    assert J.normal(J.PRNGKey(0), (1,), partitionable=True)[0] == np.float32(1.6226422)
    # the polynomial against the exact function, over the whole open interval and in the tails
    from scipy.special import erfinv
    x = np.concatenate([np.linspace(-0.9999999, 0.9999999, 20001), [np.nextafter(np.float32(-1), np.float32(0))]]).astype(np.float32)
    got, ref = J.erf_inv32(x), erfinv(x.astype(np.float64))
    assert np.abs(got / np.where(ref == 0, 1, ref) - 1)[ref != 0].max() < 5e-6
    big = J.normal(J.PRNGKey(7), (200000, 3))
    assert abs(big.mean()) < 0.005 and abs(big.std() - 1) < 0.005 and abs((big ** 3).mean()) < 0.02
    assert abs((big ** 4).mean() - 3) < 0.05


def test_notebook_samples_from_the_seed():
    """docs/tutorials/marginalization.ipynb:69, 133-135: ``flow.sample(2, seed=0)`` of the shipped flow, printed by the
    real JAX reference.  Drawing the latents of seed 0 in jax's original threefry layout and pushing them through the
    oracle's inverse reproduces the printed rows to the printed precision -- the one end-to-end anchor against real
    JAX/XLA output in the tree: it pins the stream layout AND the whole inverse chain (14 values, 1e-6 relative)."""
    from oracle import jax_prng as J
    ex = F.example_flow()
    X = golden("ref_shim_example_flow.npz")["nb_X"]
    u = J.uniform(J.PRNGKey(0), (2, 7), -5.0, 5.0)
    x, _ = O.flow_inverse(ex["bijector_info"], ex["params"], u)
    np.testing.assert_allclose(x, X, rtol=1.5e-6, atol=0)
    # the other stream layout (jax >= 0.5's default) does not produce them: the notebook was run with an older jax
    u2 = J.uniform(J.PRNGKey(0), (2, 7), -5.0, 5.0, partitionable=True)
    x2, _ = O.flow_inverse(ex["bijector_info"], ex["params"], u2)
    assert np.abs(x2 - X).max() > 0.1


def test_shuffle_and_dequantizer_semantics():
    """bijectors.py:798-810 (Shuffle) and 892-907 (UniformDequantizer as executed: identity forward, floor inverse)."""
    from oracle import jax_prng as J
    rng = np.random.default_rng(0)
    X = rng.normal(size=(5, 7)).astype(np.float32) * 3
    info = O.resolve_shuffles(("Chain", (("Shuffle", ()), ("UniformDequantizer", ([1, 4],)))), J.PRNGKey(2), 7)
    perm = np.asarray(info[1][0][1][0])
    y, ld = O.apply_bijector(info, [(), ()], X, None)
    np.testing.assert_array_equal(y, X[:, perm])
    assert not ld.any()
    back, _ = O.apply_bijector(info, [(), ()], y, None, inverse=True)
    want = X.copy()
    cols = perm[[1, 4]]   # the dequantizer sits behind the shuffle: it floors the columns that landed at 1 and 4
    want[:, cols] = np.floor(want[:, cols])
    np.testing.assert_array_equal(back, want)


def test_log_det_is_the_jacobian_of_the_map():
    """An anchor that needs no reference run: the log-determinant the restated chain returns (pzflow/utils.py:182-196
    summed through bijectors.py:155-160, 760-765) equals log|det J| of the restated MAP, J by central differences in
    float64 -- for the shipped flow (whose sampling map IS pinned against real JAX output above) and for a conditional
    chain with a StandardScaler, Roll and periodic splines.  Together with the JAX-printed samples this fixes log_prob:
    the map is JAX's, and the density follows from the map."""
    ex = F.example_flow()
    cases = [(ex["bijector_info"], O.cast_params(ex["params"], np.float64)[1], F.galaxy_rows(6, seed=3, oob_frac=0).astype(np.float64), None)]
    rng = np.random.default_rng(4)
    info = ("Chain", (("StandardScaler", (np.linspace(-1, 1, 5), np.linspace(1, 2, 5))),
                      F.rolling_info(3, 2, 8, 4.0, 2, 32, None, 2, True)))
    params = O.cast_params(((), O.synth_bijector_params(info, 5, rng, out_scale=2.0)), np.float64)[1]
    cases.append((info, params, rng.uniform(-2.5, 2.5, size=(5, 5)), rng.normal(size=(5, 2))))
    for info, params, X, cond in cases:
        n, D = X.shape
        c = np.zeros((n, 1)) if cond is None else cond
        _, ld = O.apply_bijector(info, params, X, c)
        h = 1e-6
        for i in range(n):
            J = np.empty((D, D))
            for j in range(D):
                e = np.zeros(D)
                e[j] = h
                up, _ = O.apply_bijector(info, params, (X[i] + e)[None], c[i:i + 1])
                dn, _ = O.apply_bijector(info, params, (X[i] - e)[None], c[i:i + 1])
                J[:, j] = (up[0] - dn[0]) / (2 * h)
            sign, logdet = np.linalg.slogdet(J)
            assert sign > 0
            assert abs(logdet - ld[i]) < 2e-5 * max(1.0, abs(ld[i])), (i, logdet, ld[i])


def test_latent_densities_against_scipy():
    """The latent log-densities (pzflow/distributions.py:165-194 CentBeta13, 255-275 Normal, 330-358 Tdist with the identity
    scale matrix, 414-441 Uniform) against scipy.stats -- an implementation that shares no code with the restatement."""
    import scipy.stats as st
    rng = np.random.default_rng(6)
    u = rng.uniform(-4.9, 4.9, size=(200, 3))
    want = st.beta.logpdf(u, 13, 13, loc=-5, scale=10).sum(axis=1)
    np.testing.assert_allclose(O.latent_log_prob(("CentBeta13", (3, 5)), u), want, rtol=1e-10, atol=1e-9)
    z = rng.normal(size=(200, 3))
    np.testing.assert_allclose(O.latent_log_prob(("Normal", (3,)), z), st.norm.logpdf(z).sum(axis=1), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(O.latent_log_prob(("Uniform", (3, 5)), u), np.full(200, -3 * np.log(10.0)), rtol=1e-12)
    outside = u.copy()
    outside[0, 1] = 5.5
    assert np.isneginf(O.latent_log_prob(("Uniform", (3, 5)), outside)[0])
    assert np.isneginf(O.latent_log_prob(("CentBeta13", (3, 5)), outside)[0])


# --------------------------------------------------------------------------- #
# (5) the posterior curves real JAX plotted                                     #
# --------------------------------------------------------------------------- #
def _notebook_curves():
    """Plain and marginalised redshift posteriors of the notebook's two samples, by the oracle: plain as cell 9 asks;
    marginalised by brute force over every combination of the rules' grids (cell 14), summed un-normalised with NaN -> 0,
    then normalised (pzflow/flow.py:560-664, 732-737)."""
    import itertools
    import plots as PL
    ex = F.example_flow()
    bi, li, p = ex["bijector_info"], ex["latent_info"], ex["params"]
    nb = golden("ref_shim_example_flow.npz")["nb_X"]
    grid = PL.NB_GRID
    plain = O.flow_posterior(bi, li, p, np.ascontiguousarray(nb[:, 1:]), 0, grid)
    out = []
    for row, missing in ((nb[0], [(1, PL.NB_U_GRID)]), (nb[1], [(1, PL.NB_U_GRID), (6, PL.NB_Y_GRID)])):
        combos = np.array(list(itertools.product(*[g for _, g in missing])), np.float32)
        rest = np.repeat(row[None, 1:], len(combos), axis=0).astype(np.float32)
        for q, (col, _) in enumerate(missing):
            rest[:, col - 1] = combos[:, q]
        tot = O.flow_posterior(bi, li, p, rest, 0, grid, normalize=False, nan_to_zero=True).sum(axis=0)
        out.append(tot / O.trapezoid(tot[None], grid)[0])
    return nb, plain, np.stack(out)


def check_against_the_notebook_figures(nb, plain, marg, tag=""):
    """Shared by the CPU (oracle) and the GPU (CUDA) test: every pixel of the blue / orange line of the figures the
    reference's authors made with real JAX (docs/tutorials/marginalization.ipynb cells 10 and 15) lies on the polyline
    the computed curves predict -- 2.5-pixel-wide anti-aliased lines, so <= 2.2 px from the centre line -- and the
    curve's vertices are all drawn.  One pixel is 0.8 % of the panel's tallest peak."""
    import plots as PL
    g = golden("nb_posterior_plots.npz")
    for panel in (0, 1):
        z = float(nb[panel, 0])
        reports = {"cell 10 blue": PL.overlay_report(g, 10, panel, z, [plain[panel]], 0, "blue"),
                   "cell 15 orange": PL.overlay_report(g, 15, panel, z, [plain[panel], marg[panel]], 1, "orange"),
                   "cell 15 blue": PL.overlay_report(g, 15, panel, z, [plain[panel], marg[panel]], 0, "blue")}
        for name, r in reports.items():
            print(f"FIGURE {tag} panel {panel} {name}: {r}")
            assert r["pixels"] > 20 and r["median"] < 0.8 and r["max"] < 2.2, (name, r)
        assert reports["cell 10 blue"]["vertices_hit"] >= 0.9 and reports["cell 15 orange"]["vertices_hit"] >= 0.9


def test_posterior_curves_overlay_the_figures_real_jax_plotted():
    """The forward path, log-density, posterior normalisation and the marginalisation rules against real JAX output: the
    notebook's figures.  Negative controls show what the check resolves: a 4 % lower peak, or the variants averaged
    with the wrong normalisation order, leave the drawn line."""
    import plots as PL
    nb, plain, marg = _notebook_curves()
    check_against_the_notebook_figures(nb, plain, marg, "oracle")
    g = golden("nb_posterior_plots.npz")
    # (a panel's y-axis follows its tallest curve, so a single curve is checked in SHAPE: relative to its peak)
    # (the first sample's posterior is a single spike on this grid; the second has two comparable points: 1 and 0.82)
    bent = plain[1].copy()
    bent[np.argsort(bent)[-2]] *= 0.95                     # the second-highest grid point 5 % lower
    assert PL.overlay_report(g, 10, 1, float(nb[1, 0]), [bent], 0, "blue")["max"] > 2.2
    # two curves share the axis: the marginalised posterior 5 % lower than it is leaves its drawn line
    assert PL.overlay_report(g, 15, 0, float(nb[0, 0]), [plain[0], 0.95 * marg[0]], 1, "orange")["max"] > 2.2
    shifted = np.roll(plain[1], 1)
    assert PL.overlay_report(g, 10, 1, float(nb[1, 0]), [shifted], 0, "blue")["max"] > 5
    assert PL.overlay_report(g, 15, 1, float(nb[1, 0]), [plain[1], plain[1]], 1, "orange")["max"] > 5   # not the plain curve

"""Behaviours the reference's own test-suite asserts for Flow / FlowEnsemble (pzflow tests/test_flow.py,
tests/test_ensemble.py), restated against the mirror package on the B200: shapes, error conventions, the training
loop's bookkeeping (loss lists, patience, divergence, validation, weights, condition scaling) and the error-sample
helper.  The reference exercises them on parameter-free flows -- ``Flow(columns, Reverse())`` -- and on the default
bijector; so does this file.  Values are covered by tests/test_gpu_parity.py; this file is about the API contract."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

XY = ("redshift", "y")


def _points():
    # the third point lies outside the default CentBeta13(2, 5) support: its log-probability is finfo.min (flow.py:406)
    return pd.DataFrame(np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]), columns=XY)


def _reverse_flow(**kw):
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Reverse
    return Flow(XY, Reverse(), **kw)


def test_training_bookkeeping_on_a_parameter_free_flow():
    """tests/test_flow.py: test_patience, test_train_no_errs_same, test_nan_train_stop, test_train_bad_inputs."""
    x = _points()
    losses = _reverse_flow().train(x, patience=2)
    # constant loss: epoch 1 sets the best value, epochs 2 and 3 fail to improve it -> initial + 3 entries; the
    # zero-probability point makes the loss huge (~1e38 / 3) but FINITE, so patience, not divergence, ends the run
    assert len(losses) == 4 and np.isfinite(losses).all() and losses[0] > 1e37 and len(set(losses)) == 1
    flow = _reverse_flow()
    assert np.allclose(flow.train(x, convolve_errs=True), flow.train(x, convolve_errs=False))
    nan_rows = pd.DataFrame(np.full((4, 2), np.nan), columns=["x", "y"])
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Reverse
    stopped = Flow(nan_rows.columns, Reverse()).train(nan_rows)
    assert len(stopped) == 2 and not np.isfinite(stopped).any()      # four finfo.min rows overflow the float32 mean
    for epochs in (-1, 2.4, "a"):
        with pytest.raises(ValueError):
            _reverse_flow().train(x, epochs=epochs)


def test_training_with_conditions_validation_weights_and_no_initial_loss():
    """tests/test_flow.py: test_train_w_conditions, test_validation_train, test_train_weights, test_no_initial_loss."""
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Reverse
    from pzflow_b200.distributions import Normal
    table = np.array([[1.0, 2.0, 0.1, 0.2], [3.0, 4.0, 0.3, 0.4], [5.0, 6.0, 0.5, 0.6]])
    frame = pd.DataFrame(table, columns=XY + ("a", "b"))
    flow = Flow(XY, Reverse(), latent=Normal(2), conditional_columns=("a", "b"))
    assert len(flow.train(frame, epochs=11)) == 12
    np.testing.assert_allclose(flow._condition_means, table[:, 2:].mean(axis=0), rtol=1e-6)   # flow.py:989-998
    np.testing.assert_allclose(flow._condition_stds, table[:, 2:].std(axis=0), rtol=1e-6)
    data = pd.DataFrame(np.random.default_rng(0).normal(size=(10, 2)), columns=("x", "y"))
    fit, held = data[:8], data[8:]
    for kw, n in ((dict(), 4), (dict(train_weight=np.linspace(1, 2, 8), val_weight=np.linspace(1, 2, 2)), 4),
                  (dict(initial_loss=False), 3)):
        both = Flow(fit.columns, Reverse()).train(fit, held, verbose=True, epochs=3, best_params=False, **kw)
        assert [len(part) for part in both] == [n, n]


def test_shapes_and_errors_of_the_public_calls():
    """tests/test_flow.py: test_returns_correct_shape, test_conditional_sample, test_bijector_not_set,
    test_latent_with_wrong_dimension, test_control_sample_randomness; tests/test_ensemble.py: shapes of every call."""
    from pzflow_b200 import Flow, FlowEnsemble
    from pzflow_b200.bijectors import Reverse
    from pzflow_b200.distributions import Normal, Uniform
    frame = pd.DataFrame(np.arange(12, dtype=float).reshape(-1, 4) / 12, columns=("x", "y", "a", "b"))
    grid = np.arange(0, 2.1, 0.12)
    for flow in (Flow(("x", "y"), Reverse(), latent=Normal(2)), Flow(("x", "y"), Reverse(), conditional_columns=("a", "b"))):
        assert flow.log_prob(frame).shape == (3,) and flow.log_prob(frame, err_samples=10).shape == (3,)
        assert flow.posterior(frame, column="x", grid=grid).shape == (3, grid.size)
        assert flow.posterior(frame, column="x", grid=grid, err_samples=10).shape == (3, grid.size)
        assert flow.posterior(frame.iloc[:1], column="y", grid=grid).shape == (1, grid.size)
    cond = Flow(("x", "y"), Reverse(), conditional_columns=("a", "b"))
    assert cond._get_conditions(frame).shape == (3, 2)
    with pytest.raises(ValueError):
        cond.sample(4)                                               # conditions are required (flow.py:777-780)
    assert cond.sample(4, conditions=frame).shape == (12, 4)
    assert cond.sample(4, conditions=frame, save_conditions=False).shape == (12, 2)
    plain = Flow(("x", "y"), Reverse(), latent=Uniform(2, 5))
    assert plain.sample(7, seed=3).equals(plain.sample(7, seed=3)) and not plain.sample(7, seed=3).equals(plain.sample(7, seed=4))
    unset = Flow(["x", "y"])
    with pytest.raises(ValueError):
        unset.sample(1)
    with pytest.raises(ValueError):
        unset.posterior(frame, column="x", grid=grid)
    with pytest.raises(ValueError):
        Flow(data_columns=["x", "y"], latent=Uniform(3), bijector=Reverse())
    ens = FlowEnsemble(("x", "y"), Reverse(), N=2)
    two = pd.DataFrame(np.arange(12, dtype=float).reshape(-1, 2) / 12, columns=("x", "y"))
    assert ens.log_prob(two).shape == (6,) and ens.log_prob(two, returnEnsemble=True).shape == (6, 2)
    assert ens.posterior(two, "x", grid).shape == (6, grid.size)
    assert ens.posterior(two, "x", grid, returnEnsemble=True).shape == (6, 2, grid.size)
    assert ens.sample(3, seed=0).shape == (3, 2) and ens.sample(3, seed=0, returnEnsemble=True).shape == (6, 2)
    assert len(ens.train(two, epochs=4)) == 2


def test_error_sample_helper_contract():
    """tests/test_flow.py::test_get_err_samples: [N * S, W] in row order, the skipped column removed, conditions scaled,
    a user's error model called as given, an unknown type refused."""
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Reverse
    noisy = pd.DataFrame(np.array([[1.0, 2.0, 0.1, 0.2], [3.0, 4.0, 0.3, 0.4]]), columns=("x", "y", "x_err", "y_err"))
    exact = pd.DataFrame(np.array([[1.0, 2.0, 0, 0]]), columns=("x", "y", "x_err", "y_err"))
    flow = Flow(("x", "y"), Reverse())
    assert flow._get_err_samples(0, noisy, 10).shape == (20, 2)
    np.testing.assert_allclose(flow._get_err_samples(0, exact, 10, skip="y"), np.ones((10, 1)))
    np.testing.assert_allclose(flow._get_err_samples(0, exact, 10, skip="x"), 2 * np.ones((10, 1)))
    on_y = Flow(("x"), Reverse(), conditional_columns=("y"))
    np.testing.assert_allclose(on_y._get_err_samples(0, exact, 10, type="conditions"), 2 * np.ones((10, 1)))
    with pytest.raises(ValueError):
        flow._get_err_samples(0, exact, 10, type="wrong")

    def shifted(key, X, Xerr, nsamples):
        return np.repeat(X + Xerr, nsamples, axis=0).reshape(X.shape[0], nsamples, X.shape[1])
    table = noisy.to_numpy()
    got = Flow(("x", "y"), Reverse(), data_error_model=shifted)._get_err_samples(0, noisy, 10)
    np.testing.assert_allclose(got, shifted(None, table[:, :2], table[:, 2:], 10).reshape(20, 2), rtol=1e-6)
    by_cond = Flow(("x"), Reverse(), conditional_columns=("y"), condition_error_model=shifted)
    np.testing.assert_allclose(by_cond._get_err_samples(0, noisy, 10, type="conditions"),
                               np.repeat(np.array([[2.2], [4.4]]), 10, axis=0), rtol=1e-6)


def test_default_bijector_trains_without_nans():
    """tests/test_flow.py::test_default_bijector: Flow(columns).train(data) builds ShiftBounds + RollingSplineCoupling."""
    from pzflow_b200 import Flow
    data = pd.DataFrame(np.random.default_rng(1).normal(size=(2000, 2)), columns=("x", "y"))
    flow = Flow(["x", "y"])
    losses = flow.train(data, epochs=3)
    assert len(losses) == 4 and np.isfinite(losses).all() and losses[-1] < losses[0]
    assert flow._bijector_info[0] == "Chain" and flow._bijector_info[1][0][0] == "ShiftBounds"


ROWS = np.array([[0.2, 0.1, -0.3, 0.5, 0.1, -0.4, -0.3],
                 [0.6, 0.5, 0.2, 0.2, -0.4, -0.1, 0.7],
                 [0.9, 0.2, -0.3, 0.3, 0.4, -0.4, -0.1]], np.float32)


def _battery():
    from pzflow_b200 import bijectors as B
    none = np.zeros((3, 1), np.float32)
    return [
        ("ColorTransform", lambda: B.ColorTransform(3, [1, 3, 5]), none),
        ("Reverse", lambda: B.Reverse(), none),
        ("Roll", lambda: B.Roll(2), none),
        ("Scale", lambda: B.Scale(2.0), none),
        ("Shuffle", lambda: B.Shuffle(), none),
        ("InvSoftplus", lambda: B.InvSoftplus(0), none),
        ("InvSoftplus-2", lambda: B.InvSoftplus([1, 3], [2.0, 12.0]), none),
        ("StandardScaler", lambda: B.StandardScaler(np.linspace(-1, 1, 7), np.linspace(1, 8, 7)), none),
        ("Chain", lambda: B.Chain(B.Reverse(), B.Scale(1 / 6), B.Roll(-1)), none),
        ("NeuralSplineCoupling", lambda: B.NeuralSplineCoupling(), none),
        ("NeuralSplineCoupling-args", lambda: B.NeuralSplineCoupling(16, 3, 2, 128, 3), np.arange(9, dtype=np.float32).reshape(3, 3)),
        ("RollingSplineCoupling", lambda: B.RollingSplineCoupling(2), none),
        ("RollingSplineCoupling-periodic", lambda: B.RollingSplineCoupling(2, 1, 16, 3, 2, 128, None, 0, True), none),
        ("ShiftBounds", lambda: B.ShiftBounds(-0.5, 0.9, 5), none),
        ("ShiftBounds-vector", lambda: B.ShiftBounds(-1 * np.ones(7), 1.1 * np.ones(7), 3), none),
    ]


@pytest.mark.parametrize("case", range(15))
def test_every_bijector_of_the_reference_battery_is_a_bijection(case):
    """tests/test_bijectors.py::TestBijectors (shapes, inverse(forward(x)) = x and log-det antisymmetry on its 3 x 7
    input) for every bijector the reference lists there -- Shuffle included -- through the Bijector protocol."""
    name, make, cond = _battery()[case]
    init_fun, info = make()
    params, forward_fun, inverse_fun = init_fun(0, ROWS.shape[-1])
    y, ld = forward_fun(params, ROWS, conditions=cond)
    assert y.shape == ROWS.shape and ld.shape == (3,), name
    xi, ldi = inverse_fun(params, ROWS, conditions=cond)
    assert xi.shape == ROWS.shape and ldi.shape == (3,), name
    back, ldb = inverse_fun(params, y, conditions=cond)
    np.testing.assert_allclose(back, ROWS, atol=3e-6, err_msg=name)
    np.testing.assert_allclose(ld, -ldb, atol=3e-6, err_msg=name)


def test_dequantizer_shape_bad_arguments_and_latent_contract():
    """tests/test_bijectors.py::test_uniform_dequantizer_returns_correct_shape and test_bad_inputs;
    tests/test_distributions.py (log_prob / sample shapes, seeds control the draws) for every latent."""
    from pzflow_b200 import bijectors as B
    from pzflow_b200 import distributions as Dn
    init_fun, _ = B.UniformDequantizer([1, 3, 4])
    params, forward_fun, _ = init_fun(0, 7)
    assert forward_fun(params, ROWS, conditions=np.zeros((3, 1), np.float32))[0].shape == ROWS.shape
    for make in (lambda: B.ColorTransform(0, [1, 2, 3, 4]), lambda: B.ColorTransform(1.3, [1, 2, 3, 4]),
                 lambda: B.ColorTransform(1, [2, 3, 4]), lambda: B.Roll(2.4), lambda: B.Scale(2), lambda: B.Scale(np.arange(7)),
                 lambda: B.InvSoftplus([0, 1, 2], [1.0, 2.0]),
                 lambda: B.RollingSplineCoupling(2, 1, 16, 3, 2, 128, None, 0, "fake"), lambda: B.ShiftBounds(4, 2, 1),
                 lambda: B.ShiftBounds(np.array([0, 1]), 2, 1)):
        with pytest.raises(ValueError):
            make()
    latents = [(Dn.CentBeta(2), ((0.0, 1.0), (2.0, 3.0))), (Dn.CentBeta13(2), ()), (Dn.Normal(2), ()),
               (Dn.Tdist(2), np.log(30.0)), (Dn.Uniform(2), ())]
    joint = Dn.Joint(Dn.Normal(1), Dn.Uniform(1, 3), Dn.CentBeta13(1))
    latents.append((joint, joint._params))
    for dist, p in latents:
        pts = np.full((3, dist.input_dim), 0.1, np.float32) * np.arange(1, 4, dtype=np.float32)[:, None]
        assert np.shape(dist.log_prob(p, pts)) == (3,)
        a, b, c = dist.sample(p, 4, 0), dist.sample(p, 4, 0), dist.sample(p, 4, 1)
        assert np.shape(a) == (4, dist.input_dim) and np.allclose(a, b) and not np.allclose(a, c)


def test_an_ensemble_is_its_flows_and_survives_a_file(tmp_path):
    """tests/test_ensemble.py (test_log_prob, test_posterior, test_sample, test_conditional_sample, test_train,
    test_load_ensemble, test_bad_inputs) and tests/test_flow.py::test_load_flow: flow i of FlowEnsemble(N) is
    Flow(seed=i); the ensemble rules are log(mean(exp)) and mean-then-normalise; sampling takes ceil(n / N) rows per
    flow; training flow i uses seed default_rng(0).integers(1e9, size=N)[i]; files round-trip and refuse the wrong class."""
    import dill
    from pzflow_b200 import Flow, FlowEnsemble
    from pzflow_b200.bijectors import Reverse, RollingSplineCoupling
    ens = FlowEnsemble(("x", "y"), RollingSplineCoupling(nlayers=2), N=2)
    members = [Flow(("x", "y"), RollingSplineCoupling(nlayers=2), seed=s) for s in (0, 1)]
    pts = pd.DataFrame(np.arange(6).reshape(3, 2) / 10, columns=("x", "y"))
    each = np.stack([m.log_prob(pts) for m in members], axis=1)
    np.testing.assert_allclose(ens.log_prob(pts, returnEnsemble=True), each, rtol=1e-6)
    np.testing.assert_allclose(ens.log_prob(pts), np.log(np.exp(each).mean(axis=1)), rtol=1e-6)
    grid = np.linspace(-1, 1, 5)
    per = np.stack([m.posterior(pts, "x", grid) for m in members], axis=1)
    np.testing.assert_allclose(ens.posterior(pts, "x", grid, returnEnsemble=True), per, rtol=1e-6)
    raw = np.mean([m.posterior(pts, "x", grid, normalize=False) for m in members], axis=0)
    np.testing.assert_allclose(ens.posterior(pts, "x", grid), raw / np.trapezoid(raw, grid, axis=1)[:, None], rtol=1e-5)
    drawn = ens.sample(10, seed=0).values
    stacked = np.vstack([m.sample(5, seed=0).values for m in members])
    np.testing.assert_allclose(drawn[drawn[:, 0].argsort()], stacked[stacked[:, 0].argsort()])
    np.testing.assert_allclose(ens.sample(10, seed=0, returnEnsemble=True).values,
                               np.vstack([m.sample(10, seed=0).values for m in members]))
    cond = FlowEnsemble(("x", "y"), RollingSplineCoupling(nlayers=2, n_conditions=2), conditional_columns=("a", "b"), N=2)
    one = pd.DataFrame(np.arange(2).reshape(-1, 2), columns=("a", "b"))
    five = pd.DataFrame(np.arange(10).reshape(-1, 2), columns=("a", "b"))
    assert cond.sample(nsamples=1, conditions=one, save_conditions=False).shape == (1, 2)
    assert cond.sample(nsamples=1, conditions=five, save_conditions=False).shape == (5, 2)
    assert cond.sample(nsamples=2, conditions=five, save_conditions=False).shape == (10, 2)
    assert cond.sample(nsamples=1, conditions=five, save_conditions=False, returnEnsemble=True).shape == (10, 2)
    data = pd.DataFrame(np.random.default_rng(0).normal(size=(100, 2)), columns=("x", "y"))
    path = tmp_path / "ensemble.pzflow.pkl"
    before = ens.sample(10, seed=0)
    ens.save(str(path))
    np.testing.assert_allclose(FlowEnsemble(file=str(path)).sample(10, seed=0).values, before.values)
    curves = ens.train(data, epochs=4, batch_size=50)
    seeds = np.random.default_rng(0).integers(1e9, size=2)
    for i, m in enumerate(members):
        np.testing.assert_allclose(curves[f"Flow {i}"], m.train(data, epochs=4, batch_size=50, seed=seeds[i]), rtol=1e-6)
    with open(path, "rb") as fh:
        stored = dill.load(fh)
    stored["class"] = "Flow"
    with open(path, "wb") as fh:
        dill.dump(stored, fh, recurse=True)
    with pytest.raises(TypeError):
        FlowEnsemble(file=str(path))
    for kw in (dict(), dict(bijector=Reverse()), dict(data_columns=("x", "y"), file="file"), dict(bijector=Reverse(), file="file"),
               dict(info="fake", file="file")):
        with pytest.raises(ValueError):
            FlowEnsemble(**kw)
    single = tmp_path / "flow.pzflow.pkl"
    Flow(("x", "y"), Reverse(), info=["random", 42]).save(str(single))
    loaded = Flow(file=str(single))
    ints = np.array([[1, 2], [3, 4]])
    np.testing.assert_allclose(loaded._forward(loaded._params, ints)[0], ints[:, ::-1])
    np.testing.assert_allclose(loaded._inverse(loaded._params, loaded._forward(loaded._params, ints)[0])[0], ints)
    assert loaded.info == ["random", 42]
    with open(single, "rb") as fh:
        stored = dill.load(fh)
    stored["class"] = "FlowEnsemble"
    with open(single, "wb") as fh:
        dill.dump(stored, fh, recurse=True)
    with pytest.raises(TypeError):
        Flow(file=str(single))

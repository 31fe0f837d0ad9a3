"""Host logic that needs no GPU: Bijector_Info flattening, the API mirror's error behaviour,
save/load, row sharding and the 2-rank (gloo) gather."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from oracle import fixtures as F
from oracle import nsf_oracle as O
from pzflow_b200 import Flow, FlowEnsemble, _native as N
from pzflow_b200 import bijectors as B
from pzflow_b200 import distributions as Dz
from pzflow_b200 import plan as P
from pzflow_b200 import sharding
from pzflow_b200.utils import build_bijector_from_info


def test_bijector_info_matches_reference_layout():
    """RollingSplineCoupling == Chain(*(NSC, Roll) * nlayers) (bijectors.py:651-665)."""
    _, info = B.RollingSplineCoupling(3, n_conditions=2)
    assert info == F.rolling_info(3, n_conditions=2)
    _, info = B.Chain(B.ShiftBounds(0.0, 1.0, 4.0), B.RollingSplineCoupling(2))
    assert info[0] == "Chain" and info[1][0][0] == "ShiftBounds" and info[1][1][0] == "Chain"
    rebuilt = build_bijector_from_info(info)[1]
    assert rebuilt[1][1] == info[1][1]


def test_init_params_have_the_stax_layout():
    init_fun, info = B.Chain(B.ShiftBounds(np.zeros(7), np.ones(7), 4.0), B.RollingSplineCoupling(7, n_conditions=2))
    params, fwd, inv = init_fun(0, 7)
    assert isinstance(params, list) and params[0] == () and len(params[1]) == 14
    nsc = params[1][0]
    assert [len(p) for p in nsc] == [2, 0, 2, 0, 2]
    assert nsc[0][0].shape == (5, 128) and nsc[2][0].shape == (128, 128) and nsc[4][0].shape == (128, 188)
    assert nsc[4][1].shape == (188,) and nsc[0][0].dtype == np.float32
    assert isinstance(fwd, B.ForwardFunction) and isinstance(inv, B.InverseFunction)


def test_flatten_brackets_and_stage_order():
    ex = F.example_flow()
    stages = P.flatten(ex["bijector_info"], ex["params"][1], 7)
    kinds = [s["kind"] for s in stages]
    assert kinds[:3] == [N.STAGE_CHAIN_BEGIN, N.STAGE_SHIFT_BOUNDS, N.STAGE_CHAIN_BEGIN]
    assert kinds[3:-2] == [N.STAGE_NSC, N.STAGE_ROLL] * 7
    assert kinds[-2:] == [N.STAGE_CHAIN_END, N.STAGE_CHAIN_END]
    nsc = stages[3]
    assert (nsc["K"], nsc["hidden_layers"], nsc["hidden_dim"], nsc["transformed_dim"]) == (16, 2, 128, -1)
    assert [w.shape for w in nsc["weights"]] == [(3, 128), (128,), (128, 128), (128,), (128, 188), (188,)]


def test_flatten_rejects_bad_shapes_and_foreign_bijectors():
    info = ("NeuralSplineCoupling", (16, 5, 2, 128, None, 0, False))
    params = O.synth_bijector_params(info, 7, np.random.default_rng(0))
    with pytest.raises(ValueError):
        P.flatten(info, params, 9)  # weights were built for input_dim 7
    # Shuffle draws its permutation from the key of the flow it belongs to: without one it cannot be flattened
    with pytest.raises(ValueError):
        P.flatten(("Shuffle", ()), (), 7)
    with pytest.raises(ValueError):
        P.resolve_shuffles(("Chain", (("Shuffle", ()),)), None, 7)
    with pytest.raises(NotImplementedError):
        P.flatten(("NotABijector", ()), (), 7)
    assert build_bijector_from_info(("UniformDequantizer", (0,)))[1] == ("UniformDequantizer", (0,))
    # the parameter-free bijectors flatten to their own stage kinds
    kinds = [s["kind"] for s in P.flatten(("Chain", (("Reverse", ()), ("Scale", (0.5,)), ("InvSoftplus", ([1, 3], [2.0, 12.0])),
                                                     ("ColorTransform", (3, [1, 3, 5])))), [(), (), (), ()], 7)]
    assert kinds == [P.N.STAGE_CHAIN_BEGIN, P.N.STAGE_REVERSE, P.N.STAGE_SCALE, P.N.STAGE_INV_SOFTPLUS,
                     P.N.STAGE_COLOR_TRANSFORM, P.N.STAGE_CHAIN_END]
    from pzflow_b200 import bijectors as B
    with pytest.raises(ValueError):
        B.ColorTransform(0, [0, 1])          # ref_idx must be positive (bijectors.py:230-231)
    with pytest.raises(ValueError):
        B.ColorTransform(2, [1, 3])          # ref_idx must be in mag_idx
    with pytest.raises(ValueError):
        B.InvSoftplus([0, 1], [1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        B.Scale(2)                           # scale must be a float
    assert B.Reverse()[1] == ("Reverse", ()) and B.Scale(2.0)[1] == ("Scale", (2.0,))


def test_constructor_validation_matches_reference():
    """pzflow/flow.py:127-149, 225-231; bijectors.py:446-447, 577-578, 744-751."""
    with pytest.raises(ValueError):
        Flow()
    with pytest.raises(ValueError):
        Flow(data_columns=("x", "y"), file="nope.pkl")
    with pytest.raises(ValueError):
        Flow(data_columns=("x", "y"), latent=Dz.Uniform(3, 5))
    with pytest.raises(ValueError):
        B.NeuralSplineCoupling(periodic="yes")
    with pytest.raises(ValueError):
        B.Roll(1.5)
    with pytest.raises(ValueError):
        B.ShiftBounds(np.zeros(3), np.ones(2))
    with pytest.raises(ValueError):
        B.ShiftBounds(1.0, 0.0)
    flow = Flow(data_columns=("x", "y"))
    with pytest.raises(ValueError):
        flow.log_prob(pd.DataFrame({"x": [0.0], "y": [0.0]}))  # bijector not set up
    with pytest.raises(ValueError):
        flow.sample(1)
    with pytest.raises(ValueError):
        FlowEnsemble()
    assert flow.latent.info == ("CentBeta13", (2, 5))  # default latent, flow.py:221-222
    # training needs the GPU like everything else: no CPU fallback (and no jax callables)
    from pzflow_b200._native import NativeError
    with pytest.raises((NativeError, NotImplementedError)):
        flow.train(pd.DataFrame({"x": [0.0], "y": [0.0]}), loss_fn=lambda *a: 0.0)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(NativeError):
            flow.train(pd.DataFrame({"x": [0.0], "y": [0.0]}))


def test_sample_argument_validation():
    flow = Flow(("x", "y"), B.RollingSplineCoupling(2))
    with pytest.raises(AssertionError):
        flow.sample(0)
    with pytest.raises(AssertionError):
        flow.sample(1.5)
    cflow = Flow(("x", "y"), B.RollingSplineCoupling(2, n_conditions=1), conditional_columns=("c",))
    with pytest.raises(ValueError):
        cflow.sample(2)


def test_save_load_roundtrip(tmp_path):
    """pzflow/flow.py:819-866, 152-197; flowEnsemble.py:490-505, 135-162."""
    flow = Flow(("x", "y"), B.Chain(B.ShiftBounds(np.zeros(2), np.ones(2), 4.0), B.RollingSplineCoupling(2)),
                info=["anything"])
    path = str(tmp_path / "flow.pzflow.pkl")
    flow.save(path)
    back = Flow(file=path)
    assert back.data_columns == flow.data_columns and back.info == ["anything"]
    assert back._latent_info == flow._latent_info
    assert back._bijector_info[1][1] == flow._bijector_info[1][1]
    np.testing.assert_array_equal(back._params[1][1][0][4][0], flow._params[1][1][0][4][0])
    with pytest.raises(TypeError):
        FlowEnsemble(file=path)
    ens = FlowEnsemble(("x", "y"), B.RollingSplineCoupling(2), N=2)
    epath = str(tmp_path / "ens.pkl")
    ens.save(epath)
    eback = FlowEnsemble(file=epath)
    assert list(eback._ensemble) == ["Flow 0", "Flow 1"]
    with pytest.raises(TypeError):
        Flow(file=epath)


def test_loads_reference_style_pickle(tmp_path):
    """A save dict with the reference's by-reference global pzflow.utils.gaussian_error_model."""
    import dill
    import types
    ex = F.example_flow()
    pkg, sub = types.ModuleType("pzflow"), types.ModuleType("pzflow.utils")

    def gaussian_error_model(key, X, Xerr, nsamples):
        return None
    gaussian_error_model.__module__ = "pzflow.utils"
    gaussian_error_model.__qualname__ = "gaussian_error_model"  # pickled by reference, like flow.py:238
    sub.gaussian_error_model = gaussian_error_model
    pkg.utils = sub
    sys.modules["pzflow"], sys.modules["pzflow.utils"] = pkg, sub
    try:
        d = {"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": None,
             "condition_means": None, "condition_stds": None, "data_error_model": gaussian_error_model,
             "condition_error_model": gaussian_error_model, "autoscale_conditions": True, "info": "x",
             "latent_info": ex["latent_info"], "bijector_info": ex["bijector_info"], "params": ex["params"]}
        path = str(tmp_path / "ref.pkl")
        with open(path, "wb") as fh:
            dill.dump(d, fh)
    finally:
        sys.modules.pop("pzflow"), sys.modules.pop("pzflow.utils")
    flow = Flow(file=path)
    assert flow._input_dim == 7 and flow.latent.info == ("Uniform", (7, 5))
    assert len(flow._params[1][1]) == 14


def test_loads_pickle_written_by_current_jax(tmp_path):
    """jax >= 0.4 pickles an Array as (jax._src.array._reconstruct_array, (np reconstructor, args, state, aval
    state)): the loader must hand back NumPy arrays without importing jax, and must recognise a pzflow-owned
    gaussian_error_model as the default error model."""
    import dill
    import types
    from pzflow_b200 import flow as flow_mod
    ex = F.example_flow()
    jax_pkg, jax_src, jax_arr = types.ModuleType("jax"), types.ModuleType("jax._src"), types.ModuleType("jax._src.array")

    def _reconstruct_array(fun, args, arr_state, aval_state):
        raise AssertionError("the jax reconstructor must not be called")
    _reconstruct_array.__module__, _reconstruct_array.__qualname__ = "jax._src.array", "_reconstruct_array"
    jax_arr._reconstruct_array = _reconstruct_array
    jax_pkg._src, jax_src.array = jax_src, jax_arr

    class FakeJaxArray:
        def __init__(self, a):
            self.a = np.asarray(a)

        def __reduce__(self):
            fun, args, state = self.a.__reduce__()
            return _reconstruct_array, (fun, args, state, {"weak_type": False})

    def wrap(t):
        if isinstance(t, (list, tuple)):
            return type(t)(wrap(q) for q in t)
        return FakeJaxArray(t) if isinstance(t, np.ndarray) else t

    pz_pkg, pz_utils = types.ModuleType("pzflow"), types.ModuleType("pzflow.utils")

    def gaussian_error_model(key, X, Xerr, nsamples):
        return None
    gaussian_error_model.__module__, gaussian_error_model.__qualname__ = "pzflow.utils", "gaussian_error_model"
    pz_utils.gaussian_error_model = gaussian_error_model
    pz_pkg.utils = pz_utils
    mods = {"jax": jax_pkg, "jax._src": jax_src, "jax._src.array": jax_arr, "pzflow": pz_pkg, "pzflow.utils": pz_utils}
    sys.modules.update(mods)
    try:
        d = {"class": "Flow", "data_columns": F.GALAXY_COLUMNS, "conditional_columns": ("c0",),
             "condition_means": FakeJaxArray(np.float32([1.5])), "condition_stds": FakeJaxArray(np.float32([2.0])),
             "data_error_model": gaussian_error_model, "condition_error_model": gaussian_error_model,
             "autoscale_conditions": True, "info": None, "latent_info": ex["latent_info"],
             "bijector_info": ex["bijector_info"], "params": wrap(ex["params"])}
        path = str(tmp_path / "jax.pkl")
        with open(path, "wb") as fh:
            dill.dump(d, fh)
    finally:
        for k in mods:
            sys.modules.pop(k)
    assert b"jax._src.array" in open(path, "rb").read()
    flow = Flow(file=path)
    W = flow._params[1][1][0][0][0]
    assert type(W) is np.ndarray and np.array_equal(W, ex["params"][1][1][0][0][0])
    assert type(flow._condition_means) is np.ndarray and float(flow._condition_stds[0]) == 2.0
    # the unpickled model is this package's gaussian_error_model; a foreign pzflow.utils one is accepted by name
    assert flow_mod._is_gaussian_error_model(flow.data_error_model)
    assert flow_mod._is_gaussian_error_model(gaussian_error_model)
    assert not flow_mod._is_gaussian_error_model(lambda *a: None)


@pytest.mark.skipif(not os.path.exists("/root/reference/pzflow/example_files/example-flow.pzflow.pkl"),
                    reason="the reference tree only exists in the build container")
def test_loads_the_shipped_reference_checkpoint():
    """pzflow/examples.py:105-113 loads this file with Flow(file=...): the same call here, no jax."""
    flow = Flow(file="/root/reference/pzflow/example_files/example-flow.pzflow.pkl")
    ex = F.example_flow()
    assert tuple(flow.data_columns) == tuple(F.GALAXY_COLUMNS) and flow._latent_info[0] == "Uniform"
    got, want = P._param_leaves(flow._params[1], []), P._param_leaves(ex["params"][1], [])
    assert len(got) == len(want) and all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(got, want))


def test_shard_rows_partition():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1003
    full = torch.arange(n, dtype=torch.float32)

    def fake_eval(rows):  # stands in for the per-rank kernel launch
        return rows * 2.0 + 1.0
    out = sharding.sharded_eval(fake_eval, full, gather=True)
    ok = torch.equal(out, full * 2.0 + 1.0)
    a, b = sharding.shard_rows(n, rank, world)
    local = sharding.sharded_eval(fake_eval, full, gather=False)
    ok = ok and torch.equal(local, full[a:b] * 2.0 + 1.0)
    t = sharding.max_over_ranks(float(rank + 1))
    ok = ok and t == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_sharded_eval_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_err_sample_arrays_follow_the_reference_rules():
    """Flow._get_err_arrays mirrors pzflow/flow.py:339-395: zeros for missing *_err columns, the
    skipped column removed, conditions (and their errors) standard-scaled."""
    import pandas as pd
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    flow = Flow(data_columns=("x", "y", "z"), conditional_columns=("a", "b"),
                bijector=Chain(ShiftBounds(np.zeros(3), np.ones(3), 4.0), RollingSplineCoupling(3, n_conditions=2)))
    flow._condition_means = np.array([1.0, -2.0], dtype=np.float32)
    flow._condition_stds = np.array([2.0, 0.5], dtype=np.float32)
    rng = np.random.default_rng(0)
    df = pd.DataFrame(rng.random((5, 5)), columns=["x", "y", "z", "a", "b"])
    df["y_err"] = 0.1
    df["b_err"] = 0.2
    X, Xerr = flow._get_err_arrays(df, "data")
    assert X.dtype == np.float32 and X.shape == (5, 3)
    np.testing.assert_allclose(X, df[["x", "y", "z"]].to_numpy(), rtol=1e-6)
    np.testing.assert_allclose(Xerr, np.array([[0.0, 0.1, 0.0]] * 5), rtol=1e-6)
    Xs, Xserr = flow._get_err_arrays(df, "data", skip="y")
    assert Xs.shape == (5, 2) and np.all(Xserr == 0.0)
    np.testing.assert_allclose(Xs, df[["x", "z"]].to_numpy(), rtol=1e-6)
    C, Cerr = flow._get_err_arrays(df, "conditions")
    np.testing.assert_allclose(C, (df[["a", "b"]].to_numpy() - [1.0, -2.0]) / [2.0, 0.5], rtol=1e-5)
    np.testing.assert_allclose(Cerr, np.array([[0.0, 0.4]] * 5), rtol=1e-6)
    with pytest.raises(ValueError):
        flow._get_err_arrays(df, "other")
    uncond = Flow(data_columns=("x", "y"), bijector=Chain(ShiftBounds(np.zeros(2), np.ones(2), 4.0),
                                                           RollingSplineCoupling(2)))
    assert uncond._get_err_arrays(df, "conditions") == (None, None)
    uncond.data_error_model = lambda key, X, Xerr, n: X
    with pytest.raises(NotImplementedError):
        uncond._get_err_arrays(df, "data")


# --------------------------------------------------------------------------- #
# training step: the differentiable restatement and DP   #
# --------------------------------------------------------------------------- #
def _train_case(dtype=np.float64):
    """A small conditional flow with every chain element the training path supports."""
    info = ("Chain", (("ColorTransform", (2, [1, 2, 3])), ("InvSoftplus", (0, 3.0)),
                      ("ShiftBounds", (np.full(4, -2.0, np.float32), np.full(4, 3.0, np.float32), 4.0)),
                      ("RollingSplineCoupling", (2, 1, 8, 5, 1, 16, None, 2)), ("Reverse", ()), ("Scale", (0.7,))))
    rng = np.random.default_rng(3)
    bparams = [(), (), (), O.synth_bijector_params(O.expand_rolling((2, 1, 8, 5, 1, 16, None, 2)), 4, rng, out_scale=1.5),
               (), ()]
    lat = tuple((float(a), float(b)) for a, b in rng.normal(0.9, 0.3, size=(4, 2)))
    X = rng.uniform(0.2, 1.2, size=(64, 4)).astype(dtype)
    C = rng.normal(size=(64, 2)).astype(dtype)
    return info, ("CentBeta", (4, 5)), (lat, bparams), X, C


def test_training_restatement_matches_oracle_and_its_gradient_is_right():
    import torch
    from pzflow_b200.train import TorchChain
    info, latent_info, params, X, C = _train_case()
    p64 = (params[0], O.cast_params(params[1], np.float64))
    chain = TorchChain(info, params[1], latent_info, params[0], 4, torch.device("cpu"), dtype=torch.float64)
    lp = chain.log_prob(torch.from_numpy(X), torch.from_numpy(C))
    want = O.flow_log_prob(info, latent_info, p64, X, C)
    np.testing.assert_allclose(lp.detach().numpy(), want, rtol=1e-9, atol=1e-9)
    # analytic gradient of the training loss vs central differences on a few leaves
    loss = -(lp).mean()
    loss.backward()
    rng = np.random.default_rng(0)
    for leaf in (chain.leaves[0], chain.leaves[-1], chain.leaves[len(chain.leaves) // 2]):
        flat = leaf.detach().reshape(-1)
        for j in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            g = leaf.grad.reshape(-1)[j].item()
            old = flat[j].item()
            vals = []
            for eps in (1e-6, -1e-6):
                with torch.no_grad():
                    leaf.reshape(-1)[j] = old + eps
                vals.append(-(chain.log_prob(torch.from_numpy(X), torch.from_numpy(C))).mean().item())
            with torch.no_grad():
                leaf.reshape(-1)[j] = old
            fd = (vals[0] - vals[1]) / 2e-6
            assert abs(g - fd) <= 1e-5 * max(1.0, abs(fd)), (g, fd)
    # the stax layout survives the round trip into Flow._params
    lat2, bp2 = chain.numpy_params()
    assert len(lat2) == 4 and isinstance(lat2[0][0], float)
    assert [np.asarray(l[0]).shape for l in bp2[3][0] if len(l)] == [np.asarray(l[0]).shape for l in params[1][3][0] if len(l)]


def _dp_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from pzflow_b200.train import TorchChain, dp_step, shard_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    info, latent_info, params, X, C = _train_case()
    chain = TorchChain(info, params[1], latent_info, params[0], 4, torch.device("cpu"), dtype=torch.float64)
    opt = torch.optim.Adam(chain.leaves, lr=1e-3)
    Xt, Ct, W = torch.from_numpy(X), torch.from_numpy(C), torch.linspace(0.5, 1.5, 64, dtype=torch.float64)
    for step in range(3):
        a, b = shard_batch(64, rank, world)
        dp_step(chain, opt, Xt[a:b], Ct[a:b], W[a:b], global_rows=64)
    q.put((rank, torch.cat([p.detach().reshape(-1) for p in chain.leaves]).numpy()))
    dist.destroy_process_group()


def test_data_parallel_step_equals_single_process_gloo():
    """Batch sharded over 2 ranks + one all_reduce of the flat gradient == the single-process step."""
    import torch
    import torch.multiprocessing as mp
    from pzflow_b200.train import TorchChain, dp_step
    info, latent_info, params, X, C = _train_case()
    chain = TorchChain(info, params[1], latent_info, params[0], 4, torch.device("cpu"), dtype=torch.float64)
    opt = torch.optim.Adam(chain.leaves, lr=1e-3)
    Xt, Ct, W = torch.from_numpy(X), torch.from_numpy(C), torch.linspace(0.5, 1.5, 64, dtype=torch.float64)
    start = torch.cat([p.detach().reshape(-1) for p in chain.leaves]).numpy().copy()
    for step in range(3):
        dp_step(chain, opt, Xt, Ct, W)
    single = torch.cat([p.detach().reshape(-1) for p in chain.leaves]).numpy()
    assert np.abs(single - start).max() > 1e-4  # it moved
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    np.testing.assert_array_equal(res[0], res[1])               # ranks stay bit-identical
    np.testing.assert_allclose(res[0], single, rtol=0, atol=1e-12)


def test_numa_binding_helper_is_inert_without_topology():
    """sharding.bind_to_gpu_numa_node: cpulist parsing, and no change of affinity when the topology is unreadable."""
    import os
    from pzflow_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert sharding.bind_to_gpu_numa_node(0) is None      # no CUDA device here
    assert os.sched_getaffinity(0) == before


def test_plan_cache_holds_its_params_and_sees_in_place_edits():
    """ids of freed arrays get reused and in-place edits keep the id -- neither may return a stale plan."""
    built = []

    def builder(tag):
        def build():
            built.append(tag)
            return tag
        return build

    cache = P.PlanCache(max_plans=2)
    for i in range(20):  # fresh trees in a loop: CPython reuses the freed arrays' ids
        params = [(), [(np.full((4, 4), i, np.float32), np.zeros(4, np.float32)), ()]]
        assert cache.get(params, 0, builder(i)) == i
    assert built == list(range(20)) and len(cache) == 2
    W = np.ones((8, 8), np.float32)
    params = [(W, np.zeros(8, np.float32))]
    assert cache.get(params, 0, builder("a")) == "a"
    assert cache.get(params, 0, builder("never")) == "a"          # same leaves, same bytes: hit
    assert cache.get([(W, params[0][1])], 0, builder("never")) == "a"  # a new container around the same leaves: hit
    assert cache.get(params, 1, builder("dev1")) == "dev1"        # another device: its own plan
    W *= 2.0
    assert cache.get(params, 0, builder("b")) == "b"              # edited in place: rebuilt
    W[3, 5] = 7.0
    assert cache.get(params, 0, builder("c")) == "c"
    assert cache.get(params, 0, builder("never")) == "c"
    assert "never" not in built


def test_training_latents_cover_tdist_and_joint():
    """Flow.train must be able to fit the latents Flow.log_prob evaluates (Tdist, Joint)."""
    import torch
    from pzflow_b200.train import TorchChain
    rng = np.random.default_rng(5)
    u = rng.uniform(-3, 3, size=(64, 5))
    info = ("Joint", ("Joint info", [("Uniform", (2, 5.0)), ("Tdist", (2,)), ("Normal", (1,))]))
    lat_params = [(), np.float64(np.log(12.0)), ()]
    chain = TorchChain(("Reverse", ()), (), info, lat_params, 5, torch.device("cpu"), dtype=torch.float64,
                       native_spline=False)
    got = chain._latent_any(info, lat_params, torch.as_tensor(u)).numpy()
    want = O.any_latent_log_prob(info, lat_params, u)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    t_info = ("Tdist", (5,))
    got = chain._latent_any(t_info, None, torch.as_tensor(u)).numpy()      # default nu = 30
    assert np.allclose(got, O.any_latent_log_prob(t_info, np.float64(np.log(30.0)), u), rtol=1e-12, atol=1e-12)


def test_ensemble_sample_composition_matches_reference_golden(monkeypatch):
    """FlowEnsemble.sample, unconditional (pzflow/flowEnsemble.py:397-408): given the per-flow draws, which rows of
    which flow end up where is pinned by the reference's own code (tests/golden/ref_shim_ensemble_sample.npz)."""
    from conftest import golden
    z = golden("ref_shim_ensemble_sample.npz")
    per_flow = z["per_flow"]                                   # [3 flows, ceil(10 / 3), 3 columns]
    cols = ("a", "b", "c")
    ens = FlowEnsemble(data_columns=cols, N=3)
    calls = []
    for i, flow in enumerate(ens._ensemble.values()):
        def fake(nsamples, conditions=None, save_conditions=True, seed=None, _i=i):
            calls.append((_i, nsamples, seed))
            return pd.DataFrame(per_flow[_i][:nsamples], columns=cols)
        monkeypatch.setattr(flow, "sample", fake)
    out = ens.sample(int(z["nsamples"]), seed=int(z["seed"]))
    assert calls == [(0, 4, 4), (1, 4, 4), (2, 4, 4)]          # ceil(10 / 3) from every flow, the same seed
    assert np.array_equal(out.to_numpy().astype(np.float32), z["sample"])
    assert np.array_equal(out.index.to_numpy(), z["sample_index"])
    both = ens.sample(4, seed=int(z["seed"]), returnEnsemble=True)
    assert tuple(both.shape) == tuple(z["return_ensemble_shape"])
    assert [k for k, _ in both.index] == list(z["return_ensemble_keys"])


def test_shard_bounds_cover_the_rows_in_blocks_of_four():
    for n in (0, 1, 3, 4, 5, 1000, 10_000_019):
        for parts in (1, 2, 3, 8):
            blocks = sharding.shard_bounds(n, parts)
            assert [a for a, _ in blocks] == sorted(a for a, _ in blocks)
            covered = 0
            for a, b in blocks:
                assert a == covered and b > a and a % 4 == 0
                covered = b
            assert covered == n
            if n >= 4 * parts:
                sizes = [b - a for a, b in blocks]
                assert len(blocks) == parts and max(sizes) - min(sizes) <= 8


def test_process_devices_policy(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 4)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 2)
    monkeypatch.delenv("PZFLOW_B200_DEVICES", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    sharding.set_devices(None)
    assert sharding.process_devices() == [0, 1, 2, 3]           # a plain process: every visible device
    monkeypatch.setenv("WORLD_SIZE", "4")
    assert sharding.process_devices() == [2]                    # one rank of torchrun owns its current device
    monkeypatch.setenv("PZFLOW_B200_DEVICES", "0,3")
    assert sharding.process_devices() == [0, 3]
    monkeypatch.setenv("PZFLOW_B200_DEVICES", "current")
    assert sharding.process_devices() == [2]
    sharding.set_devices("all")
    assert sharding.process_devices() == [0, 1, 2, 3]
    sharding.set_devices([1, 9])
    assert sharding.process_devices() == [1]
    sharding.set_devices(None)


# --------------------------------------------------------------------------- #
# jax-compatible key schedule (pzflow_b200.prng) and the stages that hang on it #
# --------------------------------------------------------------------------- #
def test_prng_matches_the_oracle_restatement_and_jax_documentation():
    """The host-side key schedule equals the oracle's restatement of jax.random (itself pinned by the Random123
    known answers and the notebook samples, tests/test_oracle.py); split(PRNGKey(0)) is the value jax's own
    documentation prints."""
    from oracle import jax_prng as J
    from pzflow_b200 import prng
    assert prng.layout() == "threefry_partitionable"     # the default: what the reference's required jax (>= 0.5.3) draws
    np.testing.assert_array_equal(prng.split(prng.key_from_seed(0), 2, "threefry"),
                                  [[4146024105, 967050713], [2718843009, 1272950319]])
    for seed in (0, 1, 7, 2 ** 31 + 5, 10 ** 17):
        k = prng.key_from_seed(seed)
        np.testing.assert_array_equal(k, J.PRNGKey(seed))
        for part in (False, True):
            lay = "threefry_partitionable" if part else "threefry"
            np.testing.assert_array_equal(prng.split(k, 3, lay), J.split(k, 3, part))
            for n in (1, 2, 7, 8, 33):
                np.testing.assert_array_equal(prng.words(k, n, part), J.random_bits(k, n, part))
                np.testing.assert_array_equal(prng.permutation(k, n, lay), J.permutation(k, n, part))
                assert sorted(prng.permutation(k, n, lay)) == list(range(n))


def test_shuffle_permutations_follow_the_chain_key_schedule(stream_layout):
    """plan.resolve_shuffles repeats Chain.init_fun's splits (pzflow/bijectors.py:146-149) and Shuffle's
    random.permutation (795-797): same result as the oracle's walk, different per position and per seed, and the
    flattened stage carries the permutation."""
    from oracle import jax_prng as J
    part = stream_layout == "threefry_partitionable"
    info = ("Chain", (("Shuffle", ()), ("Chain", (("Reverse", ()), ("Shuffle", ()))), ("Shuffle", ())))
    got = P.resolve_shuffles(info, 3, 7)
    want = O.resolve_shuffles(info, J.PRNGKey(3), 7, part)
    assert got == want
    perms = [got[1][0][1][0], got[1][1][1][1][1][0], got[1][2][1][0]]
    assert all(sorted(p) == list(range(7)) for p in perms) and len({tuple(p) for p in perms}) == 3
    assert P.resolve_shuffles(info, 4, 7) != got
    # a top-level Shuffle takes PRNGKey(seed) itself
    assert P.resolve_shuffles(("Shuffle", ()), 0, 7) == ("Shuffle", (tuple(int(v) for v in J.permutation(J.PRNGKey(0), 7, part)),))
    stages = P.flatten(got, [(), [(), ()], ()], 7)
    sh = [s for s in stages if s["kind"] == N.STAGE_SHUFFLE]
    assert [tuple(int(v) for v in s["vec_a"]) for s in sh] == [tuple(p) for p in perms]
    # resolved infos pass through unchanged; the bijector factory keeps the reference's Bijector_Info
    assert P.resolve_shuffles(got, None, 7) == got
    from pzflow_b200 import bijectors as B
    assert B.Shuffle()[1] == ("Shuffle", ())
    assert B.Chain(B.Shuffle(), B.UniformDequantizer([0, 2]))[1] == ("Chain", (("Shuffle", ()), ("UniformDequantizer", ([0, 2],))))
    kinds = [s["kind"] for s in P.flatten(("UniformDequantizer", (2,)), (), 7)]
    assert kinds == [N.STAGE_DEQUANTIZE]


def test_small_helpers_of_the_reference_utils_and_examples(monkeypatch, tmp_path):
    """pzflow/utils.py:230-243 (the reference's own test_sub_diag_indices_correct / _bad_input) and the example-data
    loaders (pzflow/examples.py:13-102): the data sets are not redistributed, the loaders say where they look."""
    import pandas as pd
    from pzflow_b200 import examples as E
    from pzflow_b200 import utils as U
    x = np.array([[[0, 0], [0, 0]], [[1, 1], [1, 1]], [[2, 2], [2, 2]]])
    y = np.array([[[1, 0], [0, 1]], [[2, 1], [1, 2]], [[3, 2], [2, 3]]])
    idx = U.sub_diag_indices(x)
    x[idx] = x[idx] + 1
    assert np.array_equal(x, y)
    for bad in (np.ones(2), np.ones((2, 2)), np.ones((2, 2, 2, 2))):
        with pytest.raises(ValueError):
            U.sub_diag_indices(bad)
    monkeypatch.setenv("PZFLOW_EXAMPLE_FILES", str(tmp_path))
    with pytest.raises(FileNotFoundError):
        E.get_galaxy_data()
    frame = pd.DataFrame({"x": [0.0, 1.0], "y": [2.0, 3.0]})
    frame.to_pickle(tmp_path / "two-moons-data.pkl")
    assert E.get_twomoons_data().equals(frame)
    init_fun, apply_fun = U.DenseReluNetwork(5, 2, 16)
    shape, params = init_fun(0, (10, 3))
    assert shape == (10, 5)
    assert [tuple(np.shape(q) for q in p) for p in params] == [((3, 16), (16,)), (), ((16, 16), (16,)), (), ((16, 5), (5,))]


def test_training_epochs_are_shuffled_like_the_reference(stream_layout):
    """pzflow/flow.py:1015, 1062-1064: key = PRNGKey(batch_seed); per epoch permute_key, sample_key, key = split(key, 3),
    idx = permutation(permute_key, n) -- prng.epoch_orders against the oracle's restatement of jax.random."""
    from oracle import jax_prng as J
    from pzflow_b200 import prng
    part = stream_layout == "threefry_partitionable"
    for seed, n in ((0, 10), (5, 1000), (2 ** 31 + 1, 77)):
        orders = prng.epoch_orders(seed, n)
        key = J.PRNGKey(seed)
        for _ in range(3):
            permute_key, _, key = J.split(key, 3, part)
            got = next(orders)
            np.testing.assert_array_equal(got, J.permutation(permute_key, n, part))
            assert sorted(got) == list(range(n))

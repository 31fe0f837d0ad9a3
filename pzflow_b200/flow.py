"""``Flow``: host-side mirror of pzflow/flow.py for the inference hot path.

Same constructor, attributes, ``_params`` pytree ``(latent_params,
bijector_params)``, ``_bijector_info``, save-dict keys and error behaviour as
the reference (pzflow/flow.py:49-251, 819-866); ``log_prob`` (409-464),
``posterior`` (466-738) and ``sample`` (740-817) evaluate through one fused CUDA
kernel per call via the C-ABI (``pzflow_b200.plan.Plan``).  There is no CPU
path: without a CUDA device every evaluation raises.

``err_samples`` (Gaussian error convolution, flow.py:339-395) is fused too: the samples are
drawn inside the kernel from a counter-based stream.  ``train`` lives in
``pzflow_b200/train.py``.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence, Tuple

import dill as pickle
import numpy as np
import pandas as pd
import torch

from . import distributions
from .bijectors import Bijector_Info, Chain, InitFunction, Pytree, RollingSplineCoupling, ShiftBounds
from . import sharding
from .plan import Plan, PlanCache, host_result
from .utils import build_bijector_from_info, gaussian_error_model


def _numpy_from_jax_pickle(fun, args, arr_state, aval_state=None):
    """What jax._src.array._reconstruct_array does before its device_put: rebuild the NumPy array that a
    jax Array's __reduce__ wrapped (jax >= 0.4 pickles an Array as (_reconstruct_array, (np_reduce...)))."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _JaxFreeUnpickler(pickle.Unpickler):
    """dill Unpickler for reference ``.pzflow.pkl`` files in a jax-less environment: jax Arrays come back as NumPy
    arrays and pzflow.utils globals (the save dict stores gaussian_error_model by reference, pzflow/flow.py:238)
    resolve to this package's equivalents."""

    def find_class(self, module, name):
        if module.split(".")[0] == "jax" and name == "_reconstruct_array":
            return _numpy_from_jax_pickle
        if module in ("pzflow.utils", "pzflow"):
            from . import utils as _utils
            if hasattr(_utils, name):
                return getattr(_utils, name)
        return super().find_class(module, name)


def _load_pickle(file: str) -> dict:
    """dill-load a reference ``.pzflow.pkl`` without jax or pzflow installed."""
    with open(file, "rb") as handle:
        return _JaxFreeUnpickler(handle).load()


def _is_gaussian_error_model(model) -> bool:
    """The default error model, whichever package it was imported from (a real pzflow install unpickles its own
    pzflow.utils.gaussian_error_model, which is not this module's object)."""
    return model is gaussian_error_model or (
        getattr(model, "__name__", None) == "gaussian_error_model"
        and str(getattr(model, "__module__", "")).split(".")[-1] == "utils")


def _pinned_f32(a: np.ndarray) -> torch.Tensor:
    """A C-contiguous float32 CPU tensor of a host array (one cast/copy pass, none if it already is one).  Ordinary
    pageable memory: the host-buffer entry points stage it through the plan's page-locked ring chunk by chunk,
    which is as fast as handing them page-locked arrays and saves page-locking a fresh array per call."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _to_numpy_tree(p):
    if isinstance(p, (list, tuple)):
        return type(p)(_to_numpy_tree(q) for q in p)
    if isinstance(p, (int, float)):
        return p
    return np.asarray(p)


class Flow:
    """A normalizing flow that models tabular data (inference on B200)."""

    def __init__(
        self,
        data_columns: Sequence[str] = None,
        bijector: Tuple[InitFunction, Bijector_Info] = None,
        latent: distributions.LatentDist = None,
        conditional_columns: Sequence[str] = None,
        data_error_model: Callable = None,
        condition_error_model: Callable = None,
        autoscale_conditions: bool = True,
        seed: int = 0,
        info: Any = None,
        file: str = None,
        _dictionary: dict = None,
    ) -> None:
        # validate parameters (pzflow/flow.py:126-149)
        if data_columns is None and file is None and _dictionary is None:
            raise ValueError("You must provide data_columns OR file.")
        if any((data_columns is not None, bijector is not None, conditional_columns is not None,
                latent is not None, data_error_model is not None, condition_error_model is not None,
                info is not None)):
            if file is not None:
                raise ValueError("If providing a file, please do not provide any other parameters.")
            if _dictionary is not None:
                raise ValueError("If providing a dictionary, please do not provide any other parameters.")
        if file is not None and _dictionary is not None:
            raise ValueError("Only provide file or _dictionary, not both.")

        self._plan_cache = PlanCache(1)
        if file is not None or _dictionary is not None:
            save_dict = self._save_dict()
            save_dict.update(_load_pickle(file) if file is not None else _dictionary)
            if save_dict["class"] != self.__class__.__name__:
                raise TypeError(f"This save file isn't a {self.__class__.__name__}. "
                                f"It is a {save_dict['class']}")
            self.data_columns = save_dict["data_columns"]
            self.conditional_columns = save_dict["conditional_columns"]
            self._input_dim = len(self.data_columns)
            self.info = save_dict["info"]
            self._latent_info = save_dict["latent_info"]
            self.latent = getattr(distributions, self._latent_info[0])(*self._latent_info[1])
            self.data_error_model = save_dict["data_error_model"]
            self.condition_error_model = save_dict["condition_error_model"]
            self._bijector_info = save_dict["bijector_info"]
            if self._bijector_info is not None:
                init_fun, _ = build_bijector_from_info(self._bijector_info)
                _, self._forward, self._inverse = init_fun(0, self._input_dim)   # PRNGKey(0), as flow.py:186-187
            self._bijector_seed = 0
            self._params = _to_numpy_tree(save_dict["params"])
            self._condition_means = save_dict["condition_means"]
            self._condition_stds = save_dict["condition_stds"]
            self._autoscale_conditions = save_dict["autoscale_conditions"]
        else:
            self.data_columns = tuple(data_columns)
            self._input_dim = len(self.data_columns)
            self.info = info
            if conditional_columns is None:
                self.conditional_columns = None
                self._condition_means = None
                self._condition_stds = None
            else:
                self.conditional_columns = tuple(conditional_columns)
                self._condition_means = np.zeros(len(self.conditional_columns), dtype=np.float32)
                self._condition_stds = np.ones(len(self.conditional_columns), dtype=np.float32)
            self._autoscale_conditions = autoscale_conditions
            if latent is None:
                self.latent = distributions.CentBeta13(self._input_dim, 5)
            else:
                self.latent = latent
            self._latent_info = self.latent.info
            if self.latent.input_dim != len(data_columns):
                raise ValueError(f"The latent distribution has {self.latent.input_dim} "
                                 f"dimensions, but data_columns has {len(data_columns)} "
                                 "dimensions. They must match!")
            self.data_error_model = gaussian_error_model if data_error_model is None else data_error_model
            self.condition_error_model = (gaussian_error_model if condition_error_model is None
                                          else condition_error_model)
            if bijector is not None:
                self.set_bijector(bijector, seed=seed)
            else:
                self._bijector_info = None

    # ------------------------------------------------------------------ #
    def _check_bijector(self) -> None:
        if self._bijector_info is None:
            raise ValueError("The bijector has not been set up yet! "
                             "You can do this by calling "
                             "flow.set_bijector(bijector, params), "
                             "or by calling train, in which case the default "
                             "bijector will be used.")

    def set_bijector(self, bijector: Tuple[InitFunction, Bijector_Info], params: Pytree = None,
                     seed: int = 0) -> None:
        """Set the bijector (pzflow/flow.py:263-294)."""
        init_fun, self._bijector_info = bijector
        bijector_params, self._forward, self._inverse = init_fun(seed, self._input_dim)
        self._bijector_seed = seed   # PRNGKey(seed) is what a Shuffle stage draws its permutation from (flow.py:286-287)
        bijector_params = params if params is not None else bijector_params
        self._params = (self.latent._params, bijector_params)
        self._plan_cache = PlanCache(1)

    def _set_default_bijector(self, inputs: pd.DataFrame, seed: int = 0) -> None:
        """ShiftBounds -> RollingSplineCoupling (pzflow/flow.py:296-322)."""
        data = inputs[list(self.data_columns)].to_numpy()
        mins = data.min(axis=0)
        maxs = data.max(axis=0)
        n_conditions = 0 if self.conditional_columns is None else len(self.conditional_columns)
        self.set_bijector(Chain(ShiftBounds(mins, maxs, 4.0),
                                RollingSplineCoupling(len(self.data_columns), n_conditions=n_conditions)),
                          seed=seed)

    # ------------------------------------------------------------------ #
    def _plan(self, params: Pytree = None, device: int = None) -> Plan:
        """The fused device plan of (bijector, latent) for ``params`` on ``device`` (default: the current device)."""
        params = self._params if params is None else params
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else -1
        self._plan_cache.max_plans = max(self._plan_cache.max_plans, torch.cuda.device_count() if device >= 0 else 1)
        return self._plan_cache.get(
            (params[0], params[1]), device,
            lambda: Plan(self._bijector_info, params[1], self._input_dim, self._latent_info, latent_params=params[0],
                         device=None if device < 0 else device, rng_key=getattr(self, "_bijector_seed", 0)))

    def _device_plans(self, n_rows: int) -> dict:
        """{device index: plan} a host-array call of ``n_rows`` rows (posterior: grid points) is sharded over inside
        this process (pzflow_b200.sharding.process_devices); small calls stay on the current device (one entry).
        The plans are resolved here, on the calling thread: the row blocks' worker threads only use them."""
        devs = sharding.process_devices()
        if len(devs) <= 1 or n_rows < sharding.MULTI_DEVICE_MIN_ROWS:
            return {}
        engine = self._plan().engine   # an explicit engine choice on the current device's plan carries over
        plans = {}
        for d in devs:
            plans[d] = self._plan(device=d)
            if plans[d].engine != engine:
                plans[d].set_engine(engine)
        return plans

    def _get_conditions(self, inputs: pd.DataFrame):
        """Standard-scaled conditions, or None for an unconditional flow -- the reference's
        zeros[N, 1] dummy is sliced away inside the bijector (pzflow/flow.py:324-337)."""
        if self.conditional_columns is None:
            return None
        columns = list(self.conditional_columns)
        conditions = inputs[columns].to_numpy().astype(np.float32)
        means = np.asarray(self._condition_means, dtype=np.float32)
        stds = np.asarray(self._condition_stds, dtype=np.float32)
        return (conditions - means) / stds

    def _log_prob(self, params: Pytree, inputs, conditions=None):
        """Log prob for arrays (pzflow/flow.py:397-407): forward, latent log-density +
        log_det, NaN -> zero probability -- one fused kernel.  Returns a device tensor."""
        return self._plan(params).log_prob(inputs, conditions)

    def log_prob(self, inputs: pd.DataFrame, err_samples: int = None, seed: int = None) -> np.ndarray:
        """Calculates log probability density of inputs (pzflow/flow.py:409-464)."""
        self._check_bijector()
        if err_samples is None:
            columns = list(self.data_columns)
            # float64 (or float32) columns go to the C call as they are: the down-cast (flow.py:442), the row gather and the
            # condition scaling (flow.py:330-336) run chunk by chunk inside the H2D | kernel | D2H pipeline
            cols = [inputs[c].to_numpy() for c in columns]
            for ctype in (np.float64, np.float32):
                if all(c.dtype == ctype and c.ndim == 1 for c in cols):
                    ccols = None
                    if self.conditional_columns is not None:
                        ccols = [inputs[c].to_numpy() for c in self.conditional_columns]
                        if not all(c.dtype == ctype and c.ndim == 1 for c in ccols):
                            ccols = [np.asarray(c, dtype=ctype) for c in ccols]
                    plans = self._device_plans(len(inputs))
                    if len(plans) <= 1:
                        return self._plan().log_prob_host_columns(cols, ccols, self._condition_means,
                                                                  self._condition_stds).numpy()
                    out = host_result((len(inputs),))

                    def block(dev, a, b):
                        plans[dev].log_prob_host_columns(
                            [c[a:b] for c in cols], None if ccols is None else [c[a:b] for c in ccols],
                            self._condition_means, self._condition_stds, out=out[a:b])
                    sharding.run_sharded(len(inputs), list(plans), block)
                    return out.numpy()
            X = _pinned_f32(inputs[columns].to_numpy())
            conditions = self._get_conditions(inputs)
            return self._log_prob(self._params, X, conditions).cpu().numpy()
        assert isinstance(err_samples, int), "err_samples must be a positive integer."
        assert err_samples > 0, "err_samples must be a positive integer."
        # Gaussian samples (pzflow/flow.py:455-464): drawn inside the kernel, reduced as log(mean(exp))
        seed = np.random.randint(1e18) if seed is None else seed
        if self._custom_error_models():
            # user error models (flow.py:339-395 accepts any callable): called on the host, their samples evaluated by the
            # fused kernels in sample-major order and reduced by log(mean(exp)) over the samples (flow.py:463-464)
            from .plan import ensemble_logmeanexp
            plan, n = self._plan(), len(inputs)
            out = np.empty(n, dtype=np.float32)
            Xs = self._err_samples_host(seed, inputs, err_samples, "data")          # ONE call of the model, as the
            Cs = self._err_samples_host(seed, inputs, err_samples, "conditions")    # reference makes it; [S, n, W]
            step = max(1, self._HOST_SAMPLE_ROWS // err_samples)
            for a in range(0, n, step):
                b = min(n, a + step)
                rows = np.ascontiguousarray(Xs[:, a:b]).reshape(err_samples * (b - a), -1)
                cond = None if Cs is None else torch.from_numpy(
                    np.ascontiguousarray(Cs[:, a:b]).reshape(err_samples * (b - a), -1)).to(plan.device)
                lp = plan.log_prob(torch.from_numpy(rows).to(plan.device), cond)   # device tensors: results stay on the device
                out[a:b] = ensemble_logmeanexp(lp.reshape(err_samples, b - a)).cpu().numpy()
            return out
        X, Xerr = self._get_err_arrays(inputs, "data")
        C, Cerr = self._get_err_arrays(inputs, "conditions")
        return self._plan().log_prob_err(X, Xerr, C, Cerr, err_samples, seed).cpu().numpy()

    def _get_err_samples(self, key, inputs: pd.DataFrame, err_samples: int, type: str = "data", skip: str = None) -> np.ndarray:
        """Draw error samples for each row of inputs (pzflow/flow.py:339-395): [N * err_samples, W] float32 in the
        reference's order (row i's samples are consecutive), skip column removed, conditions standard-scaled;
        zeros [N * err_samples, 1] for the conditions of an unconditional flow.  ``key``: an int seed or the uint32 pair
        jax.random.PRNGKey(seed) would be.  The fused log_prob / posterior never materialise this (the default Gaussian
        model is drawn inside the kernels); it exists for callers of the reference's helper."""
        if type == "conditions" and self.conditional_columns is None:
            return np.zeros((err_samples * inputs.shape[0], 1), dtype=np.float32)
        k = np.asarray(key).astype(np.uint64).ravel()
        seed = int(k[0]) if k.size == 1 else (int(k[0]) << 32) | int(k[1])
        s = self._err_samples_host(seed, inputs, err_samples, type, skip)          # [S, N, W]
        return np.ascontiguousarray(s.transpose(1, 0, 2)).reshape(inputs.shape[0] * err_samples, -1)

    _HOST_SAMPLE_ROWS = 1 << 22   # rows x samples evaluated per step of the custom-error-model route

    def _custom_error_models(self) -> bool:
        models = [self.data_error_model] + ([self.condition_error_model] if self.conditional_columns is not None else [])
        return any(m is not None and not _is_gaussian_error_model(m) for m in models)

    def _err_samples_host(self, seed, inputs: pd.DataFrame, err_samples: int, type: str = "data", skip: str = None):
        """Error samples of each row drawn by the flow's error-model CALLABLE on the host (pzflow/flow.py:339-395), for
        models other than the default Gaussian one: ``model(key, X[N, W], Xerr[N, W], nsamples) -> [N, nsamples, W]``
        with NumPy arrays and ``key`` = the uint32 pair jax.random.PRNGKey(seed) would be.  Returned float32 in
        SAMPLE-MAJOR order [nsamples, N, W] (sample s of input row i at [s, i]), skip column removed
        (flow.py:381-384), conditions standard-scaled (386-392); None for the conditions of an unconditional flow."""
        from . import prng
        if type == "data":
            columns, model = list(self.data_columns), self.data_error_model
        elif type == "conditions":
            if self.conditional_columns is None:
                return None
            columns, model = list(self.conditional_columns), self.condition_error_model
        else:
            raise ValueError("type must be `data` or `conditions`.")
        n = len(inputs)
        X = np.empty((n, len(columns)), dtype=np.float32)
        Xerr = np.zeros((n, len(columns)), dtype=np.float32)
        for i, col in enumerate(columns):
            if col == skip:                      # flow.py:370-373: the skipped column rides along as NaN
                X[:, i] = np.nan
                Xerr[:, i] = np.nan
                continue
            X[:, i] = inputs[col].to_numpy()
            if f"{col}_err" in inputs.columns:
                Xerr[:, i] = inputs[f"{col}_err"].to_numpy()
        model = gaussian_error_model if model is None else model
        samples = np.asarray(model(prng.key_from_seed(seed), X, Xerr, err_samples), dtype=np.float32)
        samples = samples.reshape(n, err_samples, len(columns))
        if skip is not None:
            samples = np.delete(samples, columns.index(skip), axis=2)
        if type == "conditions":
            samples = (samples - np.asarray(self._condition_means, np.float32)) / np.asarray(self._condition_stds, np.float32)
        return np.ascontiguousarray(samples.transpose(1, 0, 2), dtype=np.float32)

    def _get_err_arrays(self, inputs: pd.DataFrame, type: str = "data", skip: str = None):
        """Means and stds of the error samples for each row (pzflow/flow.py:339-395).  The reference draws
        the samples here; the fused kernels draw them on the fly, so this returns (X, Xerr) float32 with
        zeros for missing ``*_err`` columns (flow.py:366-369), the ``skip`` column removed (381-384) and
        conditions standard-scaled (386-392: (c + eps*e - mean)/std == (c - mean)/std + eps * e/std)."""
        if type == "data":
            columns = [c for c in self.data_columns if c != skip]
            model = self.data_error_model
        elif type == "conditions":
            if self.conditional_columns is None:
                return None, None
            columns = list(self.conditional_columns)
            model = self.condition_error_model
        else:
            raise ValueError("type must be `data` or `conditions`.")
        if model is not None and not _is_gaussian_error_model(model):
            raise NotImplementedError("only the default gaussian_error_model is drawn inside the B200 kernels; custom error "
                                      "models are supported by log_prob / posterior (host-side sampling), not together "
                                      "with marg_rules or convolve_errs")
        X = inputs[columns].to_numpy().astype(np.float32).reshape(len(inputs), len(columns))
        Xerr = np.zeros_like(X)
        for i, col in enumerate(columns):
            if f"{col}_err" in inputs.columns:
                Xerr[:, i] = inputs[f"{col}_err"].to_numpy().astype(np.float32)
        if type == "conditions":
            means = np.asarray(self._condition_means, dtype=np.float32)
            stds = np.asarray(self._condition_stds, dtype=np.float32)
            X = (X - means) / stds
            Xerr = Xerr / stds
        return X, Xerr

    def posterior(self, inputs: pd.DataFrame, column: str, grid, marg_rules: dict = None,
                  normalize: bool = True, err_samples: int = None, seed: int = None,
                  batch_size: int = None, nan_to_zero: bool = True) -> np.ndarray:
        """Calculates posterior distributions for the provided column
        (pzflow/flow.py:466-738).  The [rows x grid] expansion is generated inside the
        kernel; ``batch_size`` only bounds the size of one launch."""
        self._check_bijector()
        columns = list(self.data_columns)
        idx = columns.index(column)
        columns.remove(column)
        nrows = inputs.shape[0]
        batch_size = nrows if batch_size is None else batch_size
        inputs = inputs.reset_index(drop=True)
        if err_samples is not None:
            assert isinstance(err_samples, int), "err_samples must be a positive integer."
            assert err_samples > 0, "err_samples must be a positive integer."
            seed = np.random.randint(1e18) if seed is None else seed
        grid = np.asarray(grid, dtype=np.float32).reshape(-1)

        if marg_rules is not None:
            return self._posterior_marginalized(inputs, column, columns, idx, grid, marg_rules, normalize,
                                                nan_to_zero, err_samples, seed)

        if err_samples is not None:
            # flow.py:675-693, 722-724: samples of the other columns / conditions, each expanded over the grid,
            # probabilities averaged over the samples -- all inside the kernel
            plan = self._plan()
            if self._custom_error_models():
                # user error models: samples from the callables on the host, posteriors of every sample averaged
                # (flow.py:722-724), then normalised per galaxy (732-737)
                from .plan import ensemble_posterior_mean, normalize_pdfs_
                gdev = torch.from_numpy(grid).to(plan.device)
                pdfs = np.zeros((nrows, len(grid)), dtype=np.float32)
                if len(self.data_columns) == 1:
                    Rs = np.zeros((err_samples, nrows, 0), dtype=np.float32)
                else:
                    Rs = self._err_samples_host(seed, inputs, err_samples, "data", skip=column)   # one call of the model
                Cs = self._err_samples_host(seed, inputs, err_samples, "conditions")
                step = max(1, self._HOST_SAMPLE_ROWS // (err_samples * len(grid)))
                for a in range(0, nrows, step):
                    b = min(nrows, a + step)
                    rest = np.ascontiguousarray(Rs[:, a:b]).reshape(err_samples * (b - a), -1)
                    cond = None if Cs is None else torch.from_numpy(
                        np.ascontiguousarray(Cs[:, a:b]).reshape(err_samples * (b - a), -1)).to(plan.device)
                    un = plan.posterior(torch.from_numpy(rest).to(plan.device), idx, gdev, cond, normalize=False,
                                        nan_to_zero=False)
                    mean = ensemble_posterior_mean(un.reshape(err_samples, b - a, len(grid)), gdev, False)
                    if normalize or nan_to_zero:
                        normalize_pdfs_(mean, gdev, normalize, nan_to_zero)
                    pdfs[a:b] = mean.cpu().numpy()
                return pdfs
            C, Cerr = self._get_err_arrays(inputs, "conditions")
            if len(self.data_columns) == 1:
                rest = np.zeros((nrows, 0), dtype=np.float32)
                rest_err = None
            else:
                rest, rest_err = self._get_err_arrays(inputs, "data", skip=column)
            from . import prng
            if batch_size >= nrows or prng.layout() == "philox":
                pdfs = plan.posterior_err(rest, rest_err, idx, grid, C, Cerr, err_samples, seed,
                                          normalize=normalize, nan_to_zero=nan_to_zero)
                return pdfs.cpu().numpy()
            # the reference draws every batch's samples from the SAME key (flow.py:668-693: random.normal(key,
            # (len(batch), S, W)) per batch), so with a batch_size the stream restarts at every batch: evaluate the
            # batches as the separate draws they are
            pdfs = np.empty((nrows, len(grid)), dtype=np.float32)
            for a in range(0, nrows, batch_size):
                b = min(nrows, a + batch_size)
                pdfs[a:b] = plan.posterior_err(
                    rest[a:b], None if rest_err is None else rest_err[a:b], idx, grid, None if C is None else C[a:b],
                    None if Cerr is None else Cerr[a:b], err_samples, seed, normalize=normalize,
                    nan_to_zero=nan_to_zero).cpu().numpy()
            return pdfs

        # The reference loops over batch_size rows because it materialises [batch * G, D]
        # (pzflow/flow.py:697-730) and normalises once at the end (732-737).  Here the expansion is
        # generated inside the kernel and the host-buffer entry point cuts the galaxies into chunks
        # itself (normalisation is per galaxy, so chunking cannot change it): batch_size is accepted
        # for signature compatibility and has no effect on the result.
        plan = self._plan()
        conditions = self._get_conditions(inputs)
        rest = _pinned_f32(inputs[columns].to_numpy().reshape(nrows, -1))
        plans = self._device_plans(nrows * len(grid))
        if len(plans) <= 1:
            pdfs = plan.posterior_host(rest, idx, grid, conditions, normalize=normalize, nan_to_zero=nan_to_zero)
            return pdfs.numpy()
        pdfs = host_result((nrows, len(grid)))   # galaxies are independent (per-galaxy normalisation, flow.py:734)

        def block(dev, a, b):
            plans[dev].posterior_host(rest[a:b], idx, grid, None if conditions is None else conditions[a:b],
                                      normalize=normalize, nan_to_zero=nan_to_zero, out=pdfs[a:b])
        sharding.run_sharded(nrows, list(plans), block)
        return pdfs.numpy()

    def _posterior_marginalized(self, inputs, column, columns, idx, grid, marg_rules, normalize, nan_to_zero,
                                err_samples=None, seed=None):
        """Posteriors with missing values marginalised (pzflow/flow.py:560-664) as ONE extra expansion axis of the
        fused kernel instead of the reference's recursion through ``posterior`` on repeated rows.

        Rows are grouped by WHICH of the other data columns they miss.  A group's galaxies all get V variants -- the
        missing columns replaced by every combination of their marginal grids, the first rule's grid outermost, later
        rules evaluated on the row with the earlier columns already filled in, as the reference's nesting does
        (flow.py:610-643) -- and one call (``Plan.posterior_marg``) evaluates galaxy x variant x error sample x grid
        point on the device and sums the variants.  Rows that miss a column without a rule are never evaluated and
        keep the reference's zeros (flow.py:557).  Only the rule callables themselves run on the host."""
        from .plan import normalize_pdfs_
        plan = self._plan()
        nrows, G = inputs.shape[0], len(grid)
        flag = marg_rules["flag"]
        values = inputs[columns].to_numpy(dtype=np.float64).reshape(nrows, len(columns))
        missing = np.isnan(values) if np.isnan(flag) else np.isclose(values, flag)            # flow.py:563-574
        position = {c: j for j, c in enumerate(columns)}
        ruled = [name for name in marg_rules if name != "flag" and name in position]        # marginalisation order
        has_rule = np.zeros(len(columns), dtype=bool)
        has_rule[[position[c] for c in ruled]] = True
        evaluable = ~(missing & ~has_rule).any(axis=1)
        codes = (missing.astype(np.int64) << np.arange(len(columns), dtype=np.int64)).sum(axis=1)
        pdfs = torch.zeros((nrows, G), dtype=torch.float32, device=plan.device)
        all_columns = list(self.data_columns)
        for code in np.unique(codes[evaluable]) if nrows else []:
            rows = np.nonzero(evaluable & (codes == code))[0]
            frame = inputs.iloc[rows]
            names = [c for c in ruled if (int(code) >> position[c]) & 1]
            cond, cond_err = (self._get_err_arrays(frame, "conditions") if err_samples is not None
                              else (self._get_conditions(frame), None))
            if err_samples is not None and len(self.data_columns) > 1:
                rest, rest_err = self._get_err_arrays(frame, "data", skip=column)
            else:
                rest, rest_err = values[rows].astype(np.float32), None
            if not names:                                                                     # nothing missing
                if err_samples is None:
                    out = plan.posterior(torch.from_numpy(rest), idx, grid, cond, normalize=False, nan_to_zero=nan_to_zero)
                else:
                    out = plan.posterior_err(rest, rest_err, idx, grid, cond, cond_err, err_samples, seed,
                                             normalize=False, nan_to_zero=nan_to_zero)
                pdfs[torch.as_tensor(rows, device=plan.device)] = out
                continue
            if len(names) > 4:
                raise NotImplementedError("at most 4 columns of a row can be marginalised at once on the B200 path")
            # the variants: grids[q] holds, for every (row, combination of the earlier grids), the q-th rule's grid
            cur, levels = frame, []
            for q, name in enumerate(names):
                grids = np.asarray(cur.apply(marg_rules[name], axis=1, result_type="expand").to_numpy(), dtype=np.float64)
                M = grids.shape[1]
                levels = [np.repeat(v, M) for v in levels] + [grids.reshape(-1)]
                if q + 1 < len(names):      # later rules see the row with this column filled in (flow.py:617-623)
                    cur = pd.DataFrame(np.repeat(cur.to_numpy(), M, axis=0), columns=cur.columns)
                    cur[name] = grids.reshape(-1)
            V = levels[0].size // len(rows)
            mvals = np.stack(levels, axis=1).reshape(len(rows), V, len(names)).astype(np.float32)
            rest = np.where(missing[rows], np.float32(0), rest).astype(np.float32)          # never read by the kernel
            out = plan.posterior_marg(rest, idx, grid, [all_columns.index(c) for c in names], mvals, rest_err=rest_err,
                                      cond=cond, cond_err=cond_err, err_samples=err_samples or 0, seed=seed or 0,
                                      nan_to_zero=nan_to_zero)
            pdfs[torch.as_tensor(rows, device=plan.device)] = out
        if (normalize or nan_to_zero) and nrows:
            normalize_pdfs_(pdfs, torch.from_numpy(grid).to(plan.device), normalize, nan_to_zero)   # flow.py:732-737
        return pdfs.cpu().numpy()

    def sample(self, nsamples: int = 1, conditions: pd.DataFrame = None, save_conditions: bool = True,
               seed: int = None) -> pd.DataFrame:
        """Returns samples from the normalizing flow (pzflow/flow.py:740-817)."""
        self._check_bijector()
        assert isinstance(nsamples, int), "nsamples must be a positive integer."
        assert nsamples > 0, "nsamples must be a positive integer."
        if self.conditional_columns is not None and conditions is None:
            raise ValueError(f"Must provide the following conditions\n{self.conditional_columns}")

        if self.conditional_columns is None:
            cond = None
            n = nsamples
        else:
            conditions_idx = list(conditions.index)
            cond = self._get_conditions(conditions)
            conditions_idx = np.repeat(conditions_idx, nsamples)
            cond = np.repeat(cond, nsamples, axis=0)
            n = cond.shape[0]

        seed = int(np.random.randint(1e18)) if seed is None else int(seed)   # one seed for every row block
        with_params = isinstance(self.latent, (distributions.CentBeta, distributions.Tdist, distributions.Joint))

        def draw(dev, a, b):
            kw = dict(device=torch.device("cuda", dev), row_offset=a, total_rows=n)
            if with_params:
                return self.latent.sample_device(b - a, seed, self._params[0], **kw)  # latents with parameters
            return self.latent.sample_device(b - a, seed, **kw)  # draws stay on the device
        plans = self._device_plans(n)
        if len(plans) <= 1:
            plan = self._plan()
            x = self._inverse_to_host(plan, draw(plan.device_index, 0, n), cond)
        else:
            buf = np.empty((self._input_dim, n), dtype=np.float32)   # the frame's own column-major layout

            def block(dev, a, b):
                with torch.cuda.device(dev):
                    plans[dev].inverse_to_host_columns(draw(dev, a, b), None if cond is None else cond[a:b],
                                                       out=buf[:, a:b])
            sharding.run_sharded(n, list(plans), block)
            x = buf.T
        # x is a fresh array in the frame's own (column-major) layout: copy=False lets pandas wrap it as it is
        if self.conditional_columns is None:
            return pd.DataFrame(x, columns=self.data_columns, copy=False)
        if save_conditions:
            cond = cond * np.asarray(self._condition_stds, np.float32) + np.asarray(self._condition_means, np.float32)
            return pd.DataFrame(np.hstack((x, cond)),
                                columns=self.data_columns + self.conditional_columns).set_index(conditions_idx)
        return pd.DataFrame(x, columns=self.data_columns, copy=False).set_index(conditions_idx)

    @staticmethod
    def _inverse_to_host(plan: Plan, u: torch.Tensor, cond) -> np.ndarray:
        """x = inverse(u) for device-resident latent draws, returned as a NumPy array [N, D] in column-major memory
        (pzb_inverse_to_host_columns), which `pd.DataFrame` wraps without its transposing copy: the kernel of one
        chunk, the D2H copy of the previous one and the host-side scatter into the columns overlap."""
        return plan.inverse_to_host_columns(u, cond)

    # ------------------------------------------------------------------ #
    def _save_dict(self) -> dict:
        """Dictionary of everything ``save`` writes (pzflow/flow.py:819-845)."""
        save_dict = {"class": self.__class__.__name__}
        keys = ["data_columns", "conditional_columns", "condition_means", "condition_stds",
                "data_error_model", "condition_error_model", "autoscale_conditions", "info",
                "latent_info", "bijector_info", "params"]
        for key in keys:
            try:
                save_dict[key] = getattr(self, key)
            except AttributeError:
                try:
                    save_dict[key] = getattr(self, "_" + key)
                except AttributeError:
                    save_dict[key] = None
        return save_dict

    def save(self, file: str) -> None:
        """Saves the flow to a file (pzflow/flow.py:847-866)."""
        save_dict = self._save_dict()
        with open(file, "wb") as handle:
            pickle.dump(save_dict, handle, recurse=True)

    def train(self, inputs: pd.DataFrame, val_set: pd.DataFrame = None, train_weight: np.ndarray = None,
              val_weight: np.ndarray = None, epochs: int = 100, batch_size: int = 1024, optimizer: Callable = None,
              loss_fn: Callable = None, convolve_errs: bool = False, patience: int = None,
              best_params: bool = True, seed: int = 0, verbose: bool = False, progress_bar: bool = False,
              initial_loss: bool = True) -> list:
        """Trains the normalizing flow on the provided inputs (pzflow/flow.py:868-1159).

        The optional data-parallel training step of SURVEY.md §8f-4: gradients by torch autograd on the GPU
        (a library-level baseline, see pzflow_b200/train.py), every reported loss by the fused CUDA kernels,
        one all_reduce of the flat gradient per step when torch.distributed is initialised."""
        from .train import train_flow
        return train_flow(self, inputs, val_set=val_set, train_weight=train_weight, val_weight=val_weight,
                          epochs=epochs, batch_size=batch_size, optimizer=optimizer, loss_fn=loss_fn,
                          convolve_errs=convolve_errs, patience=patience, best_params=best_params, seed=seed,
                          verbose=verbose, progress_bar=progress_bar, initial_loss=initial_loss)

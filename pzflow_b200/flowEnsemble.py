"""``FlowEnsemble``: host-side mirror of pzflow/flowEnsemble.py for the hot path.

``log_prob`` (189-243), ``posterior`` (245-351) and ``sample`` (353-471) keep the
reference's signatures and reduction rules -- naive ``log(mean(exp(lp)))`` and
"average the un-normalised pdfs, then normalise" -- with every flow evaluated by
its fused kernel and the reductions done on the device.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence, Tuple

import dill as pickle
import numpy as np
import pandas as pd
import torch

from . import distributions, sharding
from .bijectors import Bijector_Info, InitFunction
from .flow import Flow, _load_pickle
from .plan import (ensemble_log_prob_host_columns, ensemble_logmeanexp, ensemble_posterior_host,
                   ensemble_posterior_mean, normalize_pdfs_, host_result)


def _frame_rows_to_device(inputs: pd.DataFrame, columns, device) -> torch.Tensor:
    """float32 rows [N, D] on the device from DataFrame columns: every column goes up as it is (a contiguous 1-D
    array) and the down-cast (pzflow/flow.py:442) and the row gather happen on the GPU -- the NumPy route
    (`to_numpy().astype(float32)`) is a strided 2-D copy that costs more than evaluating the flows."""
    import warnings
    with warnings.catch_warnings():
        # pandas hands out read-only column views; they are only read here (copied to the device)
        warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
        cols = [torch.from_numpy(np.ascontiguousarray(inputs[c].to_numpy())).to(device) for c in columns]
    return torch.stack([c.to(torch.float32) for c in cols], dim=1).contiguous()


class FlowEnsemble:
    """An ensemble of normalizing flows (inference on B200)."""

    def __init__(
        self,
        data_columns: Sequence[str] = None,
        bijector: Tuple[InitFunction, Bijector_Info] = None,
        latent: distributions.LatentDist = None,
        conditional_columns: Sequence[str] = None,
        data_error_model: Callable = None,
        condition_error_model: Callable = None,
        autoscale_conditions: bool = True,
        N: int = 1,
        info: Any = None,
        file: str = None,
    ) -> None:
        # validate parameters (pzflow/flowEnsemble.py:108-133)
        if data_columns is None and bijector is None and file is None:
            raise ValueError("You must provide data_columns and bijector OR file.")
        if file is not None and any((data_columns is not None, bijector is not None,
                                     conditional_columns is not None, latent is not None,
                                     data_error_model is not None, condition_error_model is not None,
                                     info is not None)):
            raise ValueError("If providing a file, please do not provide any other parameters.")

        if file is not None:
            save_dict = _load_pickle(file)
            c = save_dict.pop("class")
            if c != self.__class__.__name__:
                raise TypeError(f"This save file isn't a {self.__class__.__name__}. It is a {c}.")
            self._ensemble = {name: Flow(_dictionary=flow_dict)
                              for name, flow_dict in save_dict["ensemble"].items()}
            self.data_columns = save_dict["data_columns"]
            self.conditional_columns = save_dict["conditional_columns"]
            self.data_error_model = save_dict["data_error_model"]
            self.condition_error_model = save_dict["condition_error_model"]
            self.info = save_dict["info"]
            self._latent_info = save_dict["latent_info"]
            self.latent = getattr(distributions, self._latent_info[0])(*self._latent_info[1])
        else:
            self._ensemble = {
                f"Flow {i}": Flow(data_columns=data_columns, bijector=bijector,
                                  conditional_columns=conditional_columns, latent=latent,
                                  data_error_model=data_error_model,
                                  condition_error_model=condition_error_model,
                                  autoscale_conditions=autoscale_conditions, seed=i, info=f"Flow {i}")
                for i in range(N)
            }
            self.data_columns = data_columns
            self.conditional_columns = conditional_columns
            self.latent = self._ensemble["Flow 0"].latent
            self.data_error_model = data_error_model
            self.condition_error_model = condition_error_model
            self.info = info

    # ------------------------------------------------------------------ #
    def log_prob(self, inputs: pd.DataFrame, err_samples: int = None, seed: int = None,
                 returnEnsemble: bool = False) -> np.ndarray:
        """log probability density, [N] or [N, n_flows] (pzflow/flowEnsemble.py:189-243)."""
        flows = list(self._ensemble.values())
        if err_samples is not None:
            # every flow averages over its own error samples, then the ensemble rule (flowEnsemble.py:227-243)
            lps = np.stack([flow.log_prob(inputs, err_samples=err_samples, seed=seed) for flow in flows], axis=1)
            if returnEnsemble:
                return lps
            with np.errstate(divide="ignore"):
                return np.log(np.exp(lps).mean(axis=1, dtype=np.float32))
        columns = list(flows[0].data_columns)
        if not returnEnsemble:
            fused = self._log_prob_fused(flows, inputs, columns)
            if fused is not None:
                return fused
        X = None
        lps = []
        for flow in flows:
            flow._check_bijector()
            plan = flow._plan()
            if X is None:  # one host->device copy of the rows, shared by every flow
                X = _frame_rows_to_device(inputs, columns, plan.device)
            lps.append(flow._log_prob(flow._params, X, flow._get_conditions(inputs)))
        ensemble = torch.stack(lps, dim=0)  # [n_flows, N]
        if returnEnsemble:
            return ensemble.t().contiguous().cpu().numpy()
        return ensemble_logmeanexp(ensemble).cpu().numpy()

    @staticmethod
    def _log_prob_fused(flows, inputs, columns):
        """The ensemble rule inside the host pipeline (one staging of the rows for all flows), when the frame's
        columns are all float64 or all float32 and the flows scale their conditions alike; else None."""
        cols = [inputs[c].to_numpy() for c in columns]
        ctype = cols[0].dtype
        if ctype not in (np.float64, np.float32) or any(c.dtype != ctype or c.ndim != 1 for c in cols):
            return None
        ccols = means = stds = None
        cond_names = flows[0].conditional_columns
        if cond_names is not None:
            means, stds = flows[0]._condition_means, flows[0]._condition_stds
            for flow in flows[1:]:
                if flow.conditional_columns != cond_names or \
                        not np.array_equal(np.asarray(flow._condition_means), np.asarray(means)) or \
                        not np.array_equal(np.asarray(flow._condition_stds), np.asarray(stds)):
                    return None
            ccols = [np.asarray(inputs[c].to_numpy(), dtype=ctype) for c in cond_names]
        for flow in flows:
            flow._check_bijector()
        plans = [flow._plan() for flow in flows]
        if len({(p.device, p.input_dim, p.n_conditions) for p in plans}) != 1:
            return None
        devices = flows[0]._device_plans(len(inputs))
        if len(devices) > 1:   # one process, several GPUs: every device evaluates all flows on its block of rows
            per_dev = {d: [flow._plan(device=d) for flow in flows] for d in devices}
            out = host_result((len(inputs),))

            def block(dev, a, b):
                ensemble_log_prob_host_columns(per_dev[dev], [c[a:b] for c in cols],
                                               None if ccols is None else [c[a:b] for c in ccols], means, stds, out=out[a:b])
            sharding.run_sharded(len(inputs), list(devices), block)
            return out.numpy()
        return ensemble_log_prob_host_columns(plans, cols, ccols, means, stds).numpy()

    def posterior(self, inputs: pd.DataFrame, column: str, grid, marg_rules: dict = None,
                  normalize: bool = True, err_samples: int = None, seed: int = None,
                  batch_size: int = None, returnEnsemble: bool = False,
                  nan_to_zero: bool = True) -> np.ndarray:
        """posterior pdfs, [N, G] or [N, n_flows, G] (pzflow/flowEnsemble.py:245-351)."""
        grid = np.asarray(grid, dtype=np.float32).reshape(-1)
        flows = list(self._ensemble.values())
        if marg_rules is None and err_samples is None and not returnEnsemble:
            fused = self._posterior_fused(flows, inputs, column, grid, normalize, nan_to_zero)
            if fused is not None:
                return fused
        per_flow = [flow.posterior(inputs=inputs, column=column, grid=grid, marg_rules=marg_rules,
                                   err_samples=err_samples, seed=seed, batch_size=batch_size,
                                   normalize=False, nan_to_zero=nan_to_zero)
                    for flow in self._ensemble.values()]
        dev = next(iter(self._ensemble.values()))._plan().device
        ensemble = torch.from_numpy(np.stack(per_flow, axis=0)).to(dev)  # [n_flows, N, G]
        grid_t = torch.from_numpy(grid).to(dev)
        if returnEnsemble:
            nf, n, g = ensemble.shape
            if normalize:
                flat = ensemble.reshape(nf * n, g).contiguous()
                normalize_pdfs_(flat, grid_t, True, False)
                ensemble = flat.reshape(nf, n, g)
            return ensemble.permute(1, 0, 2).contiguous().cpu().numpy()
        return ensemble_posterior_mean(ensemble, grid_t, normalize).cpu().numpy()

    @staticmethod
    def _posterior_fused(flows, inputs, column, grid, normalize, nan_to_zero):
        """Mean of the flows' pdfs and its normalisation inside the host pipeline (galaxies staged once, only the
        [N, G] result comes back), when the flows scale their conditions alike; else None."""
        f0 = flows[0]
        for flow in flows:
            flow._check_bijector()
        if column not in f0.data_columns:
            return None                      # Flow.posterior raises the reference's error on the per-flow route
        for flow in flows[1:]:
            if flow.data_columns != f0.data_columns or flow.conditional_columns != f0.conditional_columns:
                return None
            if f0.conditional_columns is not None and (
                    not np.array_equal(np.asarray(flow._condition_means), np.asarray(f0._condition_means)) or
                    not np.array_equal(np.asarray(flow._condition_stds), np.asarray(f0._condition_stds))):
                return None
        plans = [flow._plan() for flow in flows]
        if len({(p.device, p.input_dim, p.n_conditions) for p in plans}) != 1:
            return None
        columns = list(f0.data_columns)
        idx = columns.index(column)
        columns.remove(column)
        nrows = inputs.shape[0]
        rest = np.ascontiguousarray(inputs[columns].to_numpy().reshape(nrows, -1), dtype=np.float32)
        cond = f0._get_conditions(inputs)
        devices = f0._device_plans(nrows * len(grid))
        if len(devices) > 1:
            per_dev = {d: [flow._plan(device=d) for flow in flows] for d in devices}
            out = host_result((nrows, len(grid)))

            def block(dev, a, b):
                ensemble_posterior_host(per_dev[dev], rest[a:b], idx, grid, None if cond is None else cond[a:b],
                                        normalize=normalize, nan_to_zero=nan_to_zero, out=out[a:b])
            sharding.run_sharded(nrows, list(devices), block)
            return out.numpy()
        return ensemble_posterior_host(plans, rest, idx, grid, cond, normalize=normalize, nan_to_zero=nan_to_zero).numpy()

    def sample(self, nsamples: int = 1, conditions: pd.DataFrame = None, save_conditions: bool = True,
               seed: int = None, returnEnsemble: bool = False) -> pd.DataFrame:
        """samples from the ensemble (pzflow/flowEnsemble.py:353-471)."""
        if returnEnsemble:
            return pd.concat([flow.sample(nsamples, conditions, save_conditions, seed)
                              for flow in self._ensemble.values()], keys=self._ensemble.keys())
        if conditions is None:
            N = int(np.ceil(nsamples / len(self._ensemble)))
            samples = pd.concat([flow.sample(N, conditions, save_conditions, seed)
                                 for flow in self._ensemble.values()])
            return samples.sample(nsamples, random_state=seed).reset_index(drop=True)
        # Conditional draws (pzflow/flowEnsemble.py:409-471): every condition row (repeated nsamples times) is handed
        # to ONE flow.  Which flow is decided on row POSITIONS here, one Flow.sample call per flow; the draws the
        # reference makes to decide -- pandas' shuffle with random_state = seed / 1e9, jax.random.permutation of the
        # chunks under PRNGKey(seed), NumPy's choice / integers for fewer rows than flows -- are the same ones.
        from . import prng
        members = list(self._ensemble.values())
        rows = pd.concat([conditions] * nsamples) if nsamples > 1 else conditions
        seed = np.random.randint(1e18) if seed is None else seed
        n_rows, n_flows = rows.shape[0], len(members)
        if n_rows > n_flows:
            shuffled = pd.Series(np.arange(n_rows)).sample(frac=1.0, random_state=int(seed / 1e9)).to_numpy()
            parts = np.array_split(shuffled, n_flows)                    # positions, ~equal parts in shuffled order
            turn = prng.permutation(prng.key_from_seed(seed), n_flows)   # flow i draws part turn[i]
            pieces = []
            for flow, part in zip(members, (parts[t] for t in turn)):
                block = rows.iloc[part]
                pieces.append(flow.sample(1, block, save_conditions, seed).set_index(block.index))
            return pd.concat(pieces).sort_index()
        picker = np.random.default_rng(seed)
        chosen = picker.choice(n_flows, size=n_rows, replace=True)
        row_seeds = picker.integers(1e18, size=n_rows)
        return pd.concat([members[f].sample(1, rows[i:i + 1], save_conditions, int(sd))
                          for i, (f, sd) in enumerate(zip(chosen, row_seeds))]).set_index(rows.index)

    # ------------------------------------------------------------------ #
    def save(self, file: str) -> None:
        """Saves the ensemble to a file (pzflow/flowEnsemble.py:473-505)."""
        save_dict = {
            "data_columns": self.data_columns,
            "conditional_columns": self.conditional_columns,
            "latent_info": self.latent.info,
            "data_error_model": self.data_error_model,
            "condition_error_model": self.condition_error_model,
            "info": self.info,
            "class": self.__class__.__name__,
            "ensemble": {name: flow._save_dict() for name, flow in self._ensemble.items()},
        }
        with open(file, "wb") as handle:
            pickle.dump(save_dict, handle, recurse=True)

    def train(self, inputs: pd.DataFrame, val_set: pd.DataFrame = None, train_weight: np.ndarray = None,
              val_weight: np.ndarray = None, epochs: int = 50, batch_size: int = 1024, optimizer=None, loss_fn=None,
              convolve_errs: bool = False, patience: int = None, best_params: bool = True, seed: int = 0,
              verbose: bool = False, progress_bar: bool = False, initial_loss: bool = True) -> dict:
        """Trains the normalizing flows on the provided inputs (pzflow/flowEnsemble.py:507-609):
        one seed per flow, dictionary of per-flow losses."""
        rng = np.random.default_rng(seed)
        seeds = rng.integers(1e9, size=len(self._ensemble))
        loss_dict = dict()
        for i, (name, flow) in enumerate(self._ensemble.items()):
            if verbose:
                print(name)
            loss_dict[name] = flow.train(inputs=inputs, val_set=val_set, train_weight=train_weight,
                                         val_weight=val_weight, epochs=epochs, batch_size=batch_size,
                                         optimizer=optimizer, loss_fn=loss_fn, convolve_errs=convolve_errs,
                                         patience=patience, best_params=best_params, seed=int(seeds[i]),
                                         verbose=verbose, progress_bar=progress_bar, initial_loss=initial_loss)
        return loss_dict

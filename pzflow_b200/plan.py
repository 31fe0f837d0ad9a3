"""Plan compiler: Bijector_Info + params pytree -> device plan (pzb_plan).

Walks the nested ``("Name", (args...))`` tuples the reference stores in
``Flow._bijector_info`` (pzflow/bijectors.py:12-13, rebuilt by
pzflow/utils.py:11-19) in lock-step with the params pytree
(``[stage_params, ...]`` per Chain, stax ``[(W, b), (), ...]`` per
NeuralSplineCoupling; pzflow/bijectors.py:146-153, pzflow/utils.py:45-48) and
flattens them into the ``pzb_stage_desc`` list ``pzb_plan_create`` takes.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import prng

HOT_PATH_BIJECTORS = ("Chain", "ColorTransform", "InvSoftplus", "NeuralSplineCoupling", "Reverse", "Roll",
                      "RollingSplineCoupling", "Scale", "ShiftBounds", "Shuffle", "StandardScaler", "UniformDequantizer")
LATENTS = {"Uniform": N.LATENT_UNIFORM, "CentBeta13": N.LATENT_CENTBETA13, "Normal": N.LATENT_NORMAL,
           "CentBeta": N.LATENT_CENTBETA, "Tdist": N.LATENT_TDIST, "Joint": N.LATENT_JOINT}


def latent_segments(latent_info, latent_params) -> list:
    """(kind, dim, B, flat float32 params or None) for every segment of a latent's ``info`` / ``_params``
    (pzflow/distributions.py: a Joint's info is ("Joint", ("Joint info", [sub infos])), its params a list)."""
    name, args = latent_info
    if name == "Joint":
        subs = args[1]
        params = latent_params if latent_params is not None else [None] * len(subs)
        out = []
        for sub, p in zip(subs, params):
            out += latent_segments(tuple(sub), p)
        return out
    if name not in LATENTS:
        raise NotImplementedError(
            f"latent {name!r} is outside the B200 hot path (supported: {', '.join(LATENTS)})")
    dim = int(args[0])
    B = float(args[1]) if len(args) > 1 and name in ("Uniform", "CentBeta13", "CentBeta") else 5.0
    flat = None
    if name == "CentBeta" and latent_params is not None and len(latent_params):
        flat = np.asarray([[float(np.asarray(a)), float(np.asarray(b))] for a, b in latent_params], np.float32).ravel()
    if name == "Tdist" and latent_params is not None and np.size(latent_params):
        flat = np.asarray([float(np.asarray(latent_params))], np.float32)
    return [(LATENTS[name], dim, B, None if flat is None else np.ascontiguousarray(flat))]


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def resolve_shuffles(info: Tuple[str, tuple], key, input_dim: int) -> Tuple[str, tuple]:
    """``info`` with every ``("Shuffle", ())`` given the permutation the reference would draw for it.

    A Shuffle's permutation is not in its Bijector_Info: the reference draws it inside ``init_fun`` from the key the
    enclosing Chains hand down (pzflow/bijectors.py:795-797; Chain: ``rng, layer_rng = random.split(rng)`` per
    stage, 146-149) -- ``PRNGKey(seed)`` at construction (flow.py:286-287), ``PRNGKey(0)`` when a flow is loaded from
    a file (186-187).  This walk repeats that key schedule (pzflow_b200.prng) and returns ``("Shuffle", (perm,))``
    nodes, the form ``flatten`` and the training chain take; everything else is returned as it is."""
    name, args = info
    if name == "Shuffle":
        if len(args) == 1:
            return info
        if key is None:
            raise ValueError("Shuffle draws its permutation from the key of the flow it belongs to: build it through "
                             "Flow(..., seed=...) / set_bijector, or an InitFunction called with a seed")
        return ("Shuffle", (tuple(int(v) for v in prng.permutation(prng.key_from_seed(key), input_dim)),))
    if name == "RollingSplineCoupling" or name != "Chain":
        return info  # a RollingSplineCoupling is a Chain of NeuralSplineCoupling / Roll only: nothing to resolve
    out = []
    rng = None if key is None else prng.key_from_seed(key)
    for sub in args:
        layer = None
        if rng is not None:
            rng, layer = prng.split(rng)
        out.append(resolve_shuffles(sub, layer, input_dim))
    return ("Chain", tuple(out))


def flatten(info: Tuple[str, tuple], params: Any, input_dim: int) -> List[dict]:
    """Flatten (Bijector_Info, params) into stage dictionaries, Chains bracketed."""
    name, args = info
    if name == "Chain":
        if len(params) != len(args):
            raise ValueError("Chain params do not match its Bijector_Info")
        out = [dict(kind=N.STAGE_CHAIN_BEGIN)]
        for sub, p in zip(args, params):
            out += flatten(sub, p, input_dim)
        out.append(dict(kind=N.STAGE_CHAIN_END))
        return out
    if name == "Roll":
        (shift,) = args
        return [dict(kind=N.STAGE_ROLL, shift=int(shift))]
    if name == "ShiftBounds":
        mn, mx = np.atleast_1d(_f32(args[0])), np.atleast_1d(_f32(args[1]))
        B = float(args[2]) if len(args) > 2 else 5.0
        if len(mn) != len(mx):
            raise ValueError("Lengths of min and max do not match. Please provide either a single "
                             "min and max, or a min and max for each dimension.")
        return [dict(kind=N.STAGE_SHIFT_BOUNDS, vec_a=mn, vec_b=mx, B=B)]
    if name == "StandardScaler":
        return [dict(kind=N.STAGE_STANDARD_SCALER, vec_a=np.atleast_1d(_f32(args[0])),
                     vec_b=np.atleast_1d(_f32(args[1])), B=0.0)]
    if name == "Reverse":
        return [dict(kind=N.STAGE_REVERSE)]
    if name == "Shuffle":
        if len(args) != 1:
            raise ValueError("Shuffle without its permutation: resolve_shuffles(info, key, input_dim) first")
        perm = np.ascontiguousarray(np.asarray(args[0], dtype=np.float32).ravel())
        return [dict(kind=N.STAGE_SHUFFLE, vec_a=perm, vec_b=perm)]
    if name == "UniformDequantizer":
        idx = np.ascontiguousarray(np.atleast_1d(np.asarray(args[0])).astype(np.float32).ravel())
        return [dict(kind=N.STAGE_DEQUANTIZE, vec_a=idx, vec_b=idx)]
    if name == "Scale":
        (scale,) = args
        if np.ndim(scale) != 0:
            raise NotImplementedError("Scale with an array of scales is outside the B200 hot path")
        return [dict(kind=N.STAGE_SCALE, B=float(scale))]
    if name == "InvSoftplus":
        idx = np.atleast_1d(np.asarray(args[0])).astype(np.float32)
        k = np.atleast_1d(np.asarray(args[1] if len(args) > 1 else 1, dtype=np.float32))
        if len(k) == 1:
            k = np.repeat(k, len(idx))
        if len(k) != len(idx):
            raise ValueError("Please provide either a single sharpness or one for each column index.")
        return [dict(kind=N.STAGE_INV_SOFTPLUS, vec_a=np.ascontiguousarray(idx), vec_b=np.ascontiguousarray(k))]
    if name == "ColorTransform":
        ref_idx, mag_idx = args
        mags = np.ascontiguousarray(np.asarray(mag_idx, dtype=np.float32).ravel())
        return [dict(kind=N.STAGE_COLOR_TRANSFORM, shift=int(ref_idx), vec_a=mags, vec_b=mags)]
    if name == "RollingSplineCoupling":
        full = tuple(args) + (None, 1, 16, 5, 2, 128, None, 0, False)[len(args):]
        nlayers, shift, K, B, hl, hd, td, nc, periodic = full
        nsc = ("NeuralSplineCoupling", (K, B, hl, hd, td, nc, periodic))
        return flatten(("Chain", (nsc, ("Roll", (shift,))) * nlayers), params, input_dim)
    if name == "NeuralSplineCoupling":
        full = tuple(args) + (16, 5, 2, 128, None, 0, False)[len(args):]
        K, B, hidden_layers, hidden_dim, transformed_dim, n_conditions, periodic = full
        weights = []
        for layer in params:
            if len(layer) == 0:
                continue
            W, b = layer
            weights += [_f32(W), _f32(b)]
        if len(weights) != 2 * (hidden_layers + 1):
            raise ValueError("NeuralSplineCoupling params do not match hidden_layers")
        lower = input_dim - input_dim // 2 if transformed_dim is None else int(transformed_dim)
        upper = input_dim - lower
        dims = [upper + n_conditions] + [hidden_dim] * hidden_layers + \
               [(3 * K - 1 + int(bool(periodic))) * lower]
        for i in range(hidden_layers + 1):
            if weights[2 * i].shape != (dims[i], dims[i + 1]) or weights[2 * i + 1].shape != (dims[i + 1],):
                raise ValueError(
                    f"NeuralSplineCoupling layer {i}: expected W{(dims[i], dims[i + 1])}, "
                    f"got W{weights[2 * i].shape}, b{weights[2 * i + 1].shape}")
        return [dict(kind=N.STAGE_NSC, K=int(K), B=float(B), hidden_layers=int(hidden_layers),
                     hidden_dim=int(hidden_dim),
                     transformed_dim=-1 if transformed_dim is None else int(transformed_dim),
                     n_conditions=int(n_conditions), periodic=int(bool(periodic)), weights=weights)]
    raise NotImplementedError(
        f"bijector {name!r} is outside the B200 hot path (supported: {', '.join(HOT_PATH_BIJECTORS)})")


def _param_leaves(params: Any, out: list) -> list:
    if isinstance(params, (list, tuple)):
        for p in params:
            _param_leaves(p, out)
    else:
        out.append(params)
    return out


def _leaf_digest(leaf) -> tuple:
    """Content check of one leaf (about 9 GB/s on one host core: 0.13 ms for the 7-D flow's 1.2 MB of weights)."""
    if leaf is None or isinstance(leaf, (bool, int, float, str)):
        return (type(leaf).__name__, leaf)
    if isinstance(leaf, torch.Tensor):
        leaf = leaf.detach().cpu().numpy()
    a = np.ascontiguousarray(leaf)
    raw = a.reshape(-1).view(np.uint8)
    n8 = raw.size // 8
    s = int(raw[: n8 * 8].view(np.uint64).sum(dtype=np.uint64)) if n8 else 0
    t = int(raw[n8 * 8:].astype(np.uint64).sum()) if raw.size > n8 * 8 else 0
    return (a.dtype.str, a.shape, s, t)


def _digest(leaves: Sequence) -> tuple:
    """Content digest of all leaves.  NumPy arrays go to pzb_host_digest in ONE call (the library's host cores; ~15 us for
    the 7-D flow's 42 arrays instead of ~130 us of NumPy reductions per Flow call); anything else, or a missing library,
    takes ``_leaf_digest``.  Same sums either way."""
    idx, ptrs, sizes = [], [], []
    out: list = [None] * len(leaves)
    for i, leaf in enumerate(leaves):
        if isinstance(leaf, np.ndarray) and leaf.flags.c_contiguous and leaf.dtype != object:
            idx.append(i)
            ptrs.append(leaf.__array_interface__["data"][0])
            sizes.append(leaf.nbytes)
        else:
            out[i] = _leaf_digest(leaf)
    if idx:
        try:
            lib = N.lib()
            n = len(idx)
            sums = (C.c_uint64 * (2 * n))()
            N.check(lib.pzb_host_digest((C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*sizes), n, sums))
            for k, i in enumerate(idx):
                a = leaves[i]
                out[i] = (a.dtype.str, a.shape, int(sums[2 * k]), int(sums[2 * k + 1]))
        except (N.NativeError, OSError, AttributeError):
            for i in idx:
                out[i] = _leaf_digest(leaves[i])
    return tuple(out)


class PlanCache:
    """Device plans keyed on a params pytree.  An entry holds a STRONG reference to every leaf it was built from
    and matches a query only if each leaf is the same object (``is``; ids of freed arrays get reused, so ids alone
    prove nothing) on the same device AND the leaves' bytes still sum to what they did at build time (catches
    ``W *= 2`` style in-place edits)."""

    def __init__(self, max_plans: int = 1) -> None:
        self.max_plans = max_plans
        self._entries: list = []  # (device, leaves, digest, plan), most recently used last

    def clear(self) -> None:
        self._entries = []

    def __len__(self) -> int:
        return len(self._entries)

    def get(self, params: Any, device: int, build) -> "Plan":
        leaves = _param_leaves(params, [])
        digest = None
        for i, (dev, held, dg, plan) in enumerate(self._entries):
            if dev != device or len(held) != len(leaves) or not all(a is b for a, b in zip(held, leaves)):
                continue
            if digest is None:
                digest = _digest(leaves)
            if dg == digest:
                self._entries.append(self._entries.pop(i))
                return plan
            del self._entries[i]  # same arrays, edited in place: the old plan is stale
            break
        if digest is None:
            digest = _digest(leaves)
        plan = build()
        self._entries.append((device, leaves, digest, plan))
        while len(self._entries) > self.max_plans:
            self._entries.pop(0)
        return plan


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Plan:
    """One compiled chain + latent on one CUDA device."""

    def __init__(self, bijector_info, bijector_params, input_dim: int,
                 latent_info=("Uniform", (None, 5)), device: Optional[int] = None, latent_params=None, rng_key=None):
        lib = N.lib()
        if lib.pzb_device_count() <= 0 or not torch.cuda.is_available():
            raise N.NativeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        latent_name, latent_args = latent_info
        if latent_name not in LATENTS:
            raise NotImplementedError(
                f"latent {latent_name!r} is outside the B200 hot path (supported: {', '.join(LATENTS)})")
        latent_B = (float(latent_args[1]) if len(latent_args) > 1 and latent_name in ("Uniform", "CentBeta13", "CentBeta")
                    else 5.0)
        self._latent_name = latent_name
        stages = flatten(resolve_shuffles(bijector_info, rng_key, int(input_dim)), bijector_params, input_dim)
        self._keep = []  # host arrays the descriptors point into (until plan_create returns)
        descs = (N.StageDesc * max(1, len(stages)))()
        for d, st in zip(descs, stages):
            d.kind = st["kind"]
            d.shift = st.get("shift", 0)
            d.B = st.get("B", 0.0)
            for f in ("K", "hidden_layers", "hidden_dim", "transformed_dim", "n_conditions", "periodic"):
                setattr(d, f, st.get(f, 0))
            if "vec_a" in st:
                a, b = st["vec_a"], st["vec_b"]
                self._keep += [a, b]
                d.vec_a = a.ctypes.data_as(N.c_float_p)
                d.vec_b = b.ctypes.data_as(N.c_float_p)
                d.n_vec = len(a)
            if "weights" in st:
                arr = (N.c_float_p * len(st["weights"]))()
                for i, w in enumerate(st["weights"]):
                    arr[i] = w.ctypes.data_as(N.c_float_p)
                self._keep += st["weights"] + [arr]
                d.weights = C.cast(arr, C.POINTER(N.c_float_p))
        handle = C.c_void_p()
        N.check(lib.pzb_plan_create(descs, len(stages), int(input_dim), LATENTS[latent_name],
                                    latent_B, self.device_index, C.byref(handle)))
        self._keep = []
        self._h = handle
        if latent_name in ("CentBeta", "Tdist", "Joint"):
            segs = latent_segments(latent_info, latent_params)
            arr = (N.LatentDesc * len(segs))()
            for d, (kind, dim, B, flat) in zip(arr, segs):
                d.kind, d.dim, d.B = kind, dim, B
                d.params = flat.ctypes.data_as(N.c_float_p) if flat is not None else None
                d.n_params = 0 if flat is None else flat.size
            N.check(lib.pzb_plan_set_latent(handle, arr, len(segs)))
        self.input_dim = int(input_dim)
        self.n_conditions = lib.pzb_plan_n_conditions(handle)
        self.n_spline_dims = lib.pzb_plan_n_spline_dims(handle)
        self.flops_per_row = lib.pzb_plan_flops_per_row(handle)
        self.tc_supported = bool(lib.pzb_plan_tc_supported(handle))
        self.wide_supported = bool(lib.pzb_plan_wide_supported(handle))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                N.lib().pzb_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- engine selection -------------------------------------------------
    def set_engine(self, engine: str) -> None:
        code = {"auto": N.ENGINE_AUTO, "simt": N.ENGINE_SIMT, "tc": N.ENGINE_TC, "wide": N.ENGINE_WIDE}[engine]
        N.check(N.lib().pzb_plan_set_engine(self._h, code))

    @property
    def engine(self) -> str:
        return {N.ENGINE_SIMT: "simt", N.ENGINE_TC: "tc", N.ENGINE_WIDE: "wide"}[N.lib().pzb_plan_engine(self._h)]

    def overflow_rows(self, reset: bool = False) -> int:
        """Rows (cumulative) whose spline logits came out non-finite on a tensor-core engine: an fp16 overflow of a
        conditioner pre-activation beyond +-65504, where the reference's fp32 stays finite (pzb_plan_overflow_rows)."""
        n = C.c_int64(0)
        N.check(N.lib().pzb_plan_overflow_rows(self._h, int(bool(reset)), C.byref(n)))
        return int(n.value)

    # ---- tensor plumbing --------------------------------------------------
    def _rows(self, X, width: int, what: str) -> torch.Tensor:
        t = torch.as_tensor(X) if not isinstance(X, torch.Tensor) else X
        if t.dim() != 2 or t.shape[1] != width:
            raise ValueError(f"{what} must have shape [N, {width}], got {tuple(t.shape)}")
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def _cond(self, cond, n_rows: int) -> Optional[torch.Tensor]:
        if self.n_conditions == 0:
            return None  # the reference's zeros[N, 1] dummy is sliced away (bijectors.py:481-483)
        if cond is None:
            raise ValueError(f"this bijector is conditional: conditions [N, {self.n_conditions}] are required")
        t = torch.as_tensor(cond) if not isinstance(cond, torch.Tensor) else cond
        if t.dim() != 2 or t.shape[0] != n_rows or t.shape[1] < self.n_conditions:
            raise ValueError(f"conditions must have shape [{n_rows}, {self.n_conditions}], got {tuple(t.shape)}")
        t = t[:, : self.n_conditions]
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    # ---- host-buffer entry points (blocking; chunks pipelined H2D | kernel | D2H) -------
    def _host_rows(self, X, width: int, what: str) -> torch.Tensor:
        """A C-contiguous float32 CPU tensor view of X (no copy when X already is one)."""
        t = torch.as_tensor(X) if not isinstance(X, torch.Tensor) else X
        if t.dim() != 2 or t.shape[1] != width:
            raise ValueError(f"{what} must have shape [N, {width}], got {tuple(t.shape)}")
        return t.to(dtype=torch.float32).contiguous()

    def _host_cond(self, cond, n_rows: int) -> Optional[torch.Tensor]:
        if self.n_conditions == 0:
            return None
        if cond is None:
            raise ValueError(f"this bijector is conditional: conditions [N, {self.n_conditions}] are required")
        t = torch.as_tensor(cond) if not isinstance(cond, torch.Tensor) else cond
        if t.dim() != 2 or t.shape[0] != n_rows or t.shape[1] < self.n_conditions:
            raise ValueError(f"conditions must have shape [{n_rows}, {self.n_conditions}], got {tuple(t.shape)}")
        return t[:, : self.n_conditions].to(dtype=torch.float32).contiguous()

    def log_prob_host(self, X, cond=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Flow._log_prob on HOST arrays -> CPU tensor [N] (pzb_log_prob_host).  Fresh results come from ``host_result``
        (page-locked up to 256 MB, pageable beyond); pageable inputs and results are staged through the library's own
        page-locked ring."""
        X = self._host_rows(X, self.input_dim, "inputs")
        cond = self._host_cond(cond, X.shape[0])
        if out is None:
            out = host_result((X.shape[0],))
        N.check(N.lib().pzb_log_prob_host(self._h, _ptr(X), _ptr(cond), X.shape[0], _ptr(out)))
        return out

    def log_prob_host_columns(self, data_cols, cond_cols=None, cond_means=None, cond_stds=None,
                              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Flow.log_prob from float64 (or float32) DataFrame COLUMNS (one contiguous 1-D array per column): the
        down-cast to float32, the row gather and the condition scaling run inside the pipelined C call
        (pzb_log_prob_host_columns / _f32).  Returns a CPU tensor [N]."""
        # float32 columns stay float32 (pzb_log_prob_host_columns_f32); anything else is taken as float64
        f32 = all(getattr(c, "dtype", None) == np.float32 for c in data_cols)
        ctype = np.float32 if f32 else np.float64
        cols = [np.ascontiguousarray(c, dtype=ctype) for c in data_cols]
        if len(cols) != self.input_dim:
            raise ValueError(f"expected {self.input_dim} data columns, got {len(cols)}")
        n = len(cols[0])
        if any(c.ndim != 1 or len(c) != n for c in cols):
            raise ValueError("data columns must be 1-D arrays of equal length")
        ccols, means, stds = [], None, None
        if self.n_conditions > 0:
            if cond_cols is None:
                raise ValueError(f"this bijector is conditional: conditions [N, {self.n_conditions}] are required")
            ccols = [np.ascontiguousarray(c, dtype=ctype) for c in list(cond_cols)[: self.n_conditions]]
            if len(ccols) < self.n_conditions or any(c.ndim != 1 or len(c) != n for c in ccols):
                raise ValueError(f"conditions must be {self.n_conditions} columns of {n} values")
            means = np.ascontiguousarray(np.zeros(self.n_conditions) if cond_means is None else cond_means,
                                         dtype=np.float32)[: self.n_conditions]
            stds = np.ascontiguousarray(np.ones(self.n_conditions) if cond_stds is None else cond_stds,
                                        dtype=np.float32)[: self.n_conditions]
        dptr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        cptr = (C.c_void_p * max(1, len(ccols)))(*[c.ctypes.data for c in ccols]) if ccols else None
        if out is None:
            out = host_result((n,))
        call = N.lib().pzb_log_prob_host_columns_f32 if f32 else N.lib().pzb_log_prob_host_columns
        N.check(call(
            self._h, dptr, cptr, means.ctypes.data if means is not None else None,
            stds.ctypes.data if stds is not None else None, n, _ptr(out)))
        return out

    def inverse_host(self, U, cond=None, want_log_det: bool = False):
        """Chain InverseFunction on HOST arrays -> (x [N, D], log_det [N] or None), CPU tensors."""
        U = self._host_rows(U, self.input_dim, "inputs")
        cond = self._host_cond(cond, U.shape[0])
        X = host_result(tuple(U.shape))
        ld = host_result((U.shape[0],)) if want_log_det else None
        N.check(N.lib().pzb_inverse_host(self._h, _ptr(U), _ptr(cond), U.shape[0], _ptr(X), _ptr(ld)))
        return X, ld

    def inverse_to_host(self, U, cond=None, want_log_det: bool = False):
        """Chain InverseFunction for DEVICE-resident latent draws U [N, D] -> (x [N, D], log_det [N] or None) as CPU
        tensors, through the pipelined C call (pzb_inverse_to_host)."""
        U = self._rows(U, self.input_dim, "inputs")
        cond = self._cond(cond, U.shape[0])
        X = host_result(tuple(U.shape))
        ld = host_result((U.shape[0],)) if want_log_det else None
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_inverse_to_host(self._h, _ptr(U), _ptr(cond), U.shape[0], _ptr(X), _ptr(ld),
                                                torch.cuda.current_stream(self.device).cuda_stream))
        return X, ld

    def inverse_to_host_columns(self, U, cond=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Like inverse_to_host, but returns the samples as a NumPy array of shape [N, D] whose memory is
        COLUMN-major (the transpose of a C-contiguous [D, N] buffer): `pd.DataFrame(x, columns=...)` wraps that
        without copying (pzb_inverse_to_host_columns).  ``out``: a float32 [D, N] array whose rows are contiguous
        (e.g. a column slice of a bigger [D, N_total] buffer) to write into."""
        U = self._rows(U, self.input_dim, "inputs")
        cond = self._cond(cond, U.shape[0])
        n = U.shape[0]
        buf = np.empty((self.input_dim, n), dtype=np.float32) if out is None else out
        if buf.dtype != np.float32 or buf.shape != (self.input_dim, n) or (n > 1 and buf.strides[1] != 4):
            raise ValueError(f"out must be a float32 array of shape ({self.input_dim}, {n}) with contiguous rows")
        cols = (C.c_void_p * self.input_dim)(*[buf[j].ctypes.data for j in range(self.input_dim)])
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_inverse_to_host_columns(self._h, _ptr(U), _ptr(cond), n, cols,
                                                        torch.cuda.current_stream(self.device).cuda_stream))
        return buf.T

    def forward_host(self, X, cond=None):
        """Chain ForwardFunction on HOST arrays -> (u [N, D], log_det [N]), CPU tensors."""
        X = self._host_rows(X, self.input_dim, "inputs")
        cond = self._host_cond(cond, X.shape[0])
        U = host_result(tuple(X.shape))
        ld = host_result((X.shape[0],))
        N.check(N.lib().pzb_forward_host(self._h, _ptr(X), _ptr(cond), X.shape[0], _ptr(U), _ptr(ld)))
        return U, ld

    def posterior_host(self, rest, col_idx: int, grid, cond=None, normalize: bool = True,
                       nan_to_zero: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Flow.posterior main loop on HOST arrays: rest [Ng, D-1], grid [G] -> pdfs [Ng, G] in host memory;
        neither the [Ng*G, D] expansion nor the whole pdfs array ever sits in HBM.  ``out`` may be page-locked
        (copied into directly) or ordinary pageable memory (the library stages chunks through its own page-locked
        ring and moves them on with all host cores while the GPU computes the next ones); the default is
        ``host_result``: page-locked up to 256 MB, pageable beyond -- page-locking gigabytes for one call costs more
        than the evaluation itself."""
        rest = self._host_rows(rest, self.input_dim - 1, "inputs")
        cond = self._host_cond(cond, rest.shape[0])
        grid = torch.as_tensor(grid).to(dtype=torch.float32).contiguous().reshape(-1).cpu()
        Ng, G = rest.shape[0], grid.numel()
        if out is None:
            out = host_result((Ng, G))
        N.check(N.lib().pzb_posterior_host(self._h, _ptr(rest) if rest.numel() else None, _ptr(cond), Ng,
                                           int(col_idx), _ptr(grid), G, int(bool(normalize)),
                                           int(bool(nan_to_zero)), _ptr(out)))
        return out

    # ---- Gaussian error convolution (err_samples), device tensors -------
    ERR_SCRATCH_FLOATS = 1 << 26  # expanded rows evaluated per launch (256 MB of scratch at most)

    def _dev_opt(self, a, shape, what):
        if a is None:
            return None
        t = torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{what} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def log_prob_err(self, X, Xerr, cond=None, cond_err=None, err_samples: int = 1, seed: int = 0) -> torch.Tensor:
        """Flow.log_prob(err_samples=S) for arrays (pzflow/flow.py:448-464): log(mean_s exp(lp(x + eps x_err)))
        with the Gaussian expansion generated inside the kernel.  cond / cond_err are standard-scaled."""
        X = self._rows(X, self.input_dim, "inputs")
        n = X.shape[0]
        Xerr = self._dev_opt(Xerr, X.shape, "input errors")
        cond = self._cond(cond, n)
        cond_err = self._dev_opt(cond_err, (n, self.n_conditions), "condition errors") if cond is not None else None
        S = int(err_samples)
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        step = max(1, self.ERR_SCRATCH_FLOATS // S)
        scratch = torch.empty(min(n, step) * S, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device), _ErrorStream(n, S, max(self.input_dim, self.n_conditions)):
            for r0 in range(0, n, step):
                m = min(step, n - r0)
                N.check(N.lib().pzb_log_prob_err(
                    self._h, _ptr(X[r0:r0 + m]), _ptr(Xerr[r0:r0 + m]) if Xerr is not None else None,
                    _ptr(cond[r0:r0 + m]) if cond is not None else None,
                    _ptr(cond_err[r0:r0 + m]) if cond_err is not None else None, m, S,
                    int(seed) & (2 ** 64 - 1), r0, _ptr(out[r0:r0 + m]), _ptr(scratch), _stream(self.device)))
        return out

    def posterior_err(self, rest, rest_err, col_idx: int, grid, cond=None, cond_err=None, err_samples: int = 1,
                      seed: int = 0, normalize: bool = True, nan_to_zero: bool = True) -> torch.Tensor:
        """Flow.posterior(err_samples=S) main loop (pzflow/flow.py:675-737) -> pdfs [Ng, G] on the device."""
        rest = self._rows(rest, self.input_dim - 1, "inputs")
        Ng = rest.shape[0]
        rest_err = self._dev_opt(rest_err, rest.shape, "input errors")
        cond = self._cond(cond, Ng)
        cond_err = self._dev_opt(cond_err, (Ng, self.n_conditions), "condition errors") if cond is not None else None
        grid = torch.as_tensor(grid).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        G, S = grid.numel(), int(err_samples)
        out = torch.empty((Ng, G), dtype=torch.float32, device=self.device)
        step = max(1, self.ERR_SCRATCH_FLOATS // (S * G))
        scratch = torch.empty(min(Ng, step) * S * G, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device), _ErrorStream(Ng, S, max(self.input_dim, self.n_conditions)):
            for r0 in range(0, Ng, step):
                m = min(step, Ng - r0)
                N.check(N.lib().pzb_posterior_err(
                    self._h, _ptr(rest[r0:r0 + m]) if rest.numel() else None,
                    _ptr(rest_err[r0:r0 + m]) if rest_err is not None and rest_err.numel() else None,
                    _ptr(cond[r0:r0 + m]) if cond is not None else None,
                    _ptr(cond_err[r0:r0 + m]) if cond_err is not None else None, m, int(col_idx), _ptr(grid), G, S,
                    int(seed) & (2 ** 64 - 1), r0, int(bool(normalize)), int(bool(nan_to_zero)),
                    _ptr(out[r0:r0 + m]), _ptr(scratch), _stream(self.device)))
        return out

    def posterior_marg(self, rest, col_idx: int, grid, mcols, mvals, rest_err=None, cond=None, cond_err=None,
                       err_samples: int = 0, seed: int = 0, nan_to_zero: bool = True) -> torch.Tensor:
        """Un-normalised posterior sums of galaxies that miss the same data columns (pzflow/flow.py:560-664):
        rest [Ng, D-1] (any value in the missing columns), mcols = indices of the missing columns among the data
        columns, mvals [Ng, V, len(mcols)] the values every variant puts there -> [Ng, G] on the device,
        sum over the variants of nan_to_num(mean over error samples of exp(log_prob)) (pzb_posterior_marg)."""
        rest = self._rows(rest, self.input_dim - 1, "inputs")
        Ng = rest.shape[0]
        mcols = [int(c) for c in mcols]
        mv = torch.as_tensor(mvals).to(device=self.device, dtype=torch.float32).contiguous()
        if mv.dim() != 3 or mv.shape[0] != Ng or mv.shape[2] != len(mcols):
            raise ValueError(f"mvals must have shape [{Ng}, V, {len(mcols)}], got {tuple(mv.shape)}")
        V, S = mv.shape[1], int(err_samples or 0)
        rest_err = self._dev_opt(rest_err, rest.shape, "input errors") if S else None
        cond = self._cond(cond, Ng)
        cond_err = self._dev_opt(cond_err, (Ng, self.n_conditions), "condition errors") if (S and cond is not None) else None
        grid = torch.as_tensor(grid).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        G = grid.numel()
        out = torch.empty((Ng, G), dtype=torch.float32, device=self.device)
        per = V * max(S, 1) * G
        step = max(1, self.ERR_SCRATCH_FLOATS // per)
        scratch = torch.empty(min(Ng, step) * per, dtype=torch.float32, device=self.device)
        cols = (C.c_int32 * len(mcols))(*mcols)
        with torch.cuda.device(self.device):
            for r0 in range(0, Ng, step):
                m = min(step, Ng - r0)
                N.check(N.lib().pzb_posterior_marg(
                    self._h, _ptr(rest[r0:r0 + m]) if rest.numel() else None,
                    _ptr(rest_err[r0:r0 + m]) if rest_err is not None and rest_err.numel() else None,
                    _ptr(cond[r0:r0 + m]) if cond is not None else None,
                    _ptr(cond_err[r0:r0 + m]) if cond_err is not None else None, m, int(col_idx), _ptr(grid), G,
                    len(mcols), cols, _ptr(mv[r0:r0 + m]), V, S, int(seed) & (2 ** 64 - 1), r0, int(bool(nan_to_zero)),
                    _ptr(out[r0:r0 + m]), _ptr(scratch), _stream(self.device)))
        return out

    # ---- entry points (all asynchronous on torch's current stream) -------
    def log_prob(self, X, cond=None, return_bins: bool = False, out: Optional[torch.Tensor] = None):
        """Flow._log_prob (pzflow/flow.py:397-407).  Device tensors are evaluated in place on torch's
        current stream; HOST arrays (NumPy / CPU tensors) take the pipelined host-buffer path and
        come back as a CPU tensor.  ``out``: a contiguous float32 device tensor [N] to write into."""
        if not return_bins and not (isinstance(X, torch.Tensor) and X.is_cuda):
            return self.log_prob_host(X, cond, out=out)
        X = self._rows(X, self.input_dim, "inputs")
        cond = self._cond(cond, X.shape[0])
        if out is None:
            out = torch.empty(X.shape[0], dtype=torch.float32, device=self.device)
        elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == X.shape[0]):
            raise ValueError("out must be a contiguous float32 device tensor with one element per row")
        bins = (torch.empty((X.shape[0], self.n_spline_dims), dtype=torch.int32, device=self.device)
                if return_bins else None)
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_log_prob(self._h, _ptr(X), _ptr(cond), X.shape[0], _ptr(out), _ptr(bins),
                                         _stream(self.device)))
        return (out, bins) if return_bins else out

    def forward(self, X, cond=None, return_bins: bool = False):
        """Chain ForwardFunction (pzflow/bijectors.py:162-164): (u, log_det)."""
        X = self._rows(X, self.input_dim, "inputs")
        cond = self._cond(cond, X.shape[0])
        U = torch.empty_like(X)
        ld = torch.empty(X.shape[0], dtype=torch.float32, device=self.device)
        bins = (torch.empty((X.shape[0], self.n_spline_dims), dtype=torch.int32, device=self.device)
                if return_bins else None)
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_forward(self._h, _ptr(X), _ptr(cond), X.shape[0], _ptr(U), _ptr(ld),
                                        _ptr(bins), _stream(self.device)))
        return (U, ld, bins) if return_bins else (U, ld)

    def inverse(self, U, cond=None, return_bins: bool = False, want_log_det: bool = True):
        """Chain InverseFunction (pzflow/bijectors.py:166-170): (x, log_det)."""
        U = self._rows(U, self.input_dim, "inputs")
        cond = self._cond(cond, U.shape[0])
        X = torch.empty_like(U)
        ld = torch.empty(U.shape[0], dtype=torch.float32, device=self.device) if want_log_det else None
        bins = (torch.empty((U.shape[0], self.n_spline_dims), dtype=torch.int32, device=self.device)
                if return_bins else None)
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_inverse(self._h, _ptr(U), _ptr(cond), U.shape[0], _ptr(X), _ptr(ld),
                                        _ptr(bins), _stream(self.device)))
        return (X, ld, bins) if return_bins else (X, ld)

    def posterior(self, rest, col_idx: int, grid, cond=None, normalize: bool = True,
                  nan_to_zero: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Flow.posterior main loop (pzflow/flow.py:666-738) without materialising the
        [Ng*G, D] expansion: rest [Ng, D-1], grid [G] -> pdfs [Ng, G]."""
        rest = self._rows(rest, self.input_dim - 1, "inputs")
        cond = self._cond(cond, rest.shape[0])
        grid = torch.as_tensor(grid).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        Ng, G = rest.shape[0], grid.numel()
        if out is None:
            out = torch.empty((Ng, G), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().pzb_posterior(self._h, _ptr(rest), _ptr(cond), Ng, int(col_idx), _ptr(grid), G,
                                          int(bool(normalize)), int(bool(nan_to_zero)), _ptr(out),
                                          _stream(self.device)))
        return out


# ---- reductions that are not tied to one plan ---------------------------------
def ensemble_log_prob_host_columns(plans, data_cols, cond_cols=None, cond_means=None, cond_stds=None,
                                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FlowEnsemble.log_prob from DataFrame COLUMNS (float64, or all float32): each chunk of rows is staged once,
    evaluated by every plan and reduced with log(mean(exp)) on the device (pzb_ensemble_log_prob_host_columns).
    Returns a CPU tensor [N]."""
    p0 = plans[0]
    f32 = all(getattr(c, "dtype", None) == np.float32 for c in data_cols)
    ctype = np.float32 if f32 else np.float64
    cols = [np.ascontiguousarray(c, dtype=ctype) for c in data_cols]
    if len(cols) != p0.input_dim:
        raise ValueError(f"expected {p0.input_dim} data columns, got {len(cols)}")
    n = len(cols[0])
    if any(c.ndim != 1 or len(c) != n for c in cols):
        raise ValueError("data columns must be 1-D arrays of equal length")
    ccols, means, stds = [], None, None
    if p0.n_conditions > 0:
        if cond_cols is None:
            raise ValueError(f"this bijector is conditional: conditions [N, {p0.n_conditions}] are required")
        ccols = [np.ascontiguousarray(c, dtype=ctype) for c in list(cond_cols)[: p0.n_conditions]]
        if len(ccols) < p0.n_conditions or any(c.ndim != 1 or len(c) != n for c in ccols):
            raise ValueError(f"conditions must be {p0.n_conditions} columns of {n} values")
        means = np.ascontiguousarray(np.zeros(p0.n_conditions) if cond_means is None else cond_means,
                                     dtype=np.float32)[: p0.n_conditions]
        stds = np.ascontiguousarray(np.ones(p0.n_conditions) if cond_stds is None else cond_stds,
                                    dtype=np.float32)[: p0.n_conditions]
    handles = (C.c_void_p * len(plans))(*[p._h for p in plans])
    dptr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    cptr = (C.c_void_p * max(1, len(ccols)))(*[c.ctypes.data for c in ccols]) if ccols else None
    if out is None:
        out = host_result((n,))
    N.check(N.lib().pzb_ensemble_log_prob_host_columns(
        handles, len(plans), dptr, int(f32), cptr, means.ctypes.data if means is not None else None,
        stds.ctypes.data if stds is not None else None, n, _ptr(out)))
    return out


def ensemble_posterior_host(plans, rest, col_idx: int, grid, cond=None, normalize: bool = True,
                            nan_to_zero: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FlowEnsemble.posterior on HOST arrays (pzb_ensemble_posterior_host): rest [Ng, D-1], grid [G] -> pdfs [Ng, G],
    the mean over the plans' un-normalised pdfs, normalised on the device chunk by chunk."""
    p0 = plans[0]
    rest = p0._host_rows(rest, p0.input_dim - 1, "inputs")
    cond = p0._host_cond(cond, rest.shape[0])
    grid = torch.as_tensor(grid).to(dtype=torch.float32).contiguous().reshape(-1).cpu()
    Ng, G = rest.shape[0], grid.numel()
    if out is None:
        out = host_result((Ng, G))
    handles = (C.c_void_p * len(plans))(*[p._h for p in plans])
    N.check(N.lib().pzb_ensemble_posterior_host(handles, len(plans), _ptr(rest) if rest.numel() else None, _ptr(cond),
                                                Ng, int(col_idx), _ptr(grid), G, int(bool(normalize)),
                                                int(bool(nan_to_zero)), _ptr(out)))
    return out


def ensemble_logmeanexp(lp: torch.Tensor) -> torch.Tensor:
    """log(mean(exp(lp), axis=flows)) for lp [n_flows, N] (pzflow/flowEnsemble.py:243)."""
    lp = lp.contiguous()
    out = torch.empty(lp.shape[1], dtype=torch.float32, device=lp.device)
    with torch.cuda.device(lp.device):
        N.check(N.lib().pzb_ensemble_logmeanexp(_ptr(lp), lp.shape[0], lp.shape[1], _ptr(out),
                                                _stream(lp.device)))
    return out


def ensemble_posterior_mean(pdfs: torch.Tensor, grid: torch.Tensor, normalize: bool) -> torch.Tensor:
    """mean over flows of pdfs [n_flows, Ng, G], then trapezoid-normalise
    (pzflow/flowEnsemble.py:346-351)."""
    pdfs = pdfs.contiguous()
    grid = grid.to(device=pdfs.device, dtype=torch.float32).contiguous()
    out = torch.empty(pdfs.shape[1:], dtype=torch.float32, device=pdfs.device)
    with torch.cuda.device(pdfs.device):
        N.check(N.lib().pzb_ensemble_posterior_mean(_ptr(pdfs), pdfs.shape[0], pdfs.shape[1], pdfs.shape[2],
                                                    _ptr(grid), int(bool(normalize)), _ptr(out),
                                                    _stream(pdfs.device)))
    return out


def normalize_pdfs_(pdfs: torch.Tensor, grid: torch.Tensor, normalize: bool, nan_to_zero: bool) -> torch.Tensor:
    """In-place per-row trapezoid normalisation + NaN->0 of pdfs [Ng, G] (pzflow/flow.py:732-737)."""
    assert pdfs.is_contiguous()
    grid = grid.to(device=pdfs.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(pdfs.device):
        N.check(N.lib().pzb_normalize_pdfs(_ptr(pdfs), pdfs.shape[0], pdfs.shape[1], _ptr(grid),
                                           int(bool(normalize)), int(bool(nan_to_zero)), _stream(pdfs.device)))
    return pdfs


def _error_stream(total_units: int, S: int, width: int):
    """Which stream the Gaussian error samples of a call over ``total_units`` rows / galaxies come from: jax's own
    (``random.normal(PRNGKey(seed), (units, S, W))``, pzflow/utils.py:59) in the layout ``prng.layout()`` names, or --
    for ``"philox"`` and for calls beyond one threefry call of jax's original layout -- the library's Philox streams.
    Returns the layout code of pzb_set_error_stream."""
    from . import prng
    lay = prng.layout()
    if lay == "philox" or (lay == "threefry" and total_units * S * max(1, width) > prng.MAX_THREEFRY_WORDS):
        return 0
    return 1 if lay == "threefry" else 2


class _ErrorStream:
    """``with _ErrorStream(units, S, W):`` -- the pzb_*_err calls of this host thread inside the block draw their
    Gaussian error samples from the stream ``_error_stream`` picks; the library's own streams are restored on exit."""

    def __init__(self, total_units: int, S: int, width: int):
        self.layout, self.units = _error_stream(total_units, S, width), int(total_units)

    def __enter__(self):
        N.check(N.lib().pzb_set_error_stream(self.layout, self.units))
        return self

    def __exit__(self, *exc):
        N.lib().pzb_set_error_stream(0, 0)
        return False


def sample_uniform(n: int, dim: int, B: float, seed: int, device=None, row_offset: int = 0) -> torch.Tensor:
    """U(-B, B)^dim latent draws from the library's Philox stream (not jax's threefry).  ``row_offset``: index of the
    first row in the caller's whole array (a sharded call draws what one big call would)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_uniform(_ptr(out), n, dim, float(B), int(seed) & (2 ** 64 - 1), int(row_offset),
                                           _stream(device)))
    return out


def sample_uniform_jax(n: int, dim: int, B: float, seed, device=None, row_offset: int = 0, total_rows: int = None,
                       layout: str = None) -> torch.Tensor:
    """Rows [row_offset, row_offset + n) of ``jax.random.uniform(PRNGKey(seed), (total_rows, dim), -B, B)``, drawn on the
    device in jax's own threefry stream (pzb_sample_uniform_threefry; pzflow/distributions.py:464-469)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    total_rows = n + row_offset if total_rows is None else int(total_rows)
    key = prng.key_from_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_uniform_threefry(_ptr(out), n, dim, -float(B), float(B), int(key[0]), int(key[1]),
                                                    total_rows, int(row_offset),
                                                    int((layout or prng.layout()) == "threefry_partitionable"),
                                                    _stream(device)))
    return out


def sample_centbeta(n: int, a, b, B: float, seed: int, device=None, row_offset: int = 0) -> torch.Tensor:
    """2B (Beta(a_j, b_j) - 0.5) latent draws [n, len(a)] from the library's Philox streams (pzb_sample_centbeta)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    out = torch.empty((n, len(a)), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_centbeta(_ptr(out), n, len(a), a.ctypes.data, b.ctypes.data, float(B),
                                            int(seed) & (2 ** 64 - 1), int(row_offset), _stream(device)))
    return out


def sample_normal(n: int, dim: int, seed: int, device=None, row_offset: int = 0) -> torch.Tensor:
    """Standard normal latent draws [n, dim] (pzb_sample_normal)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_normal(_ptr(out), n, dim, int(seed) & (2 ** 64 - 1), int(row_offset), _stream(device)))
    return out


def sample_normal_jax(n: int, dim: int, seed, device=None, row_offset: int = 0, total_rows: int = None,
                      layout: str = None) -> torch.Tensor:
    """Rows [row_offset, row_offset + n) of ``jax.random.normal(PRNGKey(seed), (total_rows, dim))``, generated on the
    device in jax's own threefry stream (pzb_sample_normal_threefry): the Normal latent's draws
    (pzflow/distributions.py:297-303) and the eps of the default error model (pzflow/utils.py:59)."""
    from . import prng
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    total_rows = row_offset + n if total_rows is None else int(total_rows)
    key = prng.key_from_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_normal_threefry(_ptr(out), n, dim, int(key[0]), int(key[1]), total_rows, int(row_offset),
                                                   int((layout or prng.layout()) == "threefry_partitionable"),
                                                   _stream(device)))
    return out


def sample_tdist(n: int, dim: int, nu: float, seed: int, device=None, row_offset: int = 0) -> torch.Tensor:
    """Multivariate-t latent draws [n, dim], unit scale matrix, nu degrees of freedom (pzb_sample_tdist)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_sample_tdist(_ptr(out), n, dim, float(nu), int(seed) & (2 ** 64 - 1), int(row_offset),
                                         _stream(device)))
    return out


PINNED_RESULT_LIMIT = 256 << 20


def host_result(shape) -> torch.Tensor:
    """A fresh float32 CPU tensor for the results of a host-buffer call.  Up to 256 MB it is page-locked: torch's
    caching host allocator hands the same block back on the next call of the same size, and the D2H copies land in
    it directly.  Larger results are ordinary pageable memory -- page-locking gigabytes for one call costs more than
    the evaluation, and the library stages pageable destinations through its own page-locked ring at the same speed."""
    n = 1
    for d in shape:
        n *= int(d)
    if 4 * n <= PINNED_RESULT_LIMIT and torch.cuda.is_available():
        return torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True)
    return torch.empty(tuple(shape), dtype=torch.float32)


def pinned_empty(shape, dtype=torch.float32) -> torch.Tensor:
    """Page-locked CPU tensor (torch's caching host allocator): D2H/H2D copies at PCIe speed."""
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def error_eps(n: int, width: int, stream_id: int, seed: int, device=None) -> torch.Tensor:
    """The standard normals pzb_log_prob_err / pzb_posterior_err draw: eps [n, width] (stream 0 data, 1 conditions)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n, width), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().pzb_error_eps(_ptr(out), n, width, stream_id, int(seed) & (2 ** 64 - 1), _stream(device)))
    return out


def launch_count() -> int:
    return int(N.lib().pzb_launch_count())

"""``Flow.train`` (pzflow/flow.py:868-1159): the optional data-parallel training step.

What runs where
---------------
* The inference hot path (log_prob / posterior / sample) is the hand-written CUDA of ``csrc/``.  Training is
  NOT on that path.  The gradient of ``-mean(w * log_prob)`` (flow.py:961-965) is assembled by torch autograd
  from two kinds of node: the coupling stage's spline (softmax / softplus knots, bin search, rational-quadratic
  spline, log-determinant) is ONE hand-written forward kernel and ONE hand-derived backward kernel
  (``csrc/nsf_train.cu``, ``_SplineStage``); the conditioner's dense layers are plain batch x 128 x 128 GEMMs
  and stay with cuBLAS (``torch.matmul``), as do the affine bijectors' few elementwise ops.  The update is
  Adam(1e-3) like ``optax.adam`` (flow.py:968-981), one fused kernel over a flat parameter buffer
  (``pzb_adam_step``), and the whole step is captured once as a CUDA graph and replayed per batch
  (``GraphedStep``; ``PZB_TRAIN_GRAPH=0`` runs it eagerly).  ``TorchChain(native_spline=False)`` is the all-torch
  restatement the gradient tests compare the kernels with.  The fused inference kernels evaluate every reported
  loss (initial, end of epoch, validation: flow.py:1033-1050, 1088-1102), so a trained flow is checked by the
  product path as it trains.
* Data parallel: one process per GPU; each rank takes its contiguous share of every batch (batch 1024 over 8
  ranks = 128 rows each), the flat fp32 gradient buffer (every leaf's .grad is a view into it) is summed with
  ONE ``all_reduce`` (NCCL over NVLink on GPUs, inside the captured graph) per step, and every rank applies the
  same Adam update, so parameters stay bit-identical across ranks.

Not reproducible from the reference without jax: its batch permutations and error samples come from
``jax.random`` (flow.py:1063-1066); here they come from NumPy / torch generators seeded from the same ``seed``.
"""
from __future__ import annotations

import math
import os
from typing import Any, List, Optional, Tuple

import numpy as np
import torch

LEAKY_SLOPE = 0.01


# --------------------------------------------------------------------------- #
# differentiable restatement of the chain (reference line numbers as in oracle/) #
# --------------------------------------------------------------------------- #
def _rqs_forward(x, W, H, D, B, periodic):
    """RationalQuadraticSpline forward (pzflow/utils.py:64-227), batched over [..., L] with K bins."""
    K = W.shape[-1]
    pad = lambda t: torch.nn.functional.pad(t, (1, 0), value=0.0)  # noqa: E731
    xk = pad(torch.cumsum(W, dim=-1)) - B          # utils.py:113-118
    yk = pad(torch.cumsum(H, dim=-1)) - B          # utils.py:120-125
    if periodic:
        dk = torch.cat([D[..., -1:], D], dim=-1)   # utils.py:127-128 (wrap)
    else:
        one = torch.ones_like(D[..., :1])
        dk = torch.cat([one, D, one], dim=-1)      # utils.py:130-135
    sk = H / W                                     # utils.py:137
    oob = (x <= -B) | (x >= B)                     # utils.py:145
    mx = torch.where(oob, x.abs() - B, x)          # utils.py:146
    if not periodic:
        # out-of-bounds elements pass through unchanged (208-209, 222-224): evaluate their (discarded) spline at a
        # harmless point so that no inf/NaN reaches the backward pass of the weights they share with other rows
        mx = torch.where(oob, torch.zeros_like(mx), mx)
    idx = (xk <= mx[..., None]).sum(dim=-1) - 1    # utils.py:150-152
    idx = idx.clamp(0, K - 1)                      # gathers outside [0, K) are NaN in the reference; masked below
    g = lambda t, i: torch.gather(t, -1, i[..., None])[..., 0]  # noqa: E731
    xki, yki = g(xk, idx), g(yk, idx)
    wk, hk = g(W, idx), g(H, idx)
    dki, dkp1, ski = g(dk, idx), g(dk, idx + 1), g(sk, idx)
    relx = (mx - xki) / wk                         # utils.py:201
    num = hk * (ski * relx ** 2 + dki * relx * (1 - relx))
    den = ski + (dkp1 + dki - 2 * ski) * relx * (1 - relx)
    out = yki + num / den                          # utils.py:202-206
    dnum = dkp1 * relx ** 2 + 2 * ski * relx * (1 - relx) + dki * (1 - relx) ** 2
    dden = ski + (dkp1 + dki - 2 * ski) * relx * (1 - relx)
    ld = 2 * torch.log(ski) + torch.log(dnum) - 2 * torch.log(dden)   # utils.py:213-220
    if not periodic:
        out = torch.where(oob, x, out)             # utils.py:208-209
        ld = torch.where(oob, torch.zeros_like(ld), ld)               # utils.py:222-224
    return out, ld.sum(dim=-1)


class _SplineStage(torch.autograd.Function):
    """The coupling stage's spline on the hand-written kernels of csrc/nsf_train.cu: forward
    ``pzb_spline_train_forward``, backward ``pzb_spline_train_backward`` (analytic gradient of the softmax /
    softplus knots, bin search and rational-quadratic spline + log-determinant w.r.t. the conditioner's logits and
    the transformed columns), periodic or not."""

    @staticmethod
    def forward(ctx, logits, x, K, B, periodic=False):
        from ._native import check, lib
        logits, x = logits.contiguous().float(), x.contiguous().float()
        N, L = x.shape
        y, ld = torch.empty_like(x), torch.empty_like(x)
        s = torch.cuda.current_stream(x.device).cuda_stream
        check(lib().pzb_spline_train_forward(logits.data_ptr(), x.data_ptr(), N, L, int(K), float(B), int(bool(periodic)),
                                             y.data_ptr(), ld.data_ptr(), s))
        ctx.save_for_backward(logits, x)
        ctx.K, ctx.B, ctx.periodic = int(K), float(B), int(bool(periodic))
        return y, ld

    @staticmethod
    def backward(ctx, gy, gld):
        from ._native import check, lib
        logits, x = ctx.saved_tensors
        N, L = x.shape
        gy, gld = gy.contiguous().float(), gld.contiguous().float()
        gl, gx = torch.empty_like(logits), torch.empty_like(x)
        s = torch.cuda.current_stream(x.device).cuda_stream
        check(lib().pzb_spline_train_backward(logits.data_ptr(), x.data_ptr(), gy.data_ptr(), gld.data_ptr(), N, L,
                                              ctx.K, ctx.B, ctx.periodic, gl.data_ptr(), gx.data_ptr(), s))
        return gl, gx, None, None, None


class TorchChain:
    """(Bijector_Info, params) -> differentiable forward; leaves of the NSC conditioners become torch Parameters
    in the stax layout ``[(W, b), (), (W, b), (), (W, b)]`` so they can be written back into ``Flow._params``."""

    def __init__(self, bijector_info, bijector_params, latent_info, latent_params, input_dim: int, device,
                 dtype=torch.float32, native_spline: bool = True):
        self.native_spline = bool(native_spline) and dtype == torch.float32
        self.info, self.latent_info, self.D = bijector_info, latent_info, int(input_dim)
        self.device, self.dtype = device, dtype
        self._consts: dict = {}
        # every trainable leaf is a view into ONE flat buffer, and its .grad a view into one flat gradient buffer:
        # the data-parallel exchange is a single all_reduce of ``flat_grad`` and Adam is one pass over ``flat``
        sizes: List[int] = []
        self._count(bijector_params, sizes)
        if latent_info[0] == "CentBeta":
            self._count(latent_params, sizes)
        self.flat = torch.zeros(sum(sizes), dtype=dtype, device=device)
        self.flat_grad = torch.zeros_like(self.flat)
        self._off = 0
        self.leaves: List[torch.Tensor] = []
        self.params = self._wrap(bijector_params)
        self.latent_params = self._wrap(latent_params) if latent_info[0] == "CentBeta" else latent_params
        self.attach_grads()

    def _count(self, p, sizes):
        if isinstance(p, (list, tuple)):
            for q in p:
                self._count(q, sizes)
        else:
            sizes.append(int(np.asarray(p).size))

    def _wrap(self, p):
        if isinstance(p, (list, tuple)):
            return type(p)(self._wrap(q) for q in p)
        a = np.asarray(p, dtype=np.float64)
        view = self.flat[self._off:self._off + a.size].view(a.shape)
        view.copy_(torch.as_tensor(a, dtype=self.dtype))
        self._off += a.size
        t = view.requires_grad_(True)
        self.leaves.append(t)
        return t

    def attach_grads(self) -> None:
        """(Re)bind every leaf's .grad to its slice of ``flat_grad`` (autograd then accumulates in place)."""
        off = 0
        for t in self.leaves:
            k = t.numel()
            t.grad = self.flat_grad[off:off + k].view(t.shape)
            off += k

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        if any(t.grad is None or t.grad.data_ptr() != self.flat_grad.data_ptr() + o * self.flat_grad.element_size()
               for t, o in zip(self.leaves, self._offsets())):
            self.attach_grads()

    def _offsets(self):
        off = 0
        for t in self.leaves:
            yield off
            off += t.numel()

    def _const(self, key, make):
        """Device constants of the affine bijectors, built once (no host-to-device copy inside a captured step)."""
        if key not in self._consts:
            self._consts[key] = make()
        return self._consts[key]

    @staticmethod
    def _unwrap(p):
        if isinstance(p, (list, tuple)):
            return type(p)(TorchChain._unwrap(q) for q in p)
        a = p.detach().cpu().numpy().astype(np.float32)
        return a if a.ndim else float(a)

    def numpy_params(self) -> Tuple[Any, Any]:
        lat = self._unwrap(self.latent_params) if self.latent_info[0] == "CentBeta" else self.latent_params
        return lat, self._unwrap(self.params)

    # ---- bijectors ----------------------------------------------------------------------
    def _t(self, a):
        return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=self.dtype, device=self.device)

    def _apply(self, info, params, x, cond):
        name, args = info
        zeros = x.new_zeros(x.shape[0])
        if name == "Chain":                                      # bijectors.py:155-160
            ld = zeros
            for sub, p in zip(args, params):
                x, l2 = self._apply(sub, p, x, cond)
                ld = ld + l2
            return x, ld
        if name == "RollingSplineCoupling":                      # bijectors.py:651-665
            full = tuple(args) + (None, 1, 16, 5, 2, 128, None, 0, False)[len(args):]
            nlayers, shift, K, B, hl, hd, td, nc, periodic = full
            nsc = ("NeuralSplineCoupling", (K, B, hl, hd, td, nc, periodic))
            return self._apply(("Chain", (nsc, ("Roll", (shift,))) * nlayers), params, x, cond)
        if name == "Roll":                                       # bijectors.py:585-593
            return torch.roll(x, int(args[0]), dims=-1), zeros
        if name == "Reverse":
            return torch.flip(x, dims=(-1,)), zeros
        if name == "Shuffle":                                    # bijectors.py:800-804 (permutation from plan.resolve_shuffles)
            return x[:, [int(v) for v in args[0]]], zeros
        if name == "UniformDequantizer":                         # bijectors.py:892-899: identity as executed
            return x, zeros
        if name == "ShiftBounds":                                # bijectors.py:754-765
            B = float(args[2]) if len(args) > 2 else 5.0

            def make():
                mn, mx = self._t(np.atleast_1d(args[0])), self._t(np.atleast_1d(args[1]))
                mean, hr = (mx + mn) / 2, (mx - mn) / 2
                return mean, hr, torch.log(torch.prod(B / hr))
            mean, hr, logdet = self._const((id(info), "sb"), make)
            return B * (x - mean) / hr, zeros + logdet
        if name == "StandardScaler":                             # bijectors.py:848-851
            means, stds = self._const((id(info), "ss"), lambda: (self._t(args[0]), self._t(args[1])))
            return (x - means) / stds, zeros + torch.log(1 / torch.prod(stds))
        if name == "Scale":                                      # bijectors.py:689-697
            s = float(args[0])
            return s * x, zeros + math.log(s ** x.shape[-1])
        if name == "InvSoftplus":                                # bijectors.py:350-358
            idx, k = self._const((id(info), "isp"), lambda: (
                torch.as_tensor(np.atleast_1d(args[0]), device=self.device),
                self._t(np.atleast_1d(args[1] if len(args) > 1 else 1))))
            y = torch.log(-1 + torch.exp(k * x[:, idx])) / k
            out = x.clone()
            out[:, idx] = y
            return out, torch.log(1 + torch.exp(-k * y)).sum(dim=1)
        if name == "ColorTransform":                             # bijectors.py:236-276
            ref_idx = int(args[0])

            def make():   # index tensors live on the device once: an index list would be copied per call (and
                mag = [int(j) for j in np.asarray(args[1]).ravel()]   # cannot be inside a CUDA-graph capture)
                front = [j for j in range(x.shape[1]) if j not in mag]
                return (torch.as_tensor(mag, dtype=torch.long, device=self.device),
                        torch.as_tensor(front, dtype=torch.long, device=self.device))
            mag_idx, front = self._const((id(info), "color", x.shape[1]), make)
            mags = x.index_select(1, mag_idx)
            out = torch.cat([x.index_select(1, front), x[:, ref_idx:ref_idx + 1], -(mags[:, 1:] - mags[:, :-1])], dim=1)
            return out, zeros
        if name == "NeuralSplineCoupling":                       # bijectors.py:462-506
            full = tuple(args) + (16, 5, 2, 128, None, 0, False)[len(args):]
            K, B, hl, hd, td, nc, periodic = full
            D = x.shape[1]
            L = D - D // 2 if td is None else int(td)
            U = D - L
            upper, lower = x[:, :U], x[:, U:]
            key = torch.cat([upper, cond], dim=1)[:, : U + int(nc)] if cond is not None else upper
            h = key
            dense = [p for p in params if len(p) == 2]
            for i, (W, b) in enumerate(dense):                   # utils.py:45-48 (LeakyReLU 0.01)
                h = h @ W + b
                if i + 1 < len(dense):
                    h = torch.nn.functional.leaky_relu(h, LEAKY_SLOPE)
            cpd = 3 * K - 1 + int(bool(periodic))
            h = h.reshape(-1, L, cpd)
            if self.native_spline and h.is_cuda and int(K) <= 64:
                y, ld = _SplineStage.apply(h, lower, int(K), float(B), bool(periodic))
                return torch.cat([upper, y], dim=1), ld.sum(dim=1)
            Wl, Hl, Dl = h[..., :K], h[..., K:2 * K], h[..., 2 * K:]
            Wk = 2 * B * torch.softmax(Wl, dim=-1)               # bijectors.py:489-491
            Hk = 2 * B * torch.softmax(Hl, dim=-1)
            Dk = torch.nn.functional.softplus(Dl)
            y, ld = _rqs_forward(lower, Wk, Hk, Dk, float(B), bool(periodic))
            return torch.cat([upper, y], dim=1), ld
        raise NotImplementedError(f"training: bijector {name!r} is outside the B200 path")

    def _latent(self, u):
        name, args = self.latent_info
        if name == "Uniform":                                    # distributions.py:414-441
            B = float(args[1])
            inside = ((u >= -B) & (u <= B)).all(dim=1)
            val = -u.shape[1] * math.log(2 * B)
            return torch.where(inside, u.new_full((u.shape[0],), val), u.new_full((u.shape[0],), -math.inf))
        if name == "Normal":                                     # distributions.py:255-275
            return -0.5 * (u.shape[1] * math.log(2 * math.pi) + (u * u).sum(dim=1))
        if name in ("CentBeta13", "CentBeta"):                   # distributions.py:165-194 / 70-99
            B = float(args[1])
            loc, scale = -B, 2 * B
            if name == "CentBeta13":
                a = b = u.new_full((u.shape[1],), 13.0)
            else:
                a = torch.stack([torch.exp(p[0]) for p in self.latent_params])
                b = torch.stack([torch.exp(p[1]) for p in self.latent_params])
            betaln = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
            y = (u - loc) / scale
            lp = -betaln + torch.xlogy(a - 1, y) + torch.special.xlog1py(b - 1, -y) - math.log(scale)
            lp = torch.where((u > loc + scale) | (u < loc), torch.full_like(lp, -math.inf), lp)
            return lp.sum(dim=1)
        raise NotImplementedError(f"training: latent {name!r} is outside the B200 path")

    def _latent_any(self, info, params, u):
        """Tdist (distributions.py:330-358, cov = I) and Joint (528-541: the sum over its sub-distributions, each on
        its own consecutive columns) on top of the closed forms of ``_latent``."""
        name, args = info
        if name == "Joint":
            subs = args[1]
            plist = params if params is not None else [None] * len(subs)
            total, start = None, 0
            for sub, p in zip(subs, plist):
                sub = tuple(sub)
                n = int(sub[1][0])
                lp = self._latent_any(sub, p, u[:, start:start + n])
                total = lp if total is None else total + lp
                start += n
            return total
        if name == "Tdist":
            n = u.shape[1]
            log_nu = params if params is not None else math.log(30.0)
            nu = torch.exp(torch.as_tensor(np.asarray(log_nu, dtype=np.float64), dtype=self.dtype, device=self.device)
                           if not isinstance(log_nu, torch.Tensor) else log_nu).reshape(())
            t = 0.5 * (nu + n)
            return (torch.lgamma(t) - torch.lgamma(0.5 * nu) - 0.5 * n * torch.log(nu * math.pi)
                    - t * torch.log1p((u * u).sum(dim=1) / nu))
        saved = self.latent_info, self.latent_params
        try:
            self.latent_info, self.latent_params = info, params
            return self._latent(u)
        finally:
            self.latent_info, self.latent_params = saved

    def log_prob(self, x, cond=None):
        """Flow._log_prob (flow.py:397-407) with gradients."""
        u, ld = self._apply(self.info, self.params, x, cond)
        lat = (self._latent_any(self.latent_info, self.latent_params, u) if self.latent_info[0] in ("Tdist", "Joint")
               else self._latent(u))
        return torch.nan_to_num(lat + ld, nan=-math.inf, posinf=math.inf, neginf=-math.inf)


# --------------------------------------------------------------------------- #
# one optimisation step, data parallel                                         #
# --------------------------------------------------------------------------- #
def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous share of an n-row batch for ``rank`` (the same split as pzflow_b200.sharding.shard_rows)."""
    base, rem = divmod(n, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


class FusedGradient:
    """The gradient of one step as ONE hand-written kernel (``pzb_train_step``, csrc/nsf_train_step.cu): forward and
    backward through the whole chain, a CTA per 16 rows, per-CTA partial weight gradients summed in a fixed order.
    Reads the parameters from ``chain.flat`` and writes ``chain.flat_grad`` -- no autograd graph, no library GEMM.
    ``FusedGradient.build`` returns None for chains outside the kernel's shapes (they keep the autograd route)."""

    def __init__(self, chain: "TorchChain", plan, n_params: int):
        self.chain, self.plan, self.n_params = chain, plan, n_params
        self._ws = {}
        self.loss = torch.zeros(1, dtype=torch.float32, device=chain.device)

    @staticmethod
    def _dense_offsets(info, params, flat, out):
        """(W offset, b offset) of every dense layer, coupling stages in chain order (the order plan.flatten walks)."""
        name, args = info
        if name == "Chain":
            for sub, p in zip(args, params):
                FusedGradient._dense_offsets(sub, p, flat, out)
        elif name == "RollingSplineCoupling":
            full = tuple(args) + (None, 1, 16, 5, 2, 128, None, 0, False)[len(args):]
            nlayers, shift, K, B, hl, hd, td, nc, periodic = full
            nsc = ("NeuralSplineCoupling", (K, B, hl, hd, td, nc, periodic))
            FusedGradient._dense_offsets(("Chain", (nsc, ("Roll", (shift,))) * nlayers), params, flat, out)
        elif name == "NeuralSplineCoupling":
            for layer in params:
                if len(layer) == 2:
                    W, b = layer
                    out.append(((W.data_ptr() - flat.data_ptr()) // flat.element_size(),
                                (b.data_ptr() - flat.data_ptr()) // flat.element_size()))

    @classmethod
    def build(cls, chain: "TorchChain"):
        import ctypes as C
        from . import _native as N
        from .plan import Plan
        if (chain.dtype != torch.float32 or not chain.flat.is_cuda or chain.latent_info[0] == "CentBeta"
                or os.environ.get("PZB_TRAIN_AUTOGRAD", "0") != "0"):
            return None
        try:
            lat, bij = chain.numpy_params()
            with torch.cuda.device(chain.device):
                plan = Plan(chain.info, bij, chain.D, chain.latent_info, device=chain.device.index, latent_params=lat)
        except (NotImplementedError, ValueError):
            return None
        if not N.lib().pzb_plan_train_supported(plan._h):
            return None
        offs = []
        cls._dense_offsets(chain.info, chain.params, chain.flat, offs)
        w = (C.c_int64 * len(offs))(*[o[0] for o in offs])
        b = (C.c_int64 * len(offs))(*[o[1] for o in offs])
        N.check(N.lib().pzb_plan_set_train_layout(plan._h, w, b, len(offs), chain.flat.numel()))
        return cls(chain, plan, chain.flat.numel())

    def _workspace(self, rows: int) -> torch.Tensor:
        from ._native import lib
        ws = self._ws.get(rows)
        if ws is None:
            need = lib().pzb_train_workspace_floats(self.plan._h, rows)
            ws = self._ws[rows] = torch.empty(max(1, need), dtype=torch.float32, device=self.chain.device)
        return ws

    def __call__(self, x, cond, w, n: int) -> None:
        """flat_grad <- gradient of -sum(w * log_prob(x | cond)) / n for the given batch rows."""
        self.gathered(x, cond, w, None, 0, int(x.shape[0]), n)

    def gathered(self, X, cond, w, idx, start: int, rows: int, n: int, Xerr=None, cond_err=None, seed: int = 0,
                 noise_base: int = 0) -> None:
        """The same for the batch rows ``idx[start:start + rows]`` of the WHOLE training arrays X / cond / w (the
        epoch's permutation is applied inside the kernel), optionally convolved with Gaussian errors drawn in the
        kernel (flow.py:1063-1079): a step launches no gather, no random-number and no copy kernel."""
        from ._native import check, lib
        c = self.chain
        nc = self.plan.n_conditions
        for t in (X, w) + ((cond,) if nc else ()) + ((Xerr,) if Xerr is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("the fused training step needs contiguous float32 device tensors")
        if nc and cond.shape[1] != nc:
            raise ValueError(f"conditions must have {nc} columns")
        if idx is not None and not (idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous()):
            raise ValueError("row indices must be a contiguous int64 device tensor")
        ws = self._workspace(rows)
        check(lib().pzb_train_step(
            self.plan._h, c.flat.data_ptr(), X.data_ptr() if rows else None,
            cond.data_ptr() if (nc and rows) else None, w.data_ptr() if rows else None, rows, int(n),
            idx.data_ptr() + 8 * int(start) if idx is not None else None,
            Xerr.data_ptr() if Xerr is not None else None,
            cond_err.data_ptr() if (cond_err is not None and nc) else None, int(seed) & (2 ** 64 - 1),
            int(noise_base) & (2 ** 64 - 1), ws.data_ptr(), c.flat_grad.data_ptr(), self.loss.data_ptr(),
            torch.cuda.current_stream(c.device).cuda_stream))


def _grads(chain: TorchChain, x, cond, w, n: int) -> None:
    """flat_grad <- d/d params of -sum_local(w * log_prob) / n, non-finite entries zeroed, summed over ranks."""
    fused = getattr(chain, "fused_gradient", None)
    if fused is not None:
        nc = fused.plan.n_conditions
        fused(x.contiguous().float(), cond[:, :nc].contiguous().float() if (nc and cond is not None) else None,
              w.contiguous().float(), n)
        d = _dist()
        if d is not None:
            d.all_reduce(chain.flat_grad, op=d.ReduceOp.SUM)
        return
    chain.zero_grad()
    if x.shape[0] > 0:
        lp = chain.log_prob(x, cond)
        fin = torch.isfinite(lp)
        loss = -(torch.where(fin, w * lp, torch.zeros_like(lp))).sum() / n
        loss.backward()
    torch.nan_to_num_(chain.flat_grad, nan=0.0, posinf=0.0, neginf=0.0)
    d = _dist()
    if d is not None:
        d.all_reduce(chain.flat_grad, op=d.ReduceOp.SUM)


def dp_step(chain: TorchChain, opt, x, cond, w, global_rows: Optional[int] = None) -> None:
    """grad of -mean_over_the_GLOBAL_batch(w * log_prob) on this rank's rows, one all_reduce(sum) of the flat
    gradient, Adam update (flow.py:961-981).  Rows whose log_prob is -inf carry no gradient (jnp.nan_to_num).
    ``opt`` is a torch optimizer over ``chain.leaves`` or a ``NativeAdam``."""
    _grads(chain, x, cond, w, int(x.shape[0]) if global_rows is None else int(global_rows))
    opt.step()


class NativeAdam:
    """optax.adam(learning_rate) (flow.py:968-981) as one fused kernel over the chain's flat parameter buffer
    (``pzb_adam_step``); moments and the step count live on the device, so a step can sit inside a CUDA graph."""

    def __init__(self, chain: TorchChain, lr: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        if chain.flat.dtype != torch.float32 or not chain.flat.is_cuda:
            raise ValueError("NativeAdam needs float32 parameters on a CUDA device")
        self.chain, self.lr, self.b1, self.b2, self.eps = chain, lr, b1, b2, eps
        self.m, self.v = torch.zeros_like(chain.flat), torch.zeros_like(chain.flat)
        self.count = torch.zeros(1, dtype=torch.int32, device=chain.flat.device)

    def step(self) -> None:
        from ._native import check, lib
        c = self.chain
        check(lib().pzb_adam_step(c.flat.data_ptr(), c.flat_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                  c.flat.numel(), self.lr, self.b1, self.b2, self.eps, self.count.data_ptr(),
                                  torch.cuda.current_stream(c.flat.device).cuda_stream))


class GraphedStep:
    """One optimisation step -- zero the flat gradient, forward, loss, backward, nan_to_num, all_reduce, Adam --
    captured ONCE as a CUDA graph for a fixed local batch shape and replayed per batch: a step of the default 7-D
    flow is ~300 kernels of a few microseconds each, so launching them one by one from Python is what bounds an
    eager step.  Inputs are copied into static buffers; everything else (parameters, gradients, Adam moments,
    step count) already lives at fixed device addresses."""

    WARMUP = 2   # eager steps before capture (cuBLAS workspaces, NCCL communicator, cached constants)

    def __init__(self, chain: TorchChain, opt: NativeAdam, rows: int, n_cols: int, n_cond: int, global_rows: int):
        dev = chain.device
        self.chain, self.opt, self.rows, self.global_rows = chain, opt, int(rows), int(global_rows)
        self.x = torch.zeros(rows, n_cols, device=dev)
        self.c = torch.zeros(rows, n_cond, device=dev) if n_cond else None
        self.w = torch.zeros(rows, device=dev)
        self.graph, self.seen = None, 0

    def _load(self, x, cond, w):
        self.x.copy_(x)
        self.w.copy_(w)
        if self.c is not None:
            self.c.copy_(cond)

    def __call__(self, x, cond, w) -> None:
        self._load(x, cond, w)
        if self.graph is not None:
            self.graph.replay()
            return
        if self.seen < self.WARMUP:
            self.seen += 1
            dp_step(self.chain, self.opt, self.x, self.c, self.w, self.global_rows)
            return
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            dp_step(self.chain, self.opt, self.x, self.c, self.w, self.global_rows)
        self.graph = g
        g.replay()            # capture records without executing: this replay IS the step for this batch


def train_flow(flow, inputs, val_set=None, train_weight=None, val_weight=None, epochs: int = 100,
               batch_size: int = 1024, optimizer=None, loss_fn=None, convolve_errs: bool = False,
               patience: int = None, best_params: bool = True, seed: int = 0, verbose: bool = False,
               progress_bar: bool = False, initial_loss: bool = True) -> list:
    """Body of Flow.train (pzflow/flow.py:936-1159)."""
    import pandas as pd  # noqa: F401
    if not torch.cuda.is_available():
        from ._native import NativeError
        raise NativeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
    if loss_fn is not None:
        raise NotImplementedError("a custom loss_fn is a jax callable; only the default -mean(w * log_prob) is supported")
    if optimizer is not None and not callable(optimizer):
        raise NotImplementedError("pass optimizer=None (Adam 1e-3, as optax.adam) or a callable leaves -> torch optimizer")
    rng = np.random.default_rng(seed)
    batch_seed, bijector_seed = rng.integers(1e9, size=2)       # flow.py:937-938
    if flow._bijector_info is None:
        flow._set_default_bijector(inputs, seed=int(bijector_seed))
    if not isinstance(epochs, int) or epochs <= 0:
        raise ValueError("epochs must be a positive integer.")
    dev = torch.device("cuda", torch.cuda.current_device())
    d = _dist()
    rank, world = (d.get_rank(), d.get_world_size()) if d is not None else (0, 1)

    columns = list(flow.data_columns)
    if flow.conditional_columns is not None and flow._autoscale_conditions:   # flow.py:989-998
        c = inputs[list(flow.conditional_columns)].to_numpy()
        flow._condition_means = c.mean(axis=0).astype(np.float32)
        stds = c.std(axis=0)
        flow._condition_stds = np.where(stds != 0, stds, 1).astype(np.float32)

    from .plan import resolve_shuffles
    info = resolve_shuffles(flow._bijector_info, getattr(flow, "_bijector_seed", 0), flow._input_dim)
    chain = TorchChain(info, flow._params[1], flow._latent_info, flow._params[0], flow._input_dim, dev)
    chain.fused_gradient = FusedGradient.build(chain)   # one kernel per step where the chain's shapes allow it
    opt = NativeAdam(chain, lr=1e-3) if optimizer is None else optimizer(chain.leaves)
    graphed: dict = {}

    def to_dev(a):
        return None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)

    def arrays(df):
        if convolve_errs:
            return flow._get_err_arrays(df, "data"), flow._get_err_arrays(df, "conditions")
        return (df[columns].to_numpy().astype(np.float32), None), (flow._get_conditions(df), None)

    def fused_loss(X, C, W):
        """-mean(w * log_prob) through the fused CUDA kernels with the CURRENT parameters."""
        flow._params = chain.numpy_params()
        flow._plan_cache.clear()
        lp = flow._plan().log_prob(torch.as_tensor(X, device=dev), C if C is None else torch.as_tensor(C, device=dev))
        # Zero-probability rows come back as finfo.min, as from the reference's nan_to_num (pzflow/flow.py:406), and enter
        # the float32 mean as such: ONE such row makes the loss huge but finite (~1e38 / n) and training goes on -- the
        # reference's own test_patience relies on it -- while a batch of them overflows the sum to inf and ends the run as
        # diverged (its test_nan_train_stop).
        return float(-(W * lp).mean().item())

    n = len(inputs)
    W = np.ones(n, np.float32) if train_weight is None else np.asarray(train_weight, np.float32)
    W = W / W.mean()
    X_all = inputs[columns].to_numpy().astype(np.float32)
    C_all = flow._get_conditions(inputs)
    Wd = to_dev(W)
    losses = [fused_loss(X_all, C_all, Wd)] if initial_loss else []
    if val_set is not None:
        Xv, Cv = val_set[columns].to_numpy().astype(np.float32), flow._get_conditions(val_set)
        Wv = np.ones(len(val_set), np.float32) if val_weight is None else np.asarray(val_weight, np.float32)
        Wv = to_dev(Wv / Wv.mean())
        val_losses = [fused_loss(Xv, Cv, Wv)] if initial_loss else []
    if verbose:
        print(f"Training {epochs} epochs \nLoss:")
        if initial_loss:
            print(f"(0) {losses[-1]:.4f}" + ("" if val_set is None else f"  {val_losses[-1]:.4f}"))

    (Xm, Xe), (Cm, Ce) = arrays(inputs)
    Xm, Xe, Cm, Ce = to_dev(Xm), to_dev(Xe), to_dev(Cm), to_dev(Ce)
    perm_rng = np.random.default_rng(int(batch_seed))
    # The reference's epoch schedule (flow.py:1015, 1062-1064): key = PRNGKey(batch_seed); every epoch
    # permute_key, sample_key, key = split(key, 3) and idx = permutation(permute_key, n) -- reproduced on the host
    # (pzflow_b200.prng) while that stays cheap; beyond 4 M rows the epochs are shuffled by NumPy's generator instead.
    from . import prng as _prng
    epoch_key = _prng.key_from_seed(int(batch_seed)) if (_prng.layout() != "philox" and n <= (1 << 20)) else None

    def jax_epoch_orders():
        """The reference's permutation of every epoch (prng.epoch_orders), each computed one epoch ahead on a helper
        thread (NumPy releases the GIL) so that the host-side threefry + sort hides behind the previous epoch's kernels."""
        from concurrent.futures import ThreadPoolExecutor
        orders = _prng.epoch_orders(int(batch_seed), n)
        with ThreadPoolExecutor(1) as pool:
            pending = pool.submit(next, orders)
            while True:
                order = pending.result()
                pending = pool.submit(next, orders)
                yield order
    epoch_orders = jax_epoch_orders() if epoch_key is not None else None
    noise = torch.Generator(device=dev)
    noise.manual_seed(int(batch_seed) + rank)   # every rank draws its own eps for its own share of the batch
    best_loss, best_vals, stall = math.inf, chain.numpy_params(), 0
    fused = chain.fused_gradient
    use_graph = fused is None and isinstance(opt, NativeAdam) and os.environ.get("PZB_TRAIN_GRAPH", "1") != "0"
    if fused is not None:   # whole arrays as the kernel wants them; the batch is picked inside the kernel
        nc = fused.plan.n_conditions
        Xg, Wg = Xm.contiguous().float(), Wd.contiguous().float()
        Cg = Cm[:, :nc].contiguous().float() if (nc and Cm is not None) else None
        Xeg = Xe.contiguous().float() if Xe is not None else None
        Ceg = Ce[:, :nc].contiguous().float() if (nc and Ce is not None) else None
    steps_done = 0
    loop = range(epochs)
    if progress_bar:
        from tqdm import tqdm
        loop = tqdm(loop)
    for epoch in loop:
        if epoch_orders is not None:
            idx = torch.as_tensor(next(epoch_orders), device=dev)              # flow.py:1062-1064
        else:
            idx = torch.as_tensor(perm_rng.permutation(n), device=dev)
        # a chain without parameters (e.g. the reference's own Flow(columns, Reverse()) test flows) has nothing to step:
        # its epochs only record the (constant) loss, and patience / divergence end them as in the reference
        for b0 in (range(0, n, batch_size) if chain.leaves else ()):
            sel = idx[b0:b0 + batch_size]
            a, b = shard_batch(len(sel), rank, world)
            if fused is not None:
                # one kernel for the gradient (gather, error samples, forward, backward), one for the partial sums,
                # the all-reduce of the flat gradient, Adam
                fused.gathered(Xg, Cg, Wg, idx, b0 + a, b - a, len(sel), Xeg, Ceg, seed=int(batch_seed),
                               noise_base=steps_done * batch_size + a)
                if d is not None:
                    d.all_reduce(chain.flat_grad, op=d.ReduceOp.SUM)
                opt.step()
                steps_done += 1
                continue
            mine = sel[a:b]
            xb = Xm[mine]
            if Xe is not None:
                xb = xb + torch.randn(xb.shape, generator=noise, device=dev) * Xe[mine]
            cb = None if Cm is None else Cm[mine]
            if cb is not None and Ce is not None:
                cb = cb + torch.randn(cb.shape, generator=noise, device=dev) * Ce[mine]
            if use_graph and len(sel) == batch_size and len(mine) > 0:
                key = len(mine)
                if key not in graphed:
                    graphed[key] = GraphedStep(chain, opt, key, xb.shape[1], 0 if cb is None else cb.shape[1],
                                               len(sel))
                graphed[key](xb, cb, Wd[mine])
            else:                                   # ragged last batch of the epoch, or a caller's own optimizer
                dp_step(chain, opt, xb, cb, Wd[mine], global_rows=len(sel))
        losses.append(fused_loss(X_all, C_all, Wd))                             # flow.py:1088-1098
        if val_set is not None:
            val_losses.append(fused_loss(Xv, Cv, Wv))
        if verbose and (epoch % max(int(0.05 * epochs), 1) == 0 or (epoch + 1) == epochs):
            print(f"({epoch + 1}) {losses[-1]:.4f}" + ("" if val_set is None else f"  {val_losses[-1]:.4f}"))
        tracked = losses if val_set is None else val_losses                     # flow.py:1113-1141
        if tracked[-1] >= best_loss or np.isclose(tracked[-1], best_loss):
            stall += 1
            if patience is not None and stall >= patience:
                print("Early stopping criterion is met.", f"Training stopping after epoch {epoch + 1}.")
                break
        else:
            best_loss, best_vals, stall = tracked[-1], chain.numpy_params(), 0
        if not np.isfinite(losses[-1]):
            print(f"Training stopping after epoch {epoch + 1}", "because training loss diverged.")
            break
    if epoch_orders is not None:
        epoch_orders.close()
    flow._params = best_vals if best_params else chain.numpy_params()            # flow.py:1151-1154
    flow._plan_cache.clear()
    return losses if val_set is None else [losses, val_losses]

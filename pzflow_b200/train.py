"""``Flow.train`` (pzflow/flow.py:868-1159): the optional data-parallel training step.

What runs where
---------------
* The inference hot path (log_prob / posterior / sample) is the hand-written CUDA of ``csrc/``.  Training is
  NOT on that path: the gradient of ``-mean(w * log_prob)`` (flow.py:961-965) is taken by torch autograd over a
  differentiable restatement of the same chain in torch tensor ops on the GPU (cuBLAS GEMMs for the
  conditioners), and the update is Adam(1e-3) like ``optax.adam`` (flow.py:968-981).  It is a library-level
  baseline, stated as such in DESIGN.md; the fused kernels evaluate every reported loss (initial, end of epoch,
  validation: flow.py:1033-1050, 1088-1102), so a trained flow is checked by the product path as it trains.
* Data parallel: one process per GPU; each rank takes its contiguous share of every batch (batch 1024 over 8
  ranks = 128 rows each), the flat fp32 gradient is summed with ONE ``all_reduce`` (NCCL over NVLink on GPUs)
  per step, and every rank applies the same Adam update, so parameters stay bit-identical across ranks.

Not reproducible from the reference without jax: its batch permutations and error samples come from
``jax.random`` (flow.py:1063-1066); here they come from NumPy / torch generators seeded from the same ``seed``.
"""
from __future__ import annotations

import math
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


class TorchChain:
    """(Bijector_Info, params) -> differentiable forward; leaves of the NSC conditioners become torch Parameters
    in the stax layout ``[(W, b), (), (W, b), (), (W, b)]`` so they can be written back into ``Flow._params``."""

    def __init__(self, bijector_info, bijector_params, latent_info, latent_params, input_dim: int, device,
                 dtype=torch.float32):
        self.info, self.latent_info, self.D = bijector_info, latent_info, int(input_dim)
        self.device, self.dtype = device, dtype
        self.leaves: List[torch.Tensor] = []
        self.params = self._wrap(bijector_params)
        self.latent_params = self._wrap(latent_params) if latent_info[0] == "CentBeta" else latent_params

    def _wrap(self, p):
        if isinstance(p, (list, tuple)):
            return type(p)(self._wrap(q) for q in p)
        t = torch.tensor(np.asarray(p, dtype=np.float64), dtype=self.dtype, device=self.device, requires_grad=True)
        self.leaves.append(t)
        return t

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
        if name == "ShiftBounds":                                # bijectors.py:754-765
            mn, mx = self._t(np.atleast_1d(args[0])), self._t(np.atleast_1d(args[1]))
            B = float(args[2]) if len(args) > 2 else 5.0
            mean, hr = (mx + mn) / 2, (mx - mn) / 2
            return B * (x - mean) / hr, zeros + torch.log(torch.prod(B / hr))
        if name == "StandardScaler":                             # bijectors.py:848-851
            means, stds = self._t(args[0]), self._t(args[1])
            return (x - means) / stds, zeros + torch.log(1 / torch.prod(stds))
        if name == "Scale":                                      # bijectors.py:689-697
            s = float(args[0])
            return s * x, zeros + math.log(s ** x.shape[-1])
        if name == "InvSoftplus":                                # bijectors.py:350-358
            idx = torch.as_tensor(np.atleast_1d(args[0]), device=self.device)
            k = self._t(np.atleast_1d(args[1] if len(args) > 1 else 1))
            y = torch.log(-1 + torch.exp(k * x[:, idx])) / k
            out = x.clone()
            out[:, idx] = y
            return out, torch.log(1 + torch.exp(-k * y)).sum(dim=1)
        if name == "ColorTransform":                             # bijectors.py:236-276
            ref_idx, mag_idx = int(args[0]), list(np.asarray(args[1]).ravel().astype(int))
            front = [j for j in range(x.shape[1]) if j not in mag_idx]
            mags = x[:, mag_idx]
            out = torch.cat([x[:, front], x[:, ref_idx:ref_idx + 1], -(mags[:, 1:] - mags[:, :-1])], dim=1)
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

    def log_prob(self, x, cond=None):
        """Flow._log_prob (flow.py:397-407) with gradients."""
        u, ld = self._apply(self.info, self.params, x, cond)
        return torch.nan_to_num(self._latent(u) + ld, nan=-math.inf, posinf=math.inf, neginf=-math.inf)


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


def dp_step(chain: TorchChain, opt: torch.optim.Optimizer, x, cond, w, global_rows: Optional[int] = None) -> None:
    """grad of -mean_over_the_GLOBAL_batch(w * log_prob) on this rank's rows, one all_reduce(sum) of the flat
    gradient, Adam update (flow.py:961-981).  Rows whose log_prob is -inf carry no gradient (jnp.nan_to_num)."""
    n = int(x.shape[0]) if global_rows is None else int(global_rows)
    opt.zero_grad(set_to_none=True)
    if x.shape[0] > 0:
        lp = chain.log_prob(x, cond)
        fin = torch.isfinite(lp)
        loss = -(torch.where(fin, w * lp, torch.zeros_like(lp))).sum() / n
        loss.backward()
    d = _dist()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in chain.leaves])
    flat = torch.nan_to_num(flat, nan=0.0, posinf=0.0, neginf=0.0)
    if d is not None:
        d.all_reduce(flat, op=d.ReduceOp.SUM)
    off = 0
    for p in chain.leaves:
        k = p.numel()
        p.grad = flat[off:off + k].reshape(p.shape).clone()
        off += k
    opt.step()


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

    chain = TorchChain(flow._bijector_info, flow._params[1], flow._latent_info, flow._params[0], flow._input_dim, dev)
    opt = torch.optim.Adam(chain.leaves, lr=1e-3) if optimizer is None else optimizer(chain.leaves)

    def to_dev(a):
        return None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)

    def arrays(df):
        if convolve_errs:
            return flow._get_err_arrays(df, "data"), flow._get_err_arrays(df, "conditions")
        return (df[columns].to_numpy().astype(np.float32), None), (flow._get_conditions(df), None)

    def fused_loss(X, C, W):
        """-mean(w * log_prob) through the fused CUDA kernels with the CURRENT parameters."""
        flow._params = chain.numpy_params()
        flow._plan_cache = {}
        lp = flow._plan().log_prob(torch.as_tensor(X, device=dev), C if C is None else torch.as_tensor(C, device=dev))
        lp = torch.where(lp < -1e30, torch.full_like(lp, -math.inf), lp)   # finfo.min is the zero-probability class
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
    noise = torch.Generator(device=dev)
    noise.manual_seed(int(batch_seed))
    best_loss, best_vals, stall = math.inf, chain.numpy_params(), 0
    loop = range(epochs)
    if progress_bar:
        from tqdm import tqdm
        loop = tqdm(loop)
    for epoch in loop:
        idx = torch.as_tensor(perm_rng.permutation(n), device=dev)            # flow.py:1063-1066
        for b0 in range(0, n, batch_size):
            sel = idx[b0:b0 + batch_size]
            a, b = shard_batch(len(sel), rank, world)
            mine = sel[a:b]
            xb = Xm[mine]
            if Xe is not None:
                xb = xb + torch.randn(xb.shape, generator=noise, device=dev) * Xe[mine]
            cb = None if Cm is None else Cm[mine]
            if cb is not None and Ce is not None:
                cb = cb + torch.randn(cb.shape, generator=noise, device=dev) * Ce[mine]
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
    flow._params = best_vals if best_params else chain.numpy_params()            # flow.py:1151-1154
    flow._plan_cache = {}
    return losses if val_set is None else [losses, val_losses]

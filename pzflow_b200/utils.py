"""Host glue of the hot path (mirror of pzflow/utils.py:11-19)."""
from __future__ import annotations

from . import bijectors


def build_bijector_from_info(info: tuple) -> tuple:
    """Build a Bijector from a Bijector_Info object (pzflow/utils.py:11-19)."""
    if info[0] == "Chain":
        return bijectors.Chain(*(build_bijector_from_info(i) for i in info[1]))
    try:
        factory = getattr(bijectors, info[0])
    except AttributeError:
        raise NotImplementedError(f"bijector {info[0]!r} is outside the B200 hot path") from None
    return factory(*info[1])


def gaussian_error_model(key, X, Xerr, nsamples):
    """Default Gaussian error model (pzflow/utils.py:52-61): X[:, None, :] + eps * Xerr[:, None, :],
    eps ~ N(0, 1) of shape [N, nsamples, D].  ``key``: an int seed or the uint32 pair jax.random.PRNGKey(seed) would be.
    In the default stream layout (pzflow_b200.prng) eps IS ``jax.random.normal(key, (N, nsamples, D))``, generated on
    the GPU; under the "philox" layout it is the library's counter-based stream (pzflow_b200.plan.error_eps).  Either
    way these are the draws the fused kernels generate on the fly -- Flow.log_prob / posterior never call this
    function, they only check that the flow's error model IS this one and let the kernel do the expansion."""
    import numpy as np

    from . import prng
    from .plan import _error_stream, error_eps, sample_normal_jax
    X = np.asarray(X, dtype=np.float32)
    Xerr = np.asarray(Xerr, dtype=np.float32)
    k = np.asarray(key).astype(np.uint64).ravel()   # an int seed, or the uint32 pair PRNGKey(seed) would be
    n, S, W = X.shape[0], int(nsamples), X.shape[1]
    if _error_stream(n, S, W):
        eps = sample_normal_jax(n * S, W, prng.key_from_seed(np.asarray(key, dtype=np.uint32)) if k.size == 2 else int(k[0]))
    else:
        seed = int(k[0]) if k.size == 1 else (int(k[0]) << 32) | int(k[1])
        eps = error_eps(n * S, W, 0, seed)
    eps = eps.cpu().numpy().reshape(n, S, W)
    return X[:, None, :] + eps * Xerr[:, None, :]


def RationalQuadraticSpline(inputs, W, H, D, B: float, periodic: bool = False, inverse: bool = False):
    """Apply rational quadratic spline to inputs and return outputs with log_det (pzflow/utils.py:64-227).

    inputs [N, L]; W, H [N, L, K] bin widths / heights (each summing to 2B); D [N, L, K - 1] knot derivatives
    (K if periodic).  Evaluated on the GPU by pzb_rational_quadratic_spline -- same bin rule, out-of-bounds identity
    and NaN beyond the last knot as inside the fused chain kernels.  Device tensors in -> device tensors out; anything
    else -> NumPy arrays."""
    import numpy as np
    import torch

    from . import _native as N
    if not torch.cuda.is_available():
        raise N.NativeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
    as_tensor = isinstance(inputs, torch.Tensor)
    dev = inputs.device if as_tensor and inputs.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def prep(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return t.to(device=dev, dtype=torch.float32).contiguous()
    x, w, h, d = prep(inputs), prep(W), prep(H), prep(D)
    if x.dim() != 2 or w.dim() != 3 or w.shape[:2] != x.shape or h.shape != w.shape:
        raise ValueError("RationalQuadraticSpline: inputs [N, L], W and H [N, L, K] expected")
    n, L, K = w.shape
    if tuple(d.shape) != (n, L, K - 1 + int(bool(periodic))):
        raise ValueError(f"RationalQuadraticSpline: D must have shape {(n, L, K - 1 + int(bool(periodic)))}")
    out = torch.empty_like(x)
    ld = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().pzb_rational_quadratic_spline(x.data_ptr(), w.data_ptr(), h.data_ptr(), d.data_ptr(), n, L, K,
                                                      float(B), int(bool(periodic)), int(bool(inverse)), out.data_ptr(),
                                                      ld.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return (out, ld) if as_tensor else (out.cpu().numpy(), ld.cpu().numpy())


def DenseReluNetwork(out_dim: int, hidden_layers: int, hidden_dim: int):
    """Create a dense neural network with Relu after hidden layers (pzflow/utils.py:22-49): stax-style
    ``(init_fun, apply_fun)`` of ``serial(*(Dense(hidden_dim), LeakyRelu) * hidden_layers, Dense(out_dim))`` -- LeakyReLU
    with slope 0.01, despite the name.  Inside a flow the conditioner is part of the fused chain kernels (tcgen05); this
    stand-alone form exists for API parity and evaluates with library GEMMs on the device (fp32, TF32 off)."""
    import numpy as np
    import torch

    from .bijectors import _dense_init, _generator

    def init_fun(rng, input_shape):
        params = _dense_init(_generator(rng), int(input_shape[-1]), hidden_layers, hidden_dim, out_dim)
        return tuple(input_shape[:-1]) + (out_dim,), params

    def apply_fun(params, inputs, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("pzflow_b200 needs a CUDA device: there is no CPU fallback")
        as_tensor = isinstance(inputs, torch.Tensor)
        dev = inputs.device if as_tensor and inputs.is_cuda else torch.device("cuda", torch.cuda.current_device())
        h = torch.as_tensor(np.asarray(inputs, dtype=np.float32) if not as_tensor else inputs).to(dev, torch.float32)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            for layer in params:
                if len(layer) == 0:
                    h = torch.where(h >= 0, h, 0.01 * h)
                else:
                    Wl, bl = (torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev) if not isinstance(a, torch.Tensor)
                              else a.to(dev, torch.float32) for a in layer)
                    h = h @ Wl + bl
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        return h if as_tensor else h.cpu().numpy()

    return init_fun, apply_fun


def sub_diag_indices(inputs):
    """Return indices for diagonal of 2D blocks in 3D array (pzflow/utils.py:230-243)."""
    import numpy as np
    inputs = np.asarray(inputs)
    if inputs.ndim != 3:
        raise ValueError("Input must be a 3D array.")
    per_block = min(inputs.shape[1:])
    block, along = np.divmod(np.arange(inputs.shape[0] * per_block), per_block)   # (block, k, k) for every diagonal entry
    return block, along, along

"""Wall-clock time of one Flow.train optimisation step (gradient + Adam), batch 1024, per chain and route:
the fused kernel (pzb_train_step), the autograd route launched eagerly, and the autograd route replayed as a CUDA graph.
Run on the B200: python profiles/time_train_step.py > profiles/r02_train_step_timing.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pzflow_b200 import Flow                                                                  # noqa: E402
from pzflow_b200.bijectors import (Chain, ColorTransform, InvSoftplus, RollingSplineCoupling,  # noqa: E402
                                   StandardScaler)
from pzflow_b200.examples import get_example_flow                                             # noqa: E402
from pzflow_b200.plan import launch_count                                                     # noqa: E402
from pzflow_b200.train import FusedGradient, GraphedStep, NativeAdam, TorchChain, dp_step     # noqa: E402


def chains():
    rng = np.random.default_rng(7)
    flow = get_example_flow()
    u = torch.rand((1024, 7), device="cuda") * 8.0 - 4.0
    rows = flow._plan().inverse(u, want_log_det=False)[0].cpu().numpy()      # rows the shipped flow itself would draw
    yield "shipped 7-D flow (ShiftBounds, RollingSplineCoupling(7))", flow, rows
    cols = ("redshift", "u", "g", "r", "i", "z", "y")
    flow = Flow(data_columns=cols, bijector=Chain(
        ColorTransform(3, [1, 2, 3, 4, 5, 6]), InvSoftplus(0, 4.0),
        StandardScaler(np.linspace(-0.5, 0.5, 7).astype(np.float32), np.linspace(0.5, 2.0, 7).astype(np.float32)),
        RollingSplineCoupling(7)))
    yield "photo-z chain (ColorTransform, InvSoftplus, StandardScaler, RollingSplineCoupling(7)), package initialiser", \
        flow, rng.uniform(0.05, 1.5, size=(1024, 7)).astype(np.float32)


def main():
    for name, flow, rows in chains():
        info, bparams, latent, lat_params = flow._bijector_info, flow._params[1], flow._latent_info, flow._params[0]
        X = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).cuda()
        w = torch.ones(len(X), device="cuda")
        print(name)
        for mode in ("fused", "autograd", "autograd+graph"):
            chain = TorchChain(info, bparams, latent, lat_params, 7, torch.device("cuda"))
            chain.fused_gradient = FusedGradient.build(chain) if mode == "fused" else None
            assert (chain.fused_gradient is not None) == (mode == "fused")
            opt = NativeAdam(chain)
            if mode == "autograd+graph":
                step = GraphedStep(chain, opt, len(X), 7, 0, len(X))
            else:
                def step(x, c, ww, chain=chain, opt=opt):
                    dp_step(chain, opt, x, c, ww)
            for _ in range(5):
                step(X, None, w)
            torch.cuda.synchronize()
            n0, t = launch_count(), time.perf_counter()
            for _ in range(50):
                step(X, None, w)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t) / 50
            print(f"  {mode}: {dt * 1e3:.3f} ms/step, launches of this library per step {(launch_count() - n0) / 50:.1f}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine auto|simt|tc]

Workload (config 2 of BASELINE.json): Flow.log_prob of the shipped 7-D galaxy flow
(redshift + ugrizy; 7 x [MLP 3->128->128->188 + RQ spline on 4 dims], ShiftBounds B=4,
Uniform(7,5)) on 10,000,000 synthetic in-distribution rows PER GPU (weak scaling; rows shard
with no data-path collective).  One step = one pass of log_prob over the rank's rows.
`value` is device-resident throughput (CUDA events, max over ranks); `e2e` is the same call
through the Flow-facing API with pinned HOST inputs and a host result: Flow._log_prob ->
pzb_log_prob_host, the host-buffer C-ABI call, which pipelines chunked H2D copies, the kernel
and D2H copies (all inside the timed region).  The own arm touches nothing under oracle/ (inputs come from the flow's own
inverse map on the device); only the CPU-baseline legs execute the oracle.  Secondary numbers (posterior galaxies/s on the config-3 geometry, sample
rows/s) ride along under "extra".

--impl reference times the CPU restatement of the reference algorithm (oracle/nsf_oracle.py,
NumPy float32 on all host cores) on a bounded sample of the same workload: the reference
itself needs jax, which is not installable here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 10_000_000
POSTERIOR_GALAXIES = 65_536   # bounded slice of config 3 (1e6 galaxies x 1000-point grid) per GPU
POSTERIOR_GRID = 1000
METRIC = "log_prob rows/s, 7-D galaxy flow (config 2: shipped example flow, 10M rows per GPU)"


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "sm_max_mhz": 1965.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, f"/tmp/pzb_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    out["sm_max_mhz"] = float(f[2])
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cpu_port_rows_per_s(sample_rows, seed=123):
    """The oracle (NumPy restatement of the reference) on a bounded sample of the workload."""
    with _all_host_threads():
        return _cpu_port_rows_per_s(sample_rows, seed)


def _cpu_port_rows_per_s(sample_rows, seed=123):
    from oracle import fixtures as F
    from oracle import nsf_oracle as O
    ex = F.example_flow()
    X = F.galaxy_rows(4096, seed=seed)
    X = np.ascontiguousarray(np.tile(X, (sample_rows // 4096 + 1, 1))[:sample_rows])
    O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], X[:4096])  # warm BLAS
    t0 = time.perf_counter()
    done = 0
    while done < sample_rows:  # the reference materialises [N, L, 3K-1] intermediates: chunk rows
        chunk = X[done:done + 200_000]
        O.flow_log_prob(ex["bijector_info"], ex["latent_info"], ex["params"], chunk)
        done += len(chunk)
    dt = time.perf_counter() - t0
    return sample_rows / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 100_000
    for _ in range(args.warmup):
        cpu_port_rows_per_s(20_000)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_port_rows_per_s(sample)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "7-D galaxy flow log_prob (config 2), bounded CPU sample",
                   "rows_per_step": sample, "params": "shipped example flow (tests/golden/example_flow_params.npz)"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} rows per step x {args.steps} steps, NumPy float32 oracle "
                                   "(restated reference algorithm; jax is not installable here)"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# The other named shapes of BASELINE.json (configs 1, 3, 4, 5; synthetic inputs as SURVEY.md section 8d states them).
# Each returns {"value", "unit", "e2e": {...}, "roofline": {...}, sizes}; C3 and C4 are STRONG-scaled by default
# (BASELINE's totals -- 1e6 galaxies, 1e8 rows -- split over the ranks), `--scaling weak` gives every rank the
# single-GPU size instead.
# ---------------------------------------------------------------------------------------------
C3_GALAXIES, C3_GRID = 1_000_000, 1000
C4_ROWS, C4_ROWS_WEAK, C4_FLOWS = 100_000_000, 10_000_000, 4
C5_ROWS, C1_ROWS = 1_000_000, 100_000
GALAXY_COLUMNS = ("redshift", "u", "g", "r", "i", "z", "y")


def timed_device(fn, reps, barrier, max_ranks, warm=1):
    """ms per call, CUDA events on torch's current stream (the stream every Plan call launches on), max over ranks."""
    import torch
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return max_ranks(e0.elapsed_time(e1)) / reps


def timed_host(fn, reps, barrier, max_ranks, warm=1):
    """ms per blocking host-buffer call (host arrays in, host result out), wall clock, max over ranks."""
    for _ in range(warm):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    return max_ranks((time.perf_counter() - t0) * 1e3) / reps


def pipe_roofline(engine, flops, ms, peaks):
    """Algorithmic conditioner FLOPs of ONE GPU's share over its time, against the pipe the engine runs on."""
    achieved = flops / (ms * 1e-3) / 1e12
    if engine in ("tc", "wide"):
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "executed_frac": 3.0 * achieved / peak, "engine": engine,
                "note": "algorithmic fp32 FLOPs over the measured sustained cuBLAS bf16 rate; the fp16x2 split "
                        "executes 3 MMAs per fp32 MAC (executed_frac), so frac cannot exceed 1/3; the sustained peak is "
                        "itself power-capped (cuBLAS: ~0.9 pipe-active x ~1.26 GHz), and these kernels sit at the same "
                        "pipe-active x clock product (profiles/README.md)"}
    peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    return {"bound": "fp32_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "engine": engine, "note": "nominal 148 SM x 128 FFMA/clk x 2 x clocks.max.sm"}


def posterior_traffic(ng):
    """DRAM bytes per galaxy of the single-pass posterior from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)["nsf_chain_tc_kernel:posterior_fused"]
        per = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["galaxies_per_launch"]
        return {"traffic_bytes_per_galaxy": per, "traffic_over_algorithmic": per / (4 * 6 + 4 * t["grid_points"]),
                "traffic_source": t["source"]}
    except (OSError, KeyError, ValueError):
        return {"traffic_bytes_per_galaxy": None}


def share(total, rank, world):
    base, rem = divmod(total, world)
    return base + (1 if rank < rem else 0)


def bench_c3(flow, plan, X, ctx):
    """Config 3: Flow.posterior over a 1000-point redshift grid for 1e6 galaxies (pzflow/flow.py:666-738)."""
    import torch
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    ng = share(C3_GALAXIES, rank, world) if ctx["strong"] else C3_GALAXIES
    total = C3_GALAXIES if ctx["strong"] else C3_GALAXIES * world
    rest = X[:ng, 1:].contiguous()
    grid = torch.linspace(0.0, 2.3, C3_GRID, device=dev)
    pdfs = torch.empty((ng, C3_GRID), dtype=torch.float32, device=dev)
    ms = timed_device(lambda: plan.posterior(rest, 0, grid, out=pdfs), 2, ctx["barrier"], ctx["max_ranks"])
    # the single-pass form (exp, trapezoid, divide, NaN -> 0 in shared memory before the only write of the pdfs)
    os.environ["PZB_FUSE_POSTERIOR"] = "1"
    try:
        fused_ms = timed_device(lambda: plan.posterior(rest, 0, grid, out=pdfs), 2, ctx["barrier"], ctx["max_ranks"])
    finally:
        os.environ.pop("PZB_FUSE_POSTERIOR", None)
    del pdfs
    torch.cuda.empty_cache()
    # end to end: photometry in host memory, pdfs [Ng, 1000] (4 GB at Ng = 1e6) into an ordinary PAGEABLE array
    rest_h = rest.cpu()
    grid_h = grid.cpu()
    out_h = torch.empty((ng, C3_GRID), dtype=torch.float32)
    e2e_ms = timed_host(lambda: plan.posterior_host(rest_h, 0, grid_h, out=out_h), 2, ctx["barrier"], ctx["max_ranks"])
    ok = bool(torch.isfinite(out_h[:: max(1, ng // 1000)]).all())
    del out_h
    return {"workload": "Flow.posterior, 7-D galaxy flow, redshift grid of 1000 points (BASELINE config 3)",
            "value": total / (ms * 1e-3), "unit": "galaxies/s", "ms": ms, "galaxies_total": total,
            "galaxies_this_rank": ng, "grid_points": C3_GRID, "scaling": "strong" if ctx["strong"] else "weak",
            "rows_per_s_equiv": total * C3_GRID / (ms * 1e-3),
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "galaxies/s", "ms": e2e_ms,
                    "h2d_bytes": int(ng * 6 * 4), "d2h_bytes": int(ng * C3_GRID * 4),
                    "result": "pageable host array through pzb_posterior_host", "finite": ok},
            "roofline": pipe_roofline(plan.engine, plan.flops_per_row * ng * C3_GRID, ms, ctx["peaks"]),
            "algorithmic_bytes_per_galaxy": 4 * 6 + 4 * C3_GRID,
            "single_pass": dict({"value": total / (fused_ms * 1e-3), "unit": "galaxies/s", "ms": fused_ms,
                                 "note": "PZB_FUSE_POSTERIOR=1: normalised inside the chain kernel, the pdfs cross HBM once; "
                                         "opt-in because a 1000-point galaxy fills 7.8 row tiles (24 idle rows) and both epilogue "
                                         "groups meet once per galaxy; the default is the chain kernel + normalize_pdfs_kernel "
                                         "(12 B per grid point instead of 4, 0.1 % of the time at 6.5 TB/s)"},
                                **posterior_traffic(ng))}


def bench_c4(flow, ctx):
    """Config 4: FlowEnsemble of 4 conditional 7-D flows (conditional_columns ra, dec), log_prob on 1e8 rows
    (pzflow/flowEnsemble.py:189-243)."""
    import pandas as pd
    import torch
    from pzflow_b200 import FlowEnsemble
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    from pzflow_b200.plan import ensemble_logmeanexp
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    n = share(C4_ROWS, rank, world) if ctx["strong"] else C4_ROWS_WEAK
    total = C4_ROWS if ctx["strong"] else C4_ROWS_WEAK * world
    mins = np.asarray(flow._bijector_info[1][0][1][0], np.float32)
    maxs = np.asarray(flow._bijector_info[1][0][1][1], np.float32)
    ens = FlowEnsemble(data_columns=GALAXY_COLUMNS, conditional_columns=("ra", "dec"), N=C4_FLOWS,
                       bijector=Chain(ShiftBounds(mins, maxs, 4.0), RollingSplineCoupling(7, n_conditions=2)))
    flows = list(ens._ensemble.values())
    plans = [f._plan() for f in flows]
    gen = ctx["gen"]
    Xc = torch.rand((n, 7), generator=gen, device=dev)
    Xc.mul_(torch.as_tensor(maxs - mins, device=dev)).add_(torch.as_tensor(mins, device=dev))
    Cc = torch.randn((n, 2), generator=gen, device=dev)
    lps = torch.empty((C4_FLOWS, n), dtype=torch.float32, device=dev)

    def step():
        for f, p in enumerate(plans):
            p.log_prob(Xc, Cc, out=lps[f])
        return ensemble_logmeanexp(lps)
    ms = timed_device(step, 2, ctx["barrier"], ctx["max_ranks"])
    del lps
    # end to end through the public API: FlowEnsemble.log_prob(DataFrame of float32 columns) -> NumPy array
    frame = pd.DataFrame({c: Xc[:, j].cpu().numpy() for j, c in enumerate(GALAXY_COLUMNS)}, copy=False)
    frame["ra"], frame["dec"] = Cc[:, 0].cpu().numpy(), Cc[:, 1].cpu().numpy()
    del Xc, Cc
    torch.cuda.empty_cache()
    e2e_ms = timed_host(lambda: ens.log_prob(frame), 2, ctx["barrier"], ctx["max_ranks"])
    flops = sum(p.flops_per_row for p in plans) * n
    return {"workload": "FlowEnsemble(4) of conditional 7-D flows, log_prob (BASELINE config 4)",
            "value": total / (ms * 1e-3), "unit": "rows/s", "ms": ms, "rows_total": total, "rows_this_rank": n,
            "flows": C4_FLOWS, "flow_rows_per_s": C4_FLOWS * total / (ms * 1e-3),
            "scaling": "strong" if ctx["strong"] else "weak",
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "rows/s", "ms": e2e_ms, "h2d_bytes": int(n * 9 * 4),
                    "d2h_bytes": int(n * 4), "call": "FlowEnsemble.log_prob(DataFrame[float32 x 9])"},
            "roofline": pipe_roofline(plans[0].engine, flops, ms, ctx["peaks"]),
            "params": "synthetic (package initialiser, seeds 0-3)"}


def bench_c5(ctx, engine_arg):
    """Config 5: 32-D RollingSplineCoupling, K = 32, hidden 256, log_prob + inverse on 1e6 rows."""
    import torch
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    n = C5_ROWS
    cols = tuple(f"x{j}" for j in range(32))
    f5 = Flow(data_columns=cols, bijector=Chain(ShiftBounds(-1.0, 1.0, 4.0),
                                                RollingSplineCoupling(nlayers=32, K=32, hidden_dim=256)))
    plan = f5._plan()
    if engine_arg == "simt":
        plan.set_engine("simt")
    X5 = torch.rand((n, 32), generator=ctx["gen"], device=dev) * 2.0 - 1.0
    out = torch.empty(n, dtype=torch.float32, device=dev)
    ms = timed_device(lambda: plan.log_prob(X5, out=out), 2, ctx["barrier"], ctx["max_ranks"])
    u5 = plan.forward(X5)[0]
    inv_ms = timed_device(lambda: plan.inverse(u5, want_log_det=False), 2, ctx["barrier"], ctx["max_ranks"])
    back = plan.inverse(u5, want_log_det=False)[0]
    round_trip = float((back - X5).abs().median().item())
    X5_h = X5.cpu().pin_memory()
    del X5, u5, back
    e2e_ms = timed_host(lambda: plan.log_prob_host(X5_h), 2, ctx["barrier"], ctx["max_ranks"])
    return {"workload": "32-D RollingSplineCoupling(32 layers, K=32, hidden 256) log_prob + inverse, 1e6 rows per GPU "
                        "(BASELINE config 5)",
            "value": n * world / (ms * 1e-3), "unit": "rows/s", "ms": ms, "rows_per_gpu": n, "scaling": "weak",
            "inverse": {"value": n * world / (inv_ms * 1e-3), "unit": "rows/s", "ms": inv_ms,
                        "round_trip_median_abs_err": round_trip},
            "e2e": {"value": n * world / (e2e_ms * 1e-3), "unit": "rows/s", "ms": e2e_ms, "h2d_bytes": int(n * 32 * 4),
                    "d2h_bytes": int(n * 4)},
            "roofline": pipe_roofline(plan.engine, plan.flops_per_row * n, ms, ctx["peaks"]),
            "inverse_roofline": pipe_roofline(plan.engine, plan.flops_per_row * n, inv_ms, ctx["peaks"]),
            "flops_per_row": plan.flops_per_row, "params": "synthetic (package initialiser, seed 0)"}


def bench_one_process(ctx):
    """Flow.log_prob(DataFrame) of a plain process on a multi-GPU box: the rows are sharded over every visible GPU
    inside the call (pzflow_b200.sharding), no torchrun."""
    import pandas as pd
    import torch
    from pzflow_b200 import sharding
    from pzflow_b200.examples import get_example_flow
    flow = get_example_flow()
    n = ROWS_PER_GPU * torch.cuda.device_count()
    plan = flow._plan()
    U = torch.rand((n, 7), generator=ctx["gen"], device=ctx["dev"]) * 10.0 - 5.0
    X = torch.nan_to_num(plan.inverse(U, want_log_det=False)[0]).cpu().numpy()
    del U
    frame = pd.DataFrame({c: np.ascontiguousarray(X[:, j]) for j, c in enumerate(GALAXY_COLUMNS)}, copy=False)
    out = {}
    for tag, devs in (("one_gpu", [0]), ("all_gpus", "all")):
        sharding.set_devices(devs)
        flow.log_prob(frame)
        t0 = time.perf_counter()
        for _ in range(3):
            lp = flow.log_prob(frame)
        out[tag] = {"value": n / ((time.perf_counter() - t0) / 3), "unit": "rows/s"}
    sharding.set_devices(None)
    out.update({"workload": "Flow.log_prob(DataFrame[float32 x 7]) end to end in ONE process", "rows": n,
                "gpus": torch.cuda.device_count(), "finite": bool(np.isfinite(lp).mean() > 0.9)})
    return out


def bench_c1(ctx):
    """Config 1: the default 2-D flow (two-moons shape: ShiftBounds + RollingSplineCoupling(2), CentBeta13 latent),
    log_prob of 100k rows."""
    import pandas as pd
    import torch
    from pzflow_b200 import Flow
    from pzflow_b200.bijectors import Chain, RollingSplineCoupling, ShiftBounds
    dev, world = ctx["dev"], ctx["world"]
    n = C1_ROWS
    mins, maxs = np.float32([-1.5, -1.0]), np.float32([2.5, 1.5])
    f1 = Flow(data_columns=("x", "y"), bijector=Chain(ShiftBounds(mins, maxs, 4.0), RollingSplineCoupling(2)))
    plan = f1._plan()
    X1 = torch.rand((n, 2), generator=ctx["gen"], device=dev)
    X1.mul_(torch.as_tensor(maxs - mins, device=dev)).add_(torch.as_tensor(mins, device=dev))
    out = torch.empty(n, dtype=torch.float32, device=dev)
    ms = timed_device(lambda: plan.log_prob(X1, out=out), 50, ctx["barrier"], ctx["max_ranks"], warm=3)
    frame = pd.DataFrame({"x": X1[:, 0].double().cpu().numpy(), "y": X1[:, 1].double().cpu().numpy()})
    e2e_ms = timed_host(lambda: f1.log_prob(frame), 20, ctx["barrier"], ctx["max_ranks"], warm=3)
    return {"workload": "default 2-D flow (two-moons shape, K=16, CentBeta13 latent) log_prob, 100k rows per GPU "
                        "(BASELINE config 1)",
            "value": n * world / (ms * 1e-3), "unit": "rows/s", "ms": ms, "rows_per_gpu": n, "scaling": "weak",
            "e2e": {"value": n * world / (e2e_ms * 1e-3), "unit": "rows/s", "ms": e2e_ms,
                    "call": "Flow.log_prob(DataFrame[float64 x 2])", "h2d_bytes": int(n * 2 * 4), "d2h_bytes": int(n * 4)},
            "roofline": pipe_roofline(plan.engine, plan.flops_per_row * n, ms, ctx["peaks"]),
            "note": "one launch of 100k rows is 782 row tiles over 148 SMs: launch- and tail-bound, not pipe-bound"}


def bench_train_step(flow, X, ctx):
    """one Flow.train optimisation step (gradient of -mean(log_prob) + Adam, pzflow/flow.py:961-982) of the
    shipped flow on a batch of 1024 rows, as the fused kernel runs it (csrc/nsf_train_step.cu); N > 1: per-rank batches of
    1024 with the NCCL all-reduce of the flat gradient inside the step."""
    import torch
    from pzflow_b200.plan import launch_count
    from pzflow_b200.train import FusedGradient, NativeAdam, TorchChain, dp_step
    rows = 1024
    chain = TorchChain(flow._bijector_info, flow._params[1], flow._latent_info, flow._params[0], flow._input_dim, ctx["dev"])
    chain.fused_gradient = FusedGradient.build(chain)
    if chain.fused_gradient is None:
        return {"unavailable": "chain outside the fused training kernel's shapes"}
    opt = NativeAdam(chain, lr=1e-6)
    xb = X[:rows].contiguous()
    w = torch.ones(rows, device=ctx["dev"])
    n0 = launch_count()
    ms = timed_device(lambda: dp_step(chain, opt, xb, None, w, rows * ctx["world"]), 50, ctx["barrier"], ctx["max_ranks"], warm=5)
    per_step = (launch_count() - n0) / 55.0
    return {"workload": "Flow.train step (gradient + Adam) of the 7-D flow, batch 1024 per GPU (SURVEY f-4)",
            "value": ms, "unit": "ms/step", "higher_is_better": False, "batch_per_gpu": rows,
            "rows_per_s": rows * ctx["world"] / (ms * 1e-3), "launches_per_step": per_step,
            "collective": "NCCL all_reduce(sum) of the flat gradient" if ctx["world"] > 1 else "none (one GPU)"}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pzflow_b200.examples import get_example_flow
    from pzflow_b200.plan import launch_count

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pzflow_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1 and os.environ.get("PZB_NUMA_BIND", "1") != "0":
        # one process per GPU: keep this rank's threads and page-locked staging buffers on the GPU's NUMA node
        from pzflow_b200.sharding import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # the contract is ONE JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    flow = get_example_flow()   # the reference's shipped 7-D galaxy flow (pzflow/examples.py:105-113)
    plan = flow._plan()
    plan.set_engine(args.engine)
    engine = plan.engine

    # synthetic in-distribution rows: x = inverse(u), u ~ U(-5,5)^7 -- the flow's own sampling
    # map, run on the device -- plus a 1 % admixture of rows from the ShiftBounds box widened by 20 %
    # (exercises the out-of-bounds paths).  Nothing under oracle/ is touched on this arm.
    n = share(args.rows, rank, world) if args.scaling == "strong" else args.rows
    total_rows = args.rows if args.scaling == "strong" else args.rows * world
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)
    U = (torch.rand((n, 7), generator=gen, device=dev) * 10.0 - 5.0)
    X = plan.inverse(U, want_log_det=False)[0]
    del U
    mins = torch.as_tensor(np.asarray(flow._bijector_info[1][0][1][0], np.float32), device=dev)
    maxs = torch.as_tensor(np.asarray(flow._bijector_info[1][0][1][1], np.float32), device=dev)
    n_oob = n // 100
    rows = torch.randperm(n, generator=gen, device=dev)[:n_oob]
    span = maxs - mins
    X[rows] = (mins - 0.2 * span) + torch.rand((n_oob, 7), generator=gen, device=dev) * (1.4 * span)
    X = torch.nan_to_num(X, nan=0.0).contiguous()
    out = torch.empty(n, dtype=torch.float32, device=dev)
    X_host = X.cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        return flow._log_prob(flow._params, X)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = launch_count() - l0
    total_ms = max_ranks(ev[0].elapsed_time(ev[-1]))
    kernel_ms = statistics.mean(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    ms_per_step = total_ms / args.steps
    value = total_rows / (ms_per_step * 1e-3)

    # ---- end to end through the Flow-facing API with pinned host buffers ----
    def e2e_step():
        # pinned HOST rows in, HOST result out, through the host-buffer C-ABI call
        # (pzb_log_prob_host): chunked H2D | kernel | D2H pipeline, blocking
        return flow._log_prob(flow._params, X_host)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = max_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = total_rows / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and e2e)

    # ---- sample (inverse map) on the headline rows ----
    u = flow.latent.sample_device(n, seed=rank)
    samp_ms = timed_device(lambda: plan.inverse(u, want_log_det=False), 2, barrier, max_ranks)
    del u

    # ---- the other named shapes of BASELINE.json (configs 3, 4, 5, 1), each device-resident and end to end ----
    peaks = measured_peaks()
    ctx = dict(dev=dev, rank=rank, world=world, barrier=barrier, max_ranks=max_ranks, peaks=peaks,
               strong=args.scaling != "weak", gen=gen)
    named = {}
    if not args.no_named_shapes:
        named["train_step"] = bench_train_step(flow, X, ctx)
        named["c3_posterior"] = bench_c3(flow, plan, X, ctx)
        del X, out, X_host
        torch.cuda.empty_cache()
        named["c4_ensemble"] = bench_c4(flow, ctx)
        torch.cuda.empty_cache()
        named["c5_high_dim"] = bench_c5(ctx, args.engine)
        torch.cuda.empty_cache()
        named["c1_two_moons"] = bench_c1(ctx)
        if world == 1 and torch.cuda.device_count() > 1:
            named["one_process_all_gpus"] = bench_one_process(ctx)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    flops_row = plan.flops_per_row
    achieved = flops_row * n / (kernel_ms * 1e-3) / 1e12  # algorithmic conditioner TFLOP/s of one GPU
    if engine in ("tc", "wide"):
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None, "peak_source": f"{peaks['_source']} cuBLAS bf16 sustained (MEASURED_PEAKS.json)",
                "note": "algorithmic fp32 FLOPs; the fp16x2 split executes 3 tensor passes per GEMM, "
                        "so the tensor pipe does 3x this",
                # what the tensor pipe actually executes (3 fp16 MMAs per algorithmic fp32 MAC), against the same peak
                "executed": 3.0 * achieved, "executed_frac": 3.0 * achieved / peak}
    else:
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        roof = {"bound": "fp32_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None, "peak_source": "nominal 148 SM x 128 FFMA/clk x 2 x clocks.max.sm "
                                                "(the FP32 pipe has no entry in MEASURED_PEAKS.json)"}
    roof["kernel"] = {"tc": "nsf_chain_tc_kernel", "wide": "nsf_chain_wide_kernel"}.get(engine, "nsf_chain_simt_kernel")
    # DRAM traffic of one launch from the committed ncu --set full capture, scaled by rows (the kernel reads each
    # input row and writes each result once, so traffic is linear in rows); algorithmic bytes: 4 * (D + 1) per row
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)[roof["kernel"]]
        roof["traffic"] = (t["dram_bytes_read"] + t["dram_bytes_write"]) * n / t["rows_per_launch"]
        roof["traffic_source"] = t["source"]
    except (OSError, KeyError, ValueError):
        roof["traffic"] = None
    roof["algorithmic_bytes"] = 4.0 * (7 + 1) * n
    roof["flops_per_row"] = flops_row
    roof["kernel_ms"] = kernel_ms

    cpu_v, cpu_dt = (None, None)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = 200_000
        cpu_v, cpu_dt = cpu_port_rows_per_s(sample)
        cpu = {"value": cpu_v, "unit": "rows/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{sample} rows of the same workload in {cpu_dt:.1f} s, NumPy float32 oracle "
                         "(restated reference algorithm, multithreaded BLAS; jax not installable)"}

    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.scaling == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32" if engine == "simt" else "f32 (fp16x2-split tensor-core conditioner)",
        "data": "synthetic",
        "config": {"workload": "7-D galaxy flow log_prob, 10M synthetic rows per GPU (BASELINE config 2)",
                   "rows_per_gpu": n, "params": "shipped example flow (7 x MLP 3-128-128-188, K=16)",
                   "engine": engine, "cache": "inputs_larger_than_L2 (280 MB per GPU per step)",
                   "parallelism": f"row-sharded x{world}, no collective",
                   "numa_bound": numa_node is not None},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": int(n * 7 * 4),
                "d2h_bytes_per_step": int(n * 4), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "extra": dict({"sample": {"value": total_rows / (samp_ms * 1e-3), "unit": "rows/s", "ms": samp_ms,
                                  "workload": "Flow.sample map (inverse chain) on the headline rows, device-resident"}},
                      **named),
    }
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc", "wide"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-named-shapes", action="store_true", help="skip the config 1/3/4/5 lines under `extra`")
    ap.add_argument("--scaling", default="default", choices=["default", "weak", "strong"],
                    help="default: headline weak (10M rows per GPU), configs 3/4 strong (BASELINE totals split over the "
                         "ranks); weak: every rank gets the single-GPU size; strong: the headline's 10M rows are split too")
    args = ap.parse_args()
    rank, world, local_rank = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # each step is ~4 s of CPU work; keep the run within minutes
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched through torchrun: re-launch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to fd 1
    # on the first collective), so fd 1 points at stderr while the benchmark runs and is restored for the result.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_ours(args, rank, world, local_rank)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

"""Row sharding across the GPUs of one box (one process per GPU).

Every row -- and every (galaxy, grid point) pair of a posterior -- is independent
(pzflow/bijectors.py:155-160 is row-wise throughout; pzflow/flow.py:734 normalises per
galaxy), so the path shards into contiguous row blocks with NO data-path collective:
weights are replicated per rank at plan creation.  The only collectives here are the
optional result gather and the max-over-ranks used for timing.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank ``rank``; sizes differ by at most one row."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sharded_eval(fn: Callable[[torch.Tensor], torch.Tensor], rows: torch.Tensor, gather: bool = False):
    """Evaluate ``fn`` on this rank's row block; optionally all-gather the [N, ...] result."""
    rank, ws = world()
    a, b = shard_rows(rows.shape[0], rank, ws)
    local = fn(rows[a:b])
    if not gather or ws == 1:
        return local
    sizes = [shard_rows(rows.shape[0], r, ws) for r in range(ws)]
    pad = max(e - s for s, e in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf)
    return torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """max of a scalar over all ranks (timing is reported as the slowest rank)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (and every thread it starts later, e.g. the OpenMP cast team of the host pipeline) to the
    CPUs of the NUMA node the GPU hangs off, so that the page-locked staging buffers it allocates afterwards are
    node-local and the chunk copies do not cross the socket interconnect.  One process per GPU calls this once,
    before the first host-buffer call.  Returns the node, or None when the topology cannot be read (single-node
    hosts, containers without sysfs) -- in which case nothing is changed."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


# ---------------------------------------------------------------------------------------------
# One process, several GPUs: what a pzflow user gets without torchrun.  Flow.log_prob / posterior / sample cut their
# host arrays into contiguous row blocks (shard_rows), one per device, and run the blocks' host pipelines
# concurrently from one host thread each (the C calls release the GIL).  No collective: every block lands in its own
# slice of the caller's result.  Rows are independent, so the result equals the one-GPU result bit for bit.
# ---------------------------------------------------------------------------------------------
MULTI_DEVICE_MIN_ROWS = 1 << 21   # below this one GPU is faster than waking the others (plans, streams, staging)

_devices_override = None


def set_devices(devices) -> None:
    """Devices the in-process sharding may use: ``"all"``, a list of CUDA device indices, or None for the default
    (environment variable PZFLOW_B200_DEVICES = "all" | "0,1,..." | "current"; unset: every visible device unless this
    process is one rank of a multi-process launch, which owns exactly its current device)."""
    global _devices_override
    _devices_override = devices


def process_devices() -> list:
    """CUDA device indices this process shards host-array calls over (always at least the current device)."""
    if not torch.cuda.is_available():
        return []
    cur = torch.cuda.current_device()
    spec = _devices_override if _devices_override is not None else os.environ.get("PZFLOW_B200_DEVICES")
    n = torch.cuda.device_count()
    if spec is None:
        launched = int(os.environ.get("WORLD_SIZE", "1") or 1) > 1 or (dist.is_available() and dist.is_initialized())
        return [cur] if launched else list(range(n))
    if isinstance(spec, str):
        spec = spec.strip().lower()
        if spec == "all":
            return list(range(n))
        if spec in ("current", ""):
            return [cur]
        spec = [int(x) for x in spec.split(",") if x.strip()]
    out = [int(d) for d in spec if 0 <= int(d) < n]
    return out or [cur]


def shard_bounds(n: int, parts: int, multiple: int = 4) -> list:
    """[(start, stop)] of ``parts`` contiguous blocks whose starts are multiples of ``multiple`` rows (the latent
    samplers draw in blocks of four values); empty blocks are dropped."""
    out = []
    for r in range(parts):
        a, b = shard_rows(n, r, parts)
        a = min(n, (a + multiple - 1) // multiple * multiple) if r else 0
        b = min(n, (b + multiple - 1) // multiple * multiple) if r + 1 < parts else n
        if b > a:
            out.append((a, b))
    return out


def run_sharded(n_rows: int, devices: list, fn) -> None:
    """fn(device_index, start, stop) for every device's row block, concurrently, one host thread per device; each
    thread gets its share of the host cores for the pipeline's casts and staging copies (pzb_set_host_threads)."""
    from . import _native as N
    blocks = shard_bounds(n_rows, len(devices))
    if len(blocks) <= 1:
        for (a, b) in blocks:
            fn(devices[0], a, b)
        return
    import threading
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    share = max(1, min(16, cores // len(blocks)))
    errors = []

    def work(dev, a, b):
        try:
            N.check(N.lib().pzb_set_host_threads(share))
            fn(dev, a, b)
        except BaseException as exc:   # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(devices[i], a, b), daemon=True) for i, (a, b) in enumerate(blocks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]

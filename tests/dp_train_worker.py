"""Worker of test_gpu_parity.test_data_parallel_training_two_gpus: launched by torchrun, one rank per GPU.
Trains the default 2-D flow for a few epochs with the graphed data-parallel step and reports, per rank, the losses
and a checksum of the parameters (they must be identical across ranks)."""
import json
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    from pzflow_b200 import Flow
    rng = np.random.default_rng(0)
    n = 8192 + 300                                     # a ragged last batch: the eager path runs beside the graph
    t = rng.uniform(0, np.pi, n)
    up = rng.random(n) < 0.5
    data = pd.DataFrame({"x": np.where(up, np.cos(t), 1 - np.cos(t)) + rng.normal(0, 0.06, n),
                         "y": np.where(up, np.sin(t), 0.5 - np.sin(t)) + rng.normal(0, 0.06, n)})
    flow = Flow(data_columns=("x", "y"))
    losses = flow.train(data, epochs=4, seed=0)
    leaves = []

    def walk(p):
        if isinstance(p, (list, tuple)):
            for q in p:
                walk(q)
        else:
            leaves.append(np.asarray(p, np.float32).ravel())
    walk(flow._params[1])
    flat = np.concatenate(leaves)
    gathered = [None] * world
    dist.all_gather_object(gathered, (losses, flat.tobytes()))
    if rank == 0:
        same = all(g[1] == gathered[0][1] for g in gathered)
        print("DPRESULT " + json.dumps({"world": world, "losses": gathered[0][0], "identical_params": same,
                                         "identical_losses": all(g[0] == gathered[0][0] for g in gathered)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

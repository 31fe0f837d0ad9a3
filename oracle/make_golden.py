"""Regenerate tests/golden/ from /root/reference (run in the build container only).

    python -m oracle.make_golden

1. example_flow_params.npz -- the shipped flow's parameters as plain arrays
   (pzflow/example_files/example-flow.pzflow.pkl is a dill dict of NumPy float32
   arrays; unpickling needs only a stub for pzflow.utils.gaussian_error_model).
2. ref_shim_*.npz -- outputs of the reference's OWN code (pzflow/*.py imported
   from /root/reference, unmodified) executed over oracle/jax_shim, the
   NumPy-backed stand-in for the jax API.  See oracle/jax_shim/__init__.py for
   what that does and does not pin.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def export_example_flow():
    import dill
    stub = types.ModuleType("pzflow")
    stub_utils = types.ModuleType("pzflow.utils")
    stub_utils.gaussian_error_model = lambda *a, **k: None
    stub.utils = stub_utils
    saved = {k: sys.modules.get(k) for k in ("pzflow", "pzflow.utils")}
    sys.modules["pzflow"], sys.modules["pzflow.utils"] = stub, stub_utils
    try:
        with open(os.path.join(REF, "pzflow", "example_files", "example-flow.pzflow.pkl"), "rb") as fh:
            d = dill.load(fh)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    info = d["bijector_info"]
    assert info[0] == "Chain" and info[1][0][0] == "ShiftBounds" and d["latent_info"] == ("Uniform", (7, 5))
    mins, maxs, sbB = info[1][0][1]
    inner = d["params"][1][1]
    out = dict(mins=np.asarray(mins, np.float32), maxs=np.asarray(maxs, np.float32), sb_B=np.int32(sbB),
               n_layers=np.int32(len(inner) // 2))
    for i in range(len(inner) // 2):
        dense = [p for p in inner[2 * i] if len(p)]
        for j, (W, b) in enumerate(dense):
            out[f"l{i}_W{j}"] = np.asarray(W, np.float32)
            out[f"l{i}_b{j}"] = np.asarray(b, np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "example_flow_params.npz"), **out)
    print("wrote example_flow_params.npz")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    export_example_flow()
    try:
        from oracle import make_golden_shim
    except ImportError:
        make_golden_shim = None
    if make_golden_shim is not None:
        make_golden_shim.main()

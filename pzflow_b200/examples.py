"""``get_example_flow`` (pzflow/examples.py:105-113): the reference's shipped 7-D galaxy flow
(redshift + LSST ugrizy; Chain(ShiftBounds(mins, maxs, 4), RollingSplineCoupling(7)), Uniform(7, 5)).

The reference loads ``example_files/example-flow.pzflow.pkl`` (a dill pickle holding jax arrays); the same
parameters ship here as plain float32 arrays in ``example_files/example-flow-params.npz`` (written by
``oracle/make_golden.py`` from that pickle), so no jax is needed.  The example DATA sets of the reference
(two moons, galaxies, checkerboard, cities) are not redistributed.
"""
from __future__ import annotations

import os

import numpy as np

from .flow import Flow

EXAMPLE_FILE_DIR = "example_files"
GALAXY_COLUMNS = ("redshift", "u", "g", "r", "i", "z", "y")


def example_flow_dict() -> dict:
    """The save-dict (pzflow/flow.py:819-845) of the shipped flow."""
    this_dir, _ = os.path.split(__file__)
    z = np.load(os.path.join(this_dir, EXAMPLE_FILE_DIR, "example-flow-params.npz"))
    nl = int(z["n_layers"])
    nsc = ("NeuralSplineCoupling", (16, 5, 2, 128, None, 0, False))
    info = ("Chain", (("ShiftBounds", (z["mins"], z["maxs"], int(z["sb_B"]))),
                      ("Chain", (nsc, ("Roll", (1,))) * nl)))
    inner = []
    for i in range(nl):
        layer = []
        for j in range(3):
            layer.append((z[f"l{i}_W{j}"], z[f"l{i}_b{j}"]))
            if j < 2:
                layer.append(())
        inner += [layer, ()]
    return {"class": "Flow", "data_columns": GALAXY_COLUMNS, "conditional_columns": None,
            "condition_means": None, "condition_stds": None, "data_error_model": None,
            "condition_error_model": None, "autoscale_conditions": True, "info": "pzflow example flow",
            "latent_info": ("Uniform", (7, 5)), "bijector_info": info, "params": ((), [(), inner])}


def get_example_flow() -> Flow:
    """Return a normalizing flow that was trained on galaxy data (pzflow/examples.py:105-113)."""
    return Flow(_dictionary=example_flow_dict())

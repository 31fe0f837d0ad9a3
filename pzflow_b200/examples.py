"""``get_example_flow`` (pzflow/examples.py:105-113): the reference's shipped 7-D galaxy flow
(redshift + LSST ugrizy; Chain(ShiftBounds(mins, maxs, 4), RollingSplineCoupling(7)), Uniform(7, 5)).

The reference loads ``example_files/example-flow.pzflow.pkl`` (a dill pickle holding jax arrays); the same
parameters ship here as plain float32 arrays in ``example_files/example-flow-params.npz`` (written by
``oracle/make_golden.py`` from that pickle), so no jax is needed.  The example DATA sets of the reference
(two moons, galaxies, checkerboard, cities; pzflow/examples.py:13-102) are not redistributed: the loaders below read
them from ``$PZFLOW_EXAMPLE_FILES`` or from an installed pzflow's ``example_files`` directory.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import pandas as pd

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


def _load_example_data(name: str) -> pd.DataFrame:
    """pzflow/examples.py:13-17, looking where the reference's pickles can be: $PZFLOW_EXAMPLE_FILES, this package's
    own example_files directory, an installed pzflow's."""
    dirs = [os.environ.get("PZFLOW_EXAMPLE_FILES"), os.path.join(os.path.split(__file__)[0], EXAMPLE_FILE_DIR)]
    try:
        spec = importlib.util.find_spec("pzflow")
        if spec is not None and spec.submodule_search_locations:
            dirs.append(os.path.join(list(spec.submodule_search_locations)[0], EXAMPLE_FILE_DIR))
    except (ImportError, ValueError):
        pass
    for d in dirs:
        if d and os.path.exists(os.path.join(d, f"{name}.pkl")):
            return pd.read_pickle(os.path.join(d, f"{name}.pkl"))
    raise FileNotFoundError(f"{name}.pkl: the reference's example data sets are not redistributed with pzflow_b200; point "
                            "PZFLOW_EXAMPLE_FILES at pzflow's example_files directory")


def get_twomoons_data() -> pd.DataFrame:
    """Return DataFrame with two moons example data (pzflow/examples.py:20-26)."""
    return _load_example_data("two-moons-data")


def get_galaxy_data() -> pd.DataFrame:
    """Return DataFrame with example galaxy data (pzflow/examples.py:29-42)."""
    return _load_example_data("galaxy-data")


def get_checkerboard_data() -> pd.DataFrame:
    """Return DataFrame with discrete checkerboard data (pzflow/examples.py:45-47)."""
    return _load_example_data("checkerboard-data")


def get_city_data() -> pd.DataFrame:
    """Return DataFrame with example city data (pzflow/examples.py:50-102)."""
    return _load_example_data("city-data")

"""Pixel coordinates of the posterior curves the reference's authors plotted from real JAX (build container only).

docs/tutorials/marginalization.ipynb holds two figures made with the SHIPPED flow under real JAX/XLA:
  cell 10: ``flow.posterior(samples, column="redshift", grid=linspace(0.1, 2.4, 100))`` of the two samples of
           ``flow.sample(2, seed=0)`` (blue), one panel per sample, x-window true redshift +- 0.25;
  cell 15: the same (blue) and the posterior with ``u`` (both rows) and ``y`` (second row) flagged 99 and marginalised
           over linspace(27, 29, 40) / linspace(25, 27, 40) (orange).
The curves are the only outputs of the reference's log_prob / posterior path that real JAX produced and that survive in
the tree.  This script decodes the two PNGs and stores, per figure and panel, the axes box (pixel positions of the
spines) and the pixel coordinates of the blue and orange line -- numbers derived from the figures, not the figures --
in tests/golden/nb_posterior_plots.npz.  tests/plots.py predicts where a computed curve falls in such a panel
(matplotlib's default 5 % y-margins, the panel's fixed x-window) and measures its distance to those pixels: one pixel is
0.8 % of the tallest curve's peak.

    python -m oracle.make_golden_plots
"""
from __future__ import annotations

import base64
import io
import json
import os

import numpy as np

from . import fixtures as F

NOTEBOOK = "/root/reference/docs/tutorials/marginalization.ipynb"
COLORS = {"blue": (31, 119, 180), "orange": (255, 127, 14), "red": (214, 39, 40)}   # matplotlib's C0, C1, C3


def main():
    from PIL import Image
    cells = json.load(open(NOTEBOOK))["cells"]
    store = {}
    for cell in (10, 15):
        png = next(o["data"]["image/png"] for o in cells[cell]["outputs"] if "image/png" in o.get("data", {}))
        im = np.asarray(Image.open(io.BytesIO(base64.b64decode(png))).convert("RGB")).astype(int)
        black = im.sum(axis=2) < 60
        rows = np.nonzero(black.sum(axis=1) > 150)[0]          # the horizontal spines
        cols = np.nonzero(black.sum(axis=0) > 80)[0]           # the vertical spines: left / right of each panel
        assert len(rows) == 2 and len(cols) == 4, (rows, cols)
        store[f"c{cell}_shape"] = np.asarray(im.shape[:2])
        for p in range(2):
            L, R, T, B = int(cols[2 * p]), int(cols[2 * p + 1]), int(rows[0]), int(rows[1])
            store[f"c{cell}_p{p}_box"] = np.asarray([L, R, T, B])
            for name, rgb in COLORS.items():
                d = np.abs(im - np.asarray(rgb)).sum(axis=2)
                ys, xs = np.nonzero(d < 60)
                keep = (xs > L + 1) & (xs < R - 1) & (ys > T + 1) & (ys < B - 1)
                store[f"c{cell}_p{p}_{name}"] = np.stack([xs[keep], ys[keep]], axis=1).astype(np.int16)
    path = os.path.join(F.GOLDEN, "nb_posterior_plots.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in store.items():
        if k.endswith(("blue", "orange", "red")):
            print(" ", k, len(v), "pixels")


if __name__ == "__main__":
    main()

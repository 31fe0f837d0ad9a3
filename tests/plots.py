"""Where a computed posterior curve falls in a panel of the reference's marginalization notebook, and how far the pixels
the authors' real-JAX figure shows are from it (tests/golden/nb_posterior_plots.npz, oracle/make_golden_plots.py).

A panel plots ``grid`` against one or two curves with ``xlim = (ztrue - 0.25, ztrue + 0.25)`` and matplotlib's default
y-limits: the data range of ALL plotted curves over the WHOLE grid, widened by 5 % on both sides (the vertical line is
drawn in axes coordinates and does not take part).  So the pixel polyline of a curve follows from the curves alone --
no free parameter -- and one pixel is 1 / 128 (1 / 145) of the y-range, i.e. about 0.8 % of the tallest curve's peak.
"""
import numpy as np

NB_GRID = np.linspace(0.1, 2.4, 100).astype(np.float32)      # marginalization.ipynb cell 9
NB_U_GRID = np.linspace(27, 29, 40)                          # cell 14
NB_Y_GRID = np.linspace(25, 27, 40)


def panel_polyline(box, ztrue, curves, which):
    """Pixel coordinates [n, 2] of curve ``which`` of ``curves`` (all the curves of the panel) in a panel whose spines
    sit at box = (left, right, top, bottom)."""
    L, R, T, B = (float(v) for v in box)
    lo = min(float(np.min(c)) for c in curves)
    hi = max(float(np.max(c)) for c in curves)
    pad = 0.05 * (hi - lo)
    ylo, yhi = lo - pad, hi + pad
    x = L + (NB_GRID.astype(np.float64) - (ztrue - 0.25)) / 0.5 * (R - L)
    y = B - (np.asarray(curves[which], np.float64) - ylo) / (yhi - ylo) * (B - T)
    return np.stack([x, y], axis=1)


def distance_to_polyline(pixels, poly):
    """Distance of every pixel [m, 2] to the polyline [n, 2] (its nearest segment), in pixels."""
    p = np.asarray(pixels, np.float64)[:, None, :]
    a, b = poly[None, :-1, :], poly[None, 1:, :]
    ab = b - a
    t = np.clip(((p - a) * ab).sum(-1) / np.maximum((ab * ab).sum(-1), 1e-12), 0.0, 1.0)
    d = np.linalg.norm(p - (a + t[..., None] * ab), axis=-1)
    return d.min(axis=1)


def overlay_report(gold, cell, panel, ztrue, curves, which, color):
    """How well curve ``which`` overlays the ``color`` pixels of that panel: (pixels, median, 99 %, max distance) and the
    fraction of the curve's vertices inside the window that have a line pixel within two pixels."""
    box = gold[f"c{cell}_p{panel}_box"]
    px = gold[f"c{cell}_p{panel}_{color}"].astype(np.float64)
    poly = panel_polyline(box, ztrue, curves, which)
    d = distance_to_polyline(px, poly)
    inside = (poly[:, 0] > box[0] + 3) & (poly[:, 0] < box[1] - 3)
    hit = np.array([np.min(np.linalg.norm(px - v, axis=1)) <= 2.0 for v in poly[inside]]) if len(px) else np.zeros(0, bool)
    return dict(pixels=len(px), median=float(np.median(d)), p99=float(np.quantile(d, 0.99)), max=float(d.max()),
                vertices_hit=float(hit.mean()) if hit.size else 0.0)

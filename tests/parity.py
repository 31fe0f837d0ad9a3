"""Parity metrics shared by the GPU tests.

north_star's tolerance is 1e-4 absolute + 1e-5 relative in fp32.  On the shipped (trained) flow
two legitimate float32 evaluations already differ by more than that on ~40 % of rows: the
float32 oracle vs its own float64 twin has a MEDIAN |error| of 1.3e-4 (p99 9e-4), and changing
only the GEMM accumulation order moves log_prob by 1e-4 (DESIGN.md "noise floor").  So parity is
gated two ways:
  * well-conditioned configs (synthetic parameters): the north_star tolerance itself;
  * the trained flow: the CUDA result must be as close to the float64 twin as the float32 oracle
    is (quantile ratio), and the fraction of rows inside the north_star tolerance is reported.
"""
import numpy as np

ZERO_PROB = 1e30
LP_UNDERFLOW = -100.0  # exp(lp) == 0 in float32 below ~-103: the density itself has underflowed
ATOL, RTOL = 1e-4, 1e-5  # north_star: 1e-4 absolute / 1e-5 relative


def zero_class(a, below=-ZERO_PROB):
    """The zero-probability class: NaN, -inf, finfo.min.  For log-probs pass
    below=LP_UNDERFLOW: rows far outside the training box are squashed onto the latent support
    boundary (|u| = 5 +- 1 ulp, lp ~ -200), where an ulp decides "finite but e^-200" vs "-inf";
    both mean probability 0.0 in float32, so they are one class."""
    a = np.asarray(a, dtype=np.float64)
    return np.isnan(a) | (a < below)


def within_tol(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) <= ATOL + RTOL * np.abs(b)


def summarize(name, cuda, o32, o64=None, below=-ZERO_PROB):
    cuda, o32 = np.asarray(cuda, np.float64), np.asarray(o32, np.float64)
    zc, zo = zero_class(cuda, below), zero_class(o32, below)
    fin = ~zc & ~zo & np.isfinite(cuda) & np.isfinite(o32)
    out = dict(name=name, n=int(cuda.size), class_mismatch=int((zc != zo).sum()),
               frac_within_tol=float(within_tol(cuda[fin], o32[fin]).mean()) if fin.any() else 1.0)
    e = np.abs(cuda - o32)[fin]
    out["med_vs_o32"], out["p99_vs_o32"], out["max_vs_o32"] = (
        (float(np.median(e)), float(np.quantile(e, 0.99)), float(e.max())) if e.size else (0.0, 0.0, 0.0))
    if o64 is not None:
        o64 = np.asarray(o64, np.float64)
        f2 = fin & ~zero_class(o64, below) & np.isfinite(o64)
        ec, eo = np.abs(cuda - o64)[f2], np.abs(o32 - o64)[f2]
        for q, tag in ((0.5, "med"), (0.99, "p99")):
            out[f"cuda_{tag}_vs_o64"] = float(np.quantile(ec, q))
            out[f"o32_{tag}_vs_o64"] = float(np.quantile(eo, q))
    print("PARITY", out)
    return out


def assert_tolerance(name, cuda, o32, min_frac=1.0, max_class_mismatch=0, below=-ZERO_PROB):
    """Well-conditioned case: every element inside the north_star tolerance."""
    s = summarize(name, cuda, o32, below=below)
    assert s["class_mismatch"] <= max_class_mismatch, s
    assert s["frac_within_tol"] >= min_frac, s
    return s


def assert_noise_floor(name, cuda, o32, o64, ratio=1.5, slack=2e-6, max_class_mismatch=None,
                       below=-ZERO_PROB):
    """Ill-conditioned case: CUDA is as close to the float64 twin as the float32 oracle is."""
    s = summarize(name, cuda, o32, o64, below=below)
    n = s["n"]
    limit = max(2, int(2e-3 * n)) if max_class_mismatch is None else max_class_mismatch
    assert s["class_mismatch"] <= limit, s  # support-boundary ties (|u| == B within an ulp)
    for tag in ("med", "p99"):
        assert s[f"cuda_{tag}_vs_o64"] <= ratio * s[f"o32_{tag}_vs_o64"] + slack, s
    return s


def count_bin_flips(bins_cuda, bins_oracle):
    b1, b2 = np.asarray(bins_cuda), np.asarray(bins_oracle)
    flips = b1 != b2
    return int(flips.sum()), int(flips.any(axis=1).sum()), int(b1.size)


# ---- knot ties ---------------------------------------------------------------------------------
def flip_tie_report(bins_cuda, trace, trace64=None, B=5.0):
    """For every bin index that differs from the oracle's: how far the searched value sits from the knot that
    separates the two bins, in float32 ulps of the spline range B (ulp(5.0) = 4.77e-7).  north_star allows flips
    "at knot ties" only: a flip must be between ADJACENT bins, and the value must sit on the separating knot to
    within the float32 uncertainty of that comparison.  That uncertainty is not 1 ulp: a stage's inputs and knots
    carry the rounding of every stage before it (the float32 oracle and its float64 twin disagree on them by up to
    hundreds of ulps by the last stage of the trained 7-D flow), so with `trace64` (the float64 twin's BinTrace of the
    same rows) each distance is also given relative to the stage's own noise floor
        noise_s = 4 ulp + q99 |value32 - value64| + q99 |knot32 - knot64|
    and `max_ratio` = max distance / noise_s.  `trace` / `trace64` are oracle BinTraces (stages in walk order)."""
    b1 = np.asarray(bins_cuda)
    b2 = np.concatenate(list(trace), axis=1)
    ulp = float(np.spacing(np.float32(B)))
    out = dict(total=int(b1.size), flips=0, rows=0, non_adjacent=0, ulps=[], ratios=[])
    flips = b1 != b2
    out["flips"], out["rows"] = int(flips.sum()), int(flips.any(axis=1).sum())
    col0 = 0
    for s, (knots, vals, idx) in enumerate(zip(trace.knots, trace.values, trace)):
        L = idx.shape[1]
        rr, dd = np.nonzero(flips[:, col0:col0 + L])
        noise = None
        if trace64 is not None and len(rr):
            fin = np.isfinite(vals) & np.isfinite(trace64.values[s])
            dv = np.abs(vals.astype(np.float64) - trace64.values[s])[fin]
            dk = np.abs(knots.astype(np.float64) - trace64.knots[s])
            dk = dk[np.isfinite(dk)]
            noise = 4 * ulp + (np.quantile(dv, 0.99) if dv.size else 0.0) + (np.quantile(dk, 0.99) if dk.size else 0.0)
        for r, d in zip(rr, dd):
            a, b = int(b1[r, col0 + d]), int(b2[r, d + col0])
            if abs(a - b) != 1 or not 0 <= max(a, b) < knots.shape[2]:
                out["non_adjacent"] += 1
                continue
            dist = abs(float(vals[r, d]) - float(knots[r, d, max(a, b)]))
            out["ulps"].append(dist / ulp)
            if noise is not None:
                out["ratios"].append(dist / noise)
        col0 += L
    u = np.asarray(out["ulps"], np.float64)
    out["max_ulps"] = float(u.max()) if u.size else 0.0
    out["median_ulps"] = float(np.median(u)) if u.size else 0.0
    out["within_4_ulps"] = int((u <= 4).sum())
    out["max_ratio"] = float(max(out["ratios"])) if out["ratios"] else 0.0
    del out["ratios"]
    out["ulps"] = [round(float(x), 1) for x in sorted(u)[-8:]]   # the largest few, for the report
    return out


_REPORT = {}


def record(config, **fields):
    """Collect per-config parity figures; tests/test_gpu_parity.py writes them to profiles/parity_r02.json."""
    _REPORT.setdefault(config, {}).update(fields)


def report():
    return _REPORT

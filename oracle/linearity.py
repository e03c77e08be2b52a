"""Linearity-analysis oracle.  TEST INFRASTRUCTURE ONLY.

Restates ``apply_thresholds`` (``/root/reference/modules/measurand.py:375-428``),
``compute_difference`` (``:620-655``) and ``compute_dimension_statistics`` (``:318-350``) as
``ExposureSeries.process_linearity`` chains them (``exposure_series.py:421-446``).  These reference
functions run unmodified at HEAD; ``tests/golden/make_golden.py`` pins this file to them.
"""
from __future__ import annotations

import warnings

import numpy as np


def apply_thresholds(val, std, lower, upper):
    n = val.shape[-1]
    lo = np.array([-np.inf if l is None else l for l in (lower or [None] * n)], dtype=val.dtype)
    hi = np.array([np.inf if u is None else u for u in (upper or [None] * n)], dtype=val.dtype)
    mask = (val < lo) | (val > hi)
    val = val.copy()
    val[mask] = np.nan
    if std is not None:
        std = std.copy()
        std[mask] = np.nan
    return val, std


def compute_difference(x_val, x_std, y_val, y_std, multiplier):
    scale_term = multiplier * y_val
    abs_diff = x_val - scale_term
    with np.errstate(all="ignore"):
        rel_diff = abs_diff / scale_term
    if x_std is None and y_std is None:
        return (abs_diff, None), (rel_diff, None)
    xs = 0 if x_std is None else x_std
    ys = 0 if y_std is None else y_std
    with np.errstate(all="ignore"):
        abs_std = np.sqrt(xs ** 2 + (multiplier * ys) ** 2)
        rel_std = np.sqrt((xs / (multiplier * y_val)) ** 2 + ((ys * x_val) / (multiplier * y_val ** 2)) ** 2)
    return (abs_diff, abs_std), (rel_diff, rel_std)


def dimension_statistics(val, std, axis=(0, 1)):
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if std is None:
            return {"mean": np.nanmean(val, axis=axis), "std": np.nanstd(val, axis=axis), "error": None}
        weights = 1 / std
        sum_w = np.nansum(weights, axis=axis)
        mean = np.nansum(val * weights, axis=axis) / sum_w
        spread = np.sqrt(np.nansum(weights * (val - mean) ** 2, axis=axis) / sum_w)
        return {"mean": mean, "std": spread, "error": np.nanmean(std, axis=axis)}


def pair_statistics(x_val, x_std, y_val, y_std, multiplier, lower=None, upper=None):
    """Thresholds (optional) -> difference -> statistics over the two spatial axes.
    Returns (absolute_stats, relative_stats), each a dict of (C,) arrays."""
    if lower is not None or upper is not None:
        x_val, x_std = apply_thresholds(x_val, x_std, lower, upper)
        y_val, y_std = apply_thresholds(y_val, y_std, lower, upper)
    (a, sa), (r, sr) = compute_difference(x_val, x_std, y_val, y_std, multiplier)
    return dimension_statistics(a, sa), dimension_statistics(r, sr)

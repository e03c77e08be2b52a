"""K4 oracle: ICRF calibration objective.  TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/modules/ICRF_calibration_exposure.py:20-44`` (candidate curve from
the PCA basis), ``:66-145`` (``analyze_linearity``), ``:148-201`` (``_energy_function``) and
``general_functions.py:149-176`` (``nanaverage``).  These reference functions run unmodified,
so this restatement is pinned bit-for-bit against them (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import warnings

import numpy as np


def candidate_curve(mean_icrf, pca_basis, params, use_mean_icrf: bool, bits: int = 256):
    """``ICRF_calibration_exposure.py:34-44``."""
    p = np.asarray(params)
    if not use_mean_icrf:
        base = np.linspace(0, 1, bits) ** p[0]
        return base + np.matmul(pca_basis, p[1:])
    return mean_icrf + np.matmul(pca_basis, p)


def nanaverage(values, weights, axis):
    """``general_functions.py:163-176``."""
    valid = ~np.isnan(values) & ~np.isnan(weights)
    weighted_sum = np.nansum(values * weights * valid, axis=axis)
    weight_sum = np.nansum(valid * weights, axis=axis)
    with np.errstate(invalid="ignore", divide="ignore"):
        result = weighted_sum / weight_sum
    result[weight_sum == 0] = np.nan
    return result


def analyze_linearity(value_stack, std_stack, lower, upper, use_relative, exposures):
    """``ICRF_calibration_exposure.py:81-145``.  value_stack is the ICRF-mapped (X, Y, N) stack."""
    if value_stack.ndim != 3:
        raise ValueError("image_stack must be a 3D CuPy array with shape (X, Y, N).")
    if exposures.ndim != 1 or exposures.size != value_stack.shape[2]:
        raise ValueError("exposure_values must be a 1D CuPy array matching the third dimension of image_stack.")
    use_std = std_stack is not None
    n = value_stack.shape[2]
    upper_pairs = np.triu_indices(n, k=1)
    ignored = np.tril_indices(n, k=0)

    masked = np.where((value_stack < lower) | (value_stack > upper), np.nan, value_stack)

    ratios = exposures[:, None] / exposures[None, :]
    ratios[ignored] = np.nan
    ratio_stack = np.expand_dims(ratios, axis=(0, 1))

    stack_i = np.expand_dims(masked, axis=3)
    stack_j = np.expand_dims(masked, axis=2)
    with np.errstate(invalid="ignore", divide="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        scaled = stack_j * ratio_stack
        diff = stack_i - scaled
        if use_relative:
            diff /= scaled
        abs_diff = np.abs(diff)

        if use_std:
            std_i = np.expand_dims(std_stack, axis=3)
            std_j = np.expand_dims(std_stack, axis=2)
            if use_relative:
                diff_std = np.sqrt((std_i / scaled) ** 2
                                   + ((stack_i * std_j) / (ratio_stack * stack_j ** 2)) ** 2)
            else:
                diff_std = np.sqrt(std_i ** 2 + (ratio_stack * std_j) ** 2)
            finite = np.logical_and(np.isfinite(abs_diff), diff_std != 0)
            weights = np.where(finite, 1 / diff_std, np.nan)
            results = nanaverage(abs_diff, weights, axis=(0, 1))
        else:
            results = np.nanmean(abs_diff, axis=(0, 1))
    return results[upper_pairs]


def energy(params, mean_icrf, pca_basis, dn_stack, std_stack, lower: int, upper: int,
           use_mean_icrf: bool, exposures, bits: int = 256) -> float:
    """``_energy_function`` (``ICRF_calibration_exposure.py:148-201``).  dn_stack must be an
    integer (X, Y, N) array (``:191`` indexes the curve with it)."""
    curve = candidate_curve(mean_icrf, pca_basis, params, use_mean_icrf, bits)
    curve += 1 - curve[-1]                                   # :167
    curve[0] = 0                                             # :168
    if np.max(curve) > 1 or np.min(curve) < 0:               # :174
        return np.inf
    if not np.all(curve[1:] > curve[:-1]):                   # :178
        return np.inf
    lo, hi = curve[lower], curve[upper]                      # :182-183
    mapped = curve[dn_stack]                                 # :191
    std_copy = std_stack.copy() if std_stack is not None else None
    pair_results = analyze_linearity(mapped, std_copy, lo, hi, True, exposures)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        e = np.nanmean(pair_results)                         # :196
    if np.isnan(e):
        e = np.inf
    return float(e)


def energy_population(population, mean_icrf, pca_basis, dn_stack, std_stack, lower, upper,
                      use_mean_icrf, exposures, bits: int = 256) -> np.ndarray:
    """SciPy ``vectorized=True`` calling convention: population is (n_params, S); returns (S,)."""
    population = np.asarray(population, dtype=np.float64)
    return np.array([energy(population[:, s].copy(), mean_icrf, pca_basis, dn_stack, std_stack,
                            lower, upper, use_mean_icrf, exposures, bits)
                     for s in range(population.shape[1])])

"""K2 oracle: weighted HDR merge of an N-exposure stack.  TEST INFRASTRUCTURE ONLY.

Restates, operation for operation, the reference's two streaming passes
(``/root/reference/modules/exposure_series.py:317-345`` sum of weights, ``:347-397`` merge,
``:399-419`` driver) and the Measurand methods they call (``measurand.py:606-618`` Gaussian
weight, ``:543-557`` bad-pixel median replace, ``:559-604`` flat-field normalisation,
``:471-541`` linearisation) with the repair set R1..R8 of SURVEY.md section 8.0.  Every line
that does arithmetic keeps the reference's operand order so the result is bit-identical to
"reference + repairs" under NumPy.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.ndimage import median_filter

from .linearize import linearize


def gaussian_weight(val: np.ndarray):
    """``measurand.py:615-616``: note ``np.e ** x`` (np.power), not ``np.exp``."""
    y = np.e ** (-30 * (val - 0.5) ** 2)
    dydx = -2 * 30 * (val - 0.5) * y
    return y, dydx


def bad_pixel_filter(val: np.ndarray, std: np.ndarray | None, dark_val: np.ndarray,
                     threshold: float, kernel: int = 3):
    """``measurand.py:543-557`` with repairs R6 (call the median filter as a plain function) and
    R6/D8 (``where(mask, median, original)``).  Median is per channel, K x K, scipy 'reflect'
    (half-sample symmetric) boundary, rank K*K//2."""
    hot = dark_val > threshold
    med = median_filter(val, size=(kernel, kernel), axes=(0, 1), mode="reflect")
    out_val = np.where(hot, med, val)
    out_std = None
    if std is not None:
        med_s = median_filter(std, size=(kernel, kernel), axes=(0, 1), mode="reflect")
        out_std = np.where(hot, med_s, std)
    return out_val, out_std


def flat_roi_bounds(im_size_x: int, im_size_y: int, mid_fraction: float):
    """ROI of ``measurand.py:569-576`` with repair R7/D10 (``int()`` the float bounds).  The
    reference slices axis 0 with the IM_SIZE_X-derived bounds and axis 1 with the
    IM_SIZE_Y-derived ones (its literal formula, kept)."""
    roi_dx = math.floor(im_size_x * mid_fraction)
    roi_dy = math.floor(im_size_y * mid_fraction)
    start = (math.floor(1 / mid_fraction) - 1) / 2
    r0, r1 = int(start * roi_dx), int((start + 1) * roi_dx)
    c0, c1 = int(start * roi_dy), int((start + 1) * roi_dy)
    return r0, r1, c0, c1


def flat_field_means(flat: np.ndarray, roi):
    r0, r1, c0, c1 = roi
    return np.mean(flat[r0:r1, c0:c1, :], axis=(0, 1))     # measurand.py:579


def normalize_by_map(val, std, flat_val, flat_std, roi, means=None):
    """``measurand.py:559-604`` (flat-field correction with uncertainty), repair R7.

    ``means = (m, ms)`` overrides the ROI means -- used by the tests to check a ROW CROP of a large
    image, where the ROI lies outside the crop and the means come from the full flat field."""
    if means is None:
        m = flat_field_means(flat_val, roi)
        ms = flat_field_means(flat_std, roi)
    else:
        m, ms = means

    u_acq = (std ** 2) / (flat_val ** 2)
    u_acq *= m ** 2

    u_ff = (val ** 2) / (flat_val ** 4)
    u_ff *= flat_std ** 2
    u_ff *= m ** 2

    u_ffm = (val ** 2) / (flat_val ** 2)
    u_ffm *= ms ** 2

    out_std = np.sqrt(u_acq + u_ff + u_ffm)
    out_val = (val / flat_val) * m
    return out_val, out_std


def hdr_merge(dn_stack, std_stack, exposures, icrf, icrf_diff, *, max_dn: int = 255,
              darks=None, dark_threshold: float = 0.05, kernel: int = 3,
              flat_val=None, flat_std=None, roi=None, flat_means=None):
    """Merge N exposures into an HDR radiance image with its uncertainty.

    dn_stack   : sequence of N integer images (H, W, C) -- what ``cv.imread`` returns
                 (``image_set.py:223``); they are converted with ``.astype(f64) / MAX_DN``.
    std_stack  : sequence of N float64 uncertainty images (H, W, C) (``image_set.py:228-243``).
    exposures  : N exposure times in seconds, ascending (``exposure_series.py:143,200``).
    darks      : None, or a sequence of N entries, each None or the dark-frame VALUE image
                 (float64, already ``dark_dn / MAX_DN`` and exposure-scaled per repair R8) that
                 ``get_dark_field`` selected for that exposure (``image_set.py:157-198``).
    flat_val / flat_std / roi : optional flat field (repair R7, ``exposure_series.py:415-417``).
    Returns (hdr_val, hdr_std) float64 (H, W, C).
    """
    n = len(dn_stack)
    values, stds = [], []
    for k in range(n):
        v = dn_stack[k].astype(np.float64) / max_dn                  # image_set.py:223
        s = std_stack[k]
        if darks is not None and darks[k] is not None:               # R5: filter result is used
            v, s = bad_pixel_filter(v, s, darks[k], dark_threshold, kernel)
        values.append(v)
        stds.append(s)

    # pass 1 -- exposure_series.py:328-343 (R3: plain arrays; R4: zero accumulators)
    sum_w = np.zeros_like(values[0])
    for k in range(n):
        sum_w += gaussian_weight(values[k])[0]
    sq_sum_w = sum_w ** 2

    # pass 2 -- exposure_series.py:372-394
    hdr_val = np.zeros_like(values[0])
    hdr_std = np.zeros_like(values[0])
    for k in range(n):
        w, dw = gaussian_weight(values[k])
        g, dg = linearize(values[k], stds[k], icrf, icrf_diff, max_dn)   # :383 (+R1)
        t = exposures[k]
        hdr_val += (w * g) / (sum_w * t)                                   # :388
        hdr_std += (((dw * g + w * dg) / sum_w - (dw * w * g) / sq_sum_w) * dg / t) ** 2  # :389
    hdr_std = hdr_std ** (1 / 2)                                           # :394

    if flat_val is not None:
        hdr_val, hdr_std = normalize_by_map(hdr_val, hdr_std, flat_val, flat_std, roi, flat_means)
    return hdr_val, hdr_std


def select_dark_field(target_exposure: float, dark_exposures, exposure_threshold: float):
    """``ImageSet.get_dark_field`` (``image_set.py:171-198``) as a pure function.

    Returns None (no dark applies), or ``(index, scale)``: the dark frame to use and the factor
    its VALUE image is multiplied by (repair R8: ``target / original``; 1.0 for an exact match).
    The scan order, the "last greater index seen" choice and the early return as soon as both a
    shorter and a longer dark have been seen are the reference's literal behaviour.
    """
    if not target_exposure >= exposure_threshold:
        return None
    lesser = greater = False
    greater_index = 0
    for i, exposure in enumerate(dark_exposures):
        if exposure < target_exposure:
            lesser = True
        if exposure > target_exposure:
            greater = True
            greater_index = i
        if target_exposure == exposure:
            return i, 1.0
        if lesser and greater:
            return greater_index, target_exposure / dark_exposures[greater_index]
    return None


def dark_value_image(dark_dn: np.ndarray, scale: float, max_dn: int = 255) -> np.ndarray:
    """Dark VALUE image as the reference holds it: ``imread/MAX_DN`` (``image_set.py:223``), then
    for a scaled dark ``(target/original) * measurand`` -> ``__rmul__`` -> ``val * [scale]``
    (``image_set.py:260``, ``measurand.py:213-215,198``).  An exact match is not multiplied."""
    v = dark_dn.astype(np.float64) / max_dn
    if scale != 1.0:
        v = v * np.array([scale], dtype=np.float64)
    return v

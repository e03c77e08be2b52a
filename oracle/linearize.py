"""K1 oracle: ICRF linearisation by LUT gather.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``/root/reference/modules/measurand.py:471-541`` (``linearize``,
``_linearize_channel``, ``_linearize_single``) with repair R1 of SURVEY.md section 8.0:
the multi-channel gather is ``ICRF[DN, arange(C)]`` exactly as the working sibling
``video_processing.py:201`` does, instead of the broken ``ICRF[DN[..., None], self.channels]``
(``measurand.py:507,512``; the reference's own ``test_linearize`` fails on it).
"""
from __future__ import annotations

import numpy as np


def lut_index(val: np.ndarray, max_dn: int = 255) -> np.ndarray:
    """LUT bin of every sample (``measurand.py:502-505`` / ``:530-533``).

    Integer images index the LUT directly.  Floating images are mapped with
    ``around(val * MAX_DN).astype(uint8)``: round-half-even, then a wrapping cast.  The reference
    hard-codes uint8; for LUTs longer than 256 rows (16-bit data, an extension the reference's
    float path cannot express) the cast widens to uint16.
    """
    if np.issubdtype(val.dtype, np.integer):
        return val.copy()
    index_dtype = np.uint8 if max_dn <= 255 else np.uint16
    with np.errstate(invalid="ignore"):
        return np.around(val * max_dn).astype(index_dtype)


def linearize(val: np.ndarray, std: np.ndarray | None, icrf: np.ndarray,
              icrf_diff: np.ndarray | None = None, max_dn: int = 255):
    """Return ``(val_out, std_out)``; ``std_out`` is None unless both std and icrf_diff are given
    (``measurand.py:498-500``)."""
    use_std = std is not None and icrf_diff is not None
    bins = lut_index(val, max_dn)
    if val.shape[-1] >= 2:                      # measurand.py:482-485 dispatch
        channels = np.arange(val.shape[-1])     # R1 / video_processing.py:201
        out = icrf[bins, channels]
        if not use_std:
            return out, None
        return out, icrf_diff[bins, channels] * std
    out = icrf[bins]                            # measurand.py:534
    if not use_std:
        return out, None
    return out, icrf_diff[bins] * std           # measurand.py:539


def default_icrf_diff(icrf: np.ndarray, bits: int = 256) -> np.ndarray:
    """Repair R2: per-channel ``np.gradient(ICRF[:, c], 2/(BITS-1))`` as in
    ``general_functions.py:269-272`` and ``tests/unit/test_measurand.py:21`` (the 2/(BITS-1)
    spacing is reference behaviour and is kept)."""
    dx = 2 / (bits - 1)
    if icrf.ndim == 1:
        return np.gradient(icrf, dx)
    out = np.zeros_like(icrf)
    for c in range(icrf.shape[1]):
        out[:, c] = np.gradient(icrf[:, c], dx)
    return out

"""Oracle for the per-channel histogram (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``AbstractMeasurand.compute_channel_histogram`` (``modules/measurand.py:430-469``)."""
from __future__ import annotations

import numpy as np


def channel_histogram(val, std, channel: int, bins: int, included_range=None, use_std: bool = False):
    v = np.asarray(val)[..., channel]                               # :450
    mask = np.isfinite(v)                                           # :451
    weights = None
    if use_std:
        s = np.asarray(std)[..., channel]                           # :454
        mask = np.logical_and(mask, s != 0)                         # :455-456
        weights = 1 / s[mask]                                       # :457-458
    return np.histogram(v[mask], bins=bins, range=included_range, weights=weights)   # :464

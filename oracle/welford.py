"""K3 oracle: Welford mean / standard-error frame over a video.  TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/modules/video_processing.py:183-217`` with repair R9 of SURVEY.md
section 8.0: ``if ICRF is not None`` instead of the raising ``if ICRF:`` (``:200``), and the
float64 mean / standard error of the mean are returned un-quantised next to the literal uint8
outputs (the reference rounds the SEM, which lives in [0, 1] units, to uint8 -> all zeros,
``:214-215``; README.md:19 says uncertainty images are 64-bit).  The arithmetic of lines
205-208 is kept in the reference's exact sequential order -- the uint8 mean depends on it.
"""
from __future__ import annotations

import numpy as np


def welford(frames, icrf: np.ndarray | None = None, use_std: bool = True, max_dn: int = 255):
    """frames: iterable of (H, W, C) uint8 frames.  Returns a dict with

    mean_u8 / std_u8 : the reference's literal return values (``video_processing.py:210-217``)
    mean / sem       : float64 running mean and sqrt(M2/(n-1))/sqrt(n) before quantisation
    m2, count        : raw Welford state after the last frame
    """
    mean = None
    m2 = None
    count = 0
    for frame in frames:
        if frame is None:
            break
        count += 1
        if mean is None:
            mean = np.zeros(frame.shape, dtype=np.float64)
            if use_std:
                m2 = np.zeros(frame.shape, dtype=np.float64)
        if icrf is not None:                                       # R9
            x = icrf[frame, np.arange(frame.shape[-1])]            # :201
        else:
            x = (frame / max_dn).astype(np.float64)                # :203
        delta = x - mean                                           # :205
        mean = mean + delta / count                                # :206
        if use_std:
            m2 = m2 + delta * (x - mean)                           # :208

    mean_u8 = np.around(mean * max_dn).astype(np.uint8)            # :210-211
    sem = None
    std_u8 = None
    if use_std:
        with np.errstate(invalid="ignore", divide="ignore"):
            sem = np.sqrt(m2 / (count - 1)) / np.sqrt(count)       # :214
            std_u8 = np.around(sem).astype(np.uint8)               # :215
    return {"mean_u8": mean_u8, "std_u8": std_u8, "mean": mean, "sem": sem, "m2": m2,
            "count": count}

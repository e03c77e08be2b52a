"""Oracle for the 8-bit export (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the array part of ``ImageSet.save_8bit`` (``modules/image_set.py:343-350`` for the value image,
``:352-357`` for a forced 8-bit uncertainty image): the file write itself is OpenCV's.
"""
from __future__ import annotations

import numpy as np


def quantize_8bit(val: np.ndarray, max_dn: float = 255.0) -> np.ndarray:
    """image_set.py:343-350."""
    val = np.array(val, dtype=np.float64, copy=True)
    max_float = np.amax(val)                      # :345
    if max_float > 1:                             # :347
        val /= max_float                          # :348
    with np.errstate(invalid="ignore"):
        return np.around(val * max_dn).astype(np.dtype('uint8'))   # :350

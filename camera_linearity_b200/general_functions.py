"""Host-side helpers shared by the modules (reference: modules/general_functions.py)."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np
import torch

from .settings import GlobalSettings as gs


def is_broadcastable(shape1, shape2) -> bool:
    """general_functions.py:14-24 (error contract pinned by tests/unit/test_general_functions.py)."""
    if not shape1 or not shape2:
        raise ValueError('Shapes cannot be empty')
    for a, b in zip(tuple(shape1)[::-1], tuple(shape2)[::-1]):
        if not (a == 1 or b == 1 or a == b):
            return False
    return True


def choose_evenly_spaced_points(array, step_x: int, step_y: Optional[int] = None):
    """Strided subsampling of the two leading axes (general_functions.py:27-44)."""
    if step_y is None:
        step_y = step_x
    return array[::step_x, ::step_y, ...]


def predict_output_shape(input_shape, step_x: int, step_y: Optional[int] = None):
    if step_y is None:
        step_y = step_x
    rows, cols = input_shape
    return (rows + step_x - 1) // step_x, (cols + step_y - 1) // step_y


def map_linearity_limits(lower_limit: Optional[int], upper_limit: Optional[int], ICRF=None):
    """general_functions.py:97-129: DN limits mapped through the ICRF (per channel)."""
    n = gs.NUM_OF_CHS
    lower = [float(gs.LOWER_LIN_LIM if lower_limit is None else lower_limit)] * n
    upper = [float(gs.UPPER_LIN_LIM if upper_limit is None else gs.MAX_DN - upper_limit)] * n
    if ICRF is None:
        return [x / gs.MAX_DN for x in lower], [x / gs.MAX_DN for x in upper]
    table = ICRF.detach().cpu().numpy() if isinstance(ICRF, torch.Tensor) else np.asarray(ICRF)
    return ([float(table[int(lower[c]), c]) for c in range(n)],
            [float(table[int(upper[c]), c]) for c in range(n)])


def icrf_derivative(ICRF):
    """Repair R2: ``np.gradient(ICRF[:, c], 2/(BITS-1))`` per channel, as in
    general_functions.py:269-272 (the 2/(BITS-1) spacing is reference behaviour)."""
    table = ICRF.detach().cpu().numpy() if isinstance(ICRF, torch.Tensor) else np.asarray(ICRF, dtype=np.float64)
    dx = 2 / (table.shape[0] - 1)
    if table.ndim == 1:
        return np.gradient(table, dx)
    out = np.zeros_like(table)
    for c in range(table.shape[1]):
        out[:, c] = np.gradient(table[:, c], dx)
    return out


def read_txt_to_array(file_name: str, path: Optional[str] = None, use_cupy: bool = True):
    """general_functions.py:280-302; returns a NumPy array (host) -- tables are tiny."""
    load_path = gs.DATA_PATH if path is None else Path(path)
    return np.loadtxt(Path(load_path).joinpath(file_name), dtype=float)


def read_ICRF_file(file_path, return_derivative: bool = True, use_cupy: bool = False):
    """general_functions.py:254-277 with defect D12 repaired (the derivative is returned)."""
    icrf = np.loadtxt(file_path, dtype=float)
    if not return_derivative:
        return icrf, None
    return icrf, icrf_derivative(icrf)


def video_frame_generator(video_path):
    """general_functions.py:226-251 (OpenCV decode on the host)."""
    import cv2 as cv
    video = cv.VideoCapture(str(video_path))
    if not video.isOpened():
        raise ValueError(f'Unable to open video file at {video_path}')
    try:
        while True:
            ret, frame = video.read()
            if not ret:
                yield None
                break
            yield frame
    finally:
        video.release()

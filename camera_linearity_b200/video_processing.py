"""Mean / standard-error frames of videos (reference: ``modules/video_processing.py:161-274``).

``welford_algorithm`` keeps the reference's signature.  Frames are decoded on the host (OpenCV),
moved to the GPU in chunks and folded into the running float64 (mean, M2) state with the exact
sequential Welford recurrence (``ops.welford_update``, bit-identical to NumPy).  When a whole
stack is already resident on the device, ``welford_stack`` uses the integer fast path (K3).
Repair R9: ``if ICRF is not None``; the float64 mean and SEM are returned next to
the reference's uint8 outputs.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Union

import numpy as np
import torch

from . import general_functions as gf
from . import ops
from .settings import GlobalSettings as gs

CHUNK_FRAMES = 32


def welford_stack(frames, ICRF=None):
    """frames: (F, H, W, C) uint8 tensor/array already in memory.  Returns the same dict as
    ``welford_algorithm``."""
    dev = gs.device()
    fr = torch.as_tensor(frames).to(dev)
    icrf = None if ICRF is None else torch.as_tensor(ICRF, dtype=torch.float64, device=dev)
    mean, sem, mean_u8 = ops.welford_stack(fr, icrf, gs.MAX_DN)
    std_u8 = torch.round(sem).nan_to_num(0.0).to(torch.uint8)       # video_processing.py:215 (D13)
    return {'mean': mean_u8, 'std': std_u8, 'mean_f64': mean, 'sem': sem, 'count': int(fr.shape[0])}


def welford_algorithm(file_paths: Union[Path, List[Path]], ICRF=None, use_std: Optional[bool] = False,
                      frame_source=None):
    """Welford mean / std frame over all frames of one or more videos (video_processing.py:161-219).

    Returns ``{'mean': uint8 mean frame, 'std': uint8 frame or None}`` exactly like the reference,
    plus ``'mean_f64'``, ``'sem'`` (float64, un-quantised; repair R9) and ``'count'``.
    ``frame_source(path)`` may replace the OpenCV frame generator (used by the tests and bench).
    """
    if not isinstance(file_paths, list):
        file_paths = [file_paths]
    source = gf.video_frame_generator if frame_source is None else frame_source
    dev = gs.device()
    icrf = None if ICRF is None else torch.as_tensor(ICRF, dtype=torch.float64, device=dev)
    mean = m2 = None
    count = 0
    pending: List[np.ndarray] = []

    def flush():
        nonlocal mean, m2, count
        if not pending:
            return
        chunk = torch.from_numpy(np.ascontiguousarray(np.stack(pending))).to(dev, non_blocking=True)
        pending.clear()
        if mean is None:
            mean = torch.zeros(chunk.shape[1:], dtype=torch.float64, device=dev)
            m2 = torch.zeros_like(mean)
        count = ops.welford_update(chunk, mean, m2, count, icrf, gs.MAX_DN)

    for file_path in file_paths:
        for frame in source(file_path):
            if frame is None:
                break
            pending.append(frame)
            if len(pending) == CHUNK_FRAMES:
                flush()
    flush()
    if mean is None:
        raise ValueError("no frames decoded")
    sem, mean_u8 = ops.welford_finalize(mean, m2 if use_std else None, count, gs.MAX_DN)
    std_u8 = None
    if use_std:
        std_u8 = torch.round(sem).nan_to_num(0.0).to(torch.uint8)   # :214-215 literal (all zeros, D13)
    return {'mean': mean_u8, 'std': std_u8, 'mean_f64': mean, 'sem': sem, 'count': count}


def process_video(video_path: Path, ICRF=None, use_std: Optional[bool] = True):
    """video_processing.py:222-236."""
    import cv2 as cv
    ret = welford_algorithm(video_path, ICRF, use_std)
    for key in ('mean', 'std'):
        if ret[key] is not None:
            save_path = str(video_path.parent.joinpath(video_path.name.replace('.avi', f'.{key}.tif')))
            cv.imwrite(save_path, ret[key].cpu().numpy())


def process_directory(dir_path: Path, ICRF=None, separately: Optional[bool] = True):
    """video_processing.py:239-274."""
    import cv2 as cv
    video_files = list(dir_path.glob("*.avi"))
    if not separately:
        ret = welford_algorithm(video_files, ICRF)
        for key in ('mean', 'std'):
            if ret[key] is not None:
                cv.imwrite(str(dir_path.joinpath(f'total_{key}.tif')), ret[key].cpu().numpy())
        return
    for path in video_files:
        ret = welford_algorithm(path, ICRF)
        for key in ('mean', 'std'):
            if ret[key] is not None:
                save_dir = path.parent.joinpath(key)
                save_dir.mkdir(exist_ok=True)
                name = path.name.replace('.avi', ' STD.tif' if key == 'std' else '.tif')
                cv.imwrite(str(save_dir.joinpath(name)), ret[key].cpu().numpy())

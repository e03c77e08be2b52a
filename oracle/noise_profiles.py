"""Noise-profile oracle: joint histogram of (uint8 mean-frame DN, frame DN) per channel.  TEST INFRASTRUCTURE ONLY.

Restates ``compute_noise_profiles`` (``/root/reference/modules/video_processing.py:77-106``): the mean frame is the
uint8 ``welford_algorithm(video_files, None, False)['mean']`` and every frame of every video is scattered with
``np.add.at(noise_profiles[:, :, c], (mean_channel, frame_channel), 1)``.  The reference function runs unmodified
(only the frame source is replaced), so this restatement is pinned against it: ``tests/golden/k9_noise_profiles.npz``.
"""
from __future__ import annotations

import numpy as np

from .welford import welford


def noise_profiles(videos, bits: int = 256):
    """videos: sequence of frame sequences, each frame (H, W, C) uint8.  Returns (profiles int64 (bits, bits, C),
    mean frame uint8)."""
    all_frames = [f for video in videos for f in video]
    mean_frame = welford(all_frames, None, False)["mean_u8"]                    # :90
    channels = mean_frame.shape[-1]
    profiles = np.zeros((bits, bits, channels), dtype=np.int64)                 # :88
    for video in videos:                                                        # :92-104
        for frame in video:
            for c in range(channels):
                np.add.at(profiles[:, :, c], (mean_frame[..., c].flatten(), frame[..., c].flatten()), 1)
    return profiles, mean_frame

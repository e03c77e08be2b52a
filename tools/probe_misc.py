"""Timings of the stand-alone / secondary entry points on 4K RGB data (development tool)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    H, W, C = 2160, 3840, 3
    n = H * W * C
    x = np.linspace(0, 1, 256)
    icrf = torch.from_numpy(np.stack([x ** (2.0 + 0.1 * c) for c in range(C)], 1)).to(dev)
    diff = torch.from_numpy(np.stack([np.gradient(x ** (2.0 + 0.1 * c), 2 / 255) for c in range(C)], 1)).to(dev)
    val = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64)
    std = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.02
    dn = (val * 255).round().to(torch.uint8)
    rows = []
    ms = timed(lambda: ops.linearize(dn, std, icrf, diff, 255.0)); rows.append(("linearize u8 + std", ms, n * 25))
    ms = timed(lambda: ops.linearize(val, std, icrf, diff, 255.0)); rows.append(("linearize f64 + std", ms, n * 32))
    ms = timed(lambda: ops.linearize(dn, None, icrf, None, 255.0)); rows.append(("linearize u8, no std", ms, n * 9))
    ms = timed(lambda: ops.gaussian_weight(val)); rows.append(("gaussian_weight (w, dw)", ms, n * 24))
    dark = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.04
    dark[torch.rand((H, W, C), generator=g, device=dev) < 0.001] = 0.5
    ms = timed(lambda: ops.bad_pixel_filter(val, std, dark, 0.05, 3)); rows.append(("bad_pixel_filter K=3", ms, n * 40))
    flat = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.2 + 0.6
    means = torch.tensor([0.7, 0.7, 0.7, 0.005, 0.005, 0.005], dtype=torch.float64, device=dev)
    ms = timed(lambda: ops.flat_field_normalize(val, std, flat, std, means)); rows.append(("flat_field_normalize", ms, n * 48))
    ms = timed(lambda: ops.quantize_8bit(val)); rows.append(("quantize_8bit", ms, n * 17))
    frames = (torch.rand((32, 1080, 1920, 3), generator=g, device=dev) * 255).to(torch.uint8)
    mean = torch.zeros((1080, 1920, 3), dtype=torch.float64, device=dev)
    m2 = torch.zeros_like(mean)
    ms = timed(lambda: ops.welford_update(frames, mean, m2, 0, None, 255.0), reps=5)
    rows.append(("welford_update 32 frames 1080p", ms, frames.numel() + mean.numel() * 32))
    ms = timed(lambda: ops.channel_histogram(val, std, 1, 256, (0.0, 1.0)), reps=5); rows.append(("channel_histogram 256 bins", ms, n // 3 * 16))
    for name, ms, nb in rows:
        print(f"{name:34s} {ms:8.3f} ms  {nb / ms / 1e6:8.0f} GB/s")


if __name__ == "__main__":
    main()

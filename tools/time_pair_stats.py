"""Time ops.pair_statistics on one synthetic 4K RGB exposure pair (CUDA events) and check it against a torch
float64 evaluation of the same chain (thresholds -> scaled difference -> inverse-sigma weighted statistics).

    python tools/time_pair_stats.py [reps]          (last line: "<ms> ms  <GB/s>  maxrel <err>")
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def torch_reference(xv, xs, yv, ys, m, lo, hi):
    """measurand.py:375-428, 620-655, 318-350 as plain torch float64 (NaN-skipping sums)."""
    nan = float("nan")
    bx = (xv < lo) | (xv > hi)
    by = (yv < lo) | (yv > hi)
    xv = torch.where(bx, nan, xv); xs = torch.where(bx, nan, xs)
    yv = torch.where(by, nan, yv); ys = torch.where(by, nan, ys)
    scale = m * yv
    a = xv - scale
    r = a / scale
    sa = torch.sqrt(xs ** 2 + (m * ys) ** 2)
    sr = torch.sqrt((xs / scale) ** 2 + (ys * xv / (m * yv * yv)) ** 2)
    out = []
    for v, s in ((a, sa), (r, sr)):
        w = 1.0 / s
        sw = torch.nansum(w, dim=(0, 1))
        mean = torch.nansum(v * w, dim=(0, 1)) / sw
        var = torch.nansum(w * (v - mean) ** 2, dim=(0, 1)) / sw
        err = torch.nansum(s, dim=(0, 1)) / (~torch.isnan(s)).sum(dim=(0, 1))
        out.append(torch.stack([mean, torch.sqrt(var), err]))
    return torch.stack(out)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    shape = (2160, 3840, 3)
    xv = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) + 0.05
    yv = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) + 0.05
    xs = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 0.02 + 0.001
    ys = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 0.02 + 0.001
    lo, hi = [0.1] * 3, [0.9] * 3
    for _ in range(5):
        out = ops.pair_statistics(xv, xs, yv, ys, 0.5, lo, hi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ops.pair_statistics(xv, xs, yv, ys, 0.5, lo, hi)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    again = ops.pair_statistics(xv, xs, yv, ys, 0.5, lo, hi)
    ref = torch_reference(xv, xs, yv, ys, 0.5, 0.1, 0.9)
    rel = ((out.reshape(ref.shape) - ref).abs() / ref.abs()).max().item()
    gbs = xv.numel() * 32 / ms / 1e6
    import hashlib
    digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12]          # equal across builds = bit-identical
    print(f"{ms:.4f} ms  {gbs:.0f} GB/s  maxrel {rel:.2e}  repeat_identical {bool(torch.equal(out, again))}  sha1 {digest}")


if __name__ == "__main__":
    main()

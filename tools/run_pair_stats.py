"""Run ops.pair_statistics on one synthetic 4K RGB exposure pair (profiling target for ncu).

    python tools/run_pair_stats.py [reps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    shape = (2160, 3840, 3)
    xv = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) + 0.05
    yv = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) + 0.05
    xs = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 0.02 + 0.001
    ys = torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 0.02 + 0.001
    for _ in range(reps):
        out = ops.pair_statistics(xv, xs, yv, ys, 0.5, [0.1] * 3, [0.9] * 3)
    torch.cuda.synchronize()
    print(out.cpu().numpy())


if __name__ == "__main__":
    main()

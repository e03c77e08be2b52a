"""Run ops.hdr_merge on one cfg5-like 16-bit stack (12 x 7680 x 4320 x 3 uint16 + f64 std); ncu target.

    python tools/run_merge16.py [algo] [reps] [height]
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def main():
    algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 4320
    W, N = 7680, 12
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    x16 = np.linspace(0, 1, 65536)
    icrf = torch.from_numpy(np.stack([x16 ** (2.0 + 0.1 * c) for c in range(3)], axis=1)).to(dev)
    diff = torch.from_numpy(np.stack([np.gradient(x16 ** (2.0 + 0.1 * c), 2 / 65535) for c in range(3)], axis=1)).to(dev)
    rad = torch.rand((H, W, 3), generator=g, device=dev, dtype=torch.float32) * 25
    t = [0.002 * 1.6 ** k for k in range(N)]
    dn = [torch.round(65535 * torch.clamp(rad * tk, 0, 1) ** (1 / 2.2)).to(torch.int32).to(torch.uint16) for tk in t]
    del rad
    std = [torch.rand((H, W, 3), generator=g, device=dev, dtype=torch.float64) * 0.018 + 0.002 for _ in t]
    out = (torch.empty((H, W, 3), dtype=torch.float64, device=dev), torch.empty((H, W, 3), dtype=torch.float64, device=dev))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=algo)
        ev[r + 1].record()
    torch.cuda.synchronize()
    print("algo", algo, "ms per call:", [round(ev[r].elapsed_time(ev[r + 1]), 3) for r in range(reps)])


if __name__ == "__main__":
    main()

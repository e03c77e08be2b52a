"""Run the Welford stack kernel on cfg4 (600 x 1080x1920x3 uint8), with or without an ICRF; ncu target.

    python tools/run_k3.py [icrf:0|1] [reps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def main():
    use_icrf = len(sys.argv) > 1 and sys.argv[1] == "1"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    icrf_np, _ = bench.icrf_tables(3)
    icrf = torch.from_numpy(icrf_np).to(dev)
    base = torch.randint(20, 231, (1, 1080, 1920, 3), generator=g, device=dev, dtype=torch.int16)
    frames = torch.empty((600, 1080, 1920, 3), dtype=torch.uint8, device=dev)
    for f0 in range(0, 600, 50):
        noise = torch.round(torch.randn((50, 1080, 1920, 3), generator=g, device=dev) * 3).to(torch.int16)
        frames[f0:f0 + 50] = torch.clamp(base + noise, 0, 255).to(torch.uint8)
    del noise
    ws = torch.empty(ops._lib.load().cl_welford_stack_workspace_bytes(600, frames[0].numel()), dtype=torch.uint8, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        ops.welford_stack(frames, icrf if use_icrf else None, 255.0, ws)
        ev[r + 1].record()
    torch.cuda.synchronize()
    print("icrf", use_icrf, "ms per call:", [round(ev[r].elapsed_time(ev[r + 1]), 3) for r in range(reps)])


if __name__ == "__main__":
    main()

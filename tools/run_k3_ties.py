"""Welford stack on a video in which a chosen fraction of the samples is an exact rounding tie (frames alternate
d, d + 1 there): times the tie replay.    python tools/run_k3_ties.py [fraction] [frames]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def main():
    frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    shape = (1080, 1920, 3)
    base = torch.randint(20, 231, shape, generator=g, device=dev, dtype=torch.uint8)
    tie = torch.rand(shape, generator=g, device=dev) < frac
    frames = base.unsqueeze(0).repeat(F, 1, 1, 1)
    frames[1::2] += tie.to(torch.uint8)
    for _ in range(2):
        ops.welford_stack(frames)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for r in range(3):
        mean, sem, mean_u8 = ops.welford_stack(frames)
        ev[r + 1].record()
    torch.cuda.synchronize()
    print(f"ties {int(tie.sum())} of {tie.numel()}  ms per call:", [round(ev[r].elapsed_time(ev[r + 1]), 3) for r in range(3)])


if __name__ == "__main__":
    main()

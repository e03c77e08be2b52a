"""cfg5 merge time on a spatially smooth scene (what a camera delivers) against bench's i.i.d.-random radiance, which is
the worst case for the 16-bit kernel: its bound is one L1 wavefront per DISTINCT table row a warp gathers.

    python tools/cfg5_smooth.py
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, W, N = bench.CFG5["H"], bench.CFG5["W"], bench.CFG5["N"]
icrf, diff, stdlut = bench.cfg5_tables(dev)
t = [bench.CFG5["t0"] * bench.CFG5["ratio"] ** k for k in range(N)]
g = torch.Generator(device=dev).manual_seed(7)
std = [torch.rand((H, W, 1), generator=g, device=dev, dtype=torch.float64) * 0.018 + 0.002 for _ in t]
out = (torch.empty((H, W, 1), dtype=torch.float64, device=dev), torch.empty((H, W, 1), dtype=torch.float64, device=dev))


def stack_from(rad):
    return [torch.round(65535 * torch.clamp(rad * tk, 0, 1) ** (1 / 2.2)).to(torch.int32).to(torch.uint16) for tk in t]


def timed(dn, label):
    for _ in range(2):
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for r in range(5):
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(5))
    print(f"{label:64s} {ms[2]:.3f} ms")


rad = torch.rand((H, W, 1), generator=g, device=dev, dtype=torch.float32) * 25
timed(stack_from(rad), "i.i.d. random radiance (bench data)")
for cells, noise in ((64, 0.0), (64, 0.01), (512, 0.01), (512, 0.05)):
    coarse = torch.rand((1, 1, cells * H // W + 2, cells), generator=g, device=dev, dtype=torch.float32)
    smooth = F.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)[0, 0].unsqueeze(-1) * 25
    smooth = smooth * (1 + noise * torch.randn((H, W, 1), generator=g, device=dev, dtype=torch.float32))
    timed(stack_from(smooth.clamp_min(0)), f"smooth scene: {cells} random cells across the width, {noise * 100:.0f} % pixel noise")

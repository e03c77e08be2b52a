import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from camera_linearity_b200 import ops
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
for C, W in ((1, 3840 * 3), (3, 3840)):
    H, N = 2160, 16
    x = np.linspace(0, 1, 256)
    icrf = torch.from_numpy(np.stack([x ** (2.0 + 0.1 * c) for c in range(C)], 1)).to(dev)
    diff = torch.from_numpy(np.stack([np.gradient(x ** (2.0 + 0.1 * c), 2 / 255) for c in range(C)], 1)).to(dev)
    t = [0.001 * 1.6 ** k for k in range(N)]
    rad = torch.rand((H, W, C), generator=g, device=dev) * 25
    dn = [torch.round(255 * torch.clamp(rad * tk, 0, 1) ** (1 / 2.2)).to(torch.uint8) for tk in t]
    std = [torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.018 + 0.002 for _ in t]
    out = (torch.empty((H, W, C), dtype=torch.float64, device=dev), torch.empty((H, W, C), dtype=torch.float64, device=dev))
    for _ in range(3): ops.hdr_merge(dn, std, t, icrf, diff, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.hdr_merge(dn, std, t, icrf, diff, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nb = N * H * W * C * 9 + H * W * C * 16
    print(f"C={C} W={W}: {ms:.3f} ms  {nb/ms/1e6:.0f} GB/s")
    del dn, std, out

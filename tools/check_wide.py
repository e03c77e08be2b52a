"""16-bit merge kernel (algo 3) against the generic kernel (algo 1) on a cfg5-like stack with dark frames and a flat
field, and its time on one full cfg5 stack.  Used with tools/ab_variants.sh.

    python tools/check_wide.py [channels] [rows] [std_table:0|1]
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def rel(a, b):
    d = (a - b).abs() / b.abs().clamp_min(1e-300)
    d = torch.where(a == b, torch.zeros_like(d), d)
    return float(d.max())


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    std_table = len(sys.argv) > 3 and sys.argv[3] == "1"
    W, N = 1000, 12
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    x16 = np.linspace(0, 1, 65536)
    icrf = torch.from_numpy(np.stack([x16 ** (2.0 + 0.1 * c) for c in range(C)], axis=1)).to(dev)
    diff = torch.from_numpy(np.stack([np.gradient(x16 ** (2.0 + 0.1 * c), 2 / 65535) for c in range(C)], axis=1)).to(dev)
    stdlut = torch.from_numpy(np.stack([0.002 + 0.02 * np.sqrt(x16) * (1 + 0.1 * c) for c in range(C)], axis=1)).to(dev)
    rad = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float32) * 25
    t = [0.0005 * 1.7 ** k for k in range(N)]
    dn = [torch.round(65535 * torch.clamp(rad * tk, 0, 1) ** (1 / 2.2)).to(torch.int32).to(torch.uint16) for tk in t]
    std = None if std_table else [torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.018 + 0.002 for _ in t]
    darks = []
    for k in range(N):
        if k < 6:
            darks.append(None)
            continue
        d = torch.randint(0, 600, (H, W, C), generator=g, device=dev, dtype=torch.int32)
        hot = torch.rand((H, W, C), generator=g, device=dev) < 2e-3
        darks.append(torch.where(hot, d + 12000, d).to(torch.uint16))
    flat = torch.randint(30000, 60000, (H, W, C), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
    flat_std = torch.rand((H, W, C), generator=g, device=dev, dtype=torch.float64) * 0.009 + 0.001
    means = torch.tensor([0.7 + 0.01 * c for c in range(C)] + [0.001] * C, dtype=torch.float64, device=dev)
    kw = dict(darks=darks, dark_threshold=0.05, flat=flat, flat_std=flat_std, flat_means=means)
    if std_table:
        kw["std_lut"] = stdlut
    res = {}
    for algo in (1, 3):
        v, s = ops.hdr_merge(dn, std, t, icrf, diff, algo=algo, **kw)
        res[algo] = (v.clone(), s.clone())
    v3b, s3b = ops.hdr_merge(dn, std, t, icrf, diff, algo=3, **kw)
    msg = (f"C={C} std_table={std_table}: val equal {torch.equal(res[1][0], res[3][0])} rel {rel(res[3][0], res[1][0]):.2e}; "
           f"std equal {torch.equal(res[1][1], res[3][1])} rel {rel(res[3][1], res[1][1]):.2e}; "
           f"repeat identical {torch.equal(v3b, res[3][0]) and torch.equal(s3b, res[3][1])}")
    del dn, std, darks, flat, flat_std, res
    # timing on one full cfg5 stack
    icrf5, diff5, stdlut5 = bench.cfg5_tables(dev)
    dn5, std5, t5 = bench.cfg5_stack_device(5000, dev, with_std=not std_table)
    shape = (bench.CFG5["H"], bench.CFG5["W"], 1)
    out = (torch.empty(shape, dtype=torch.float64, device=dev), torch.empty(shape, dtype=torch.float64, device=dev))
    kw5 = dict(std_lut=stdlut5) if std_table else {}
    reps = 5
    for _ in range(2):
        ops.hdr_merge(dn5, std5, t5, icrf5, diff5, out=out, algo=3, **kw5)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        ops.hdr_merge(dn5, std5, t5, icrf5, diff5, out=out, algo=3, **kw5)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
    print(msg, "| cfg5 stack ms (median of 5):", round(ms[reps // 2], 4))


if __name__ == "__main__":
    main()

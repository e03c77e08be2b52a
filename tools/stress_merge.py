"""Randomised stress of the HDR merge: staged (algo 2) vs generic (algo 1) kernels on random shapes,
exposure counts, dark-frame / flat-field combinations and hot-pixel densities (development tool).

    python tools/stress_merge.py [n_cases] [seed] [8|16]     (16: uint16 stacks, fused-table kernel vs generic)
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bits16 = len(sys.argv) > 3 and sys.argv[3] == "16"
    top, rows, fast = (65535, 65536, 3) if bits16 else (255, 256, 2)
    np_dt, t_view = (np.uint16, torch.uint16) if bits16 else (np.uint8, torch.uint8)

    def to_dev(a):
        return torch.from_numpy(a.view(np.int16)).to(dev).view(torch.uint16) if bits16 else torch.from_numpy(a).to(dev)
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    worst = 0.0
    for case in range(n_cases):
        C = int(rng.choice([1, 3]))
        n = int(rng.integers(1, 17))
        big = rng.uniform() < 0.1
        H = int(rng.integers(300, 1200)) if big else int(rng.integers(20, 200))
        W = int(rng.integers(300, 1600)) if big else int(rng.integers(30, 300))
        if H * W * C < 1536:
            continue
        x = np.linspace(0, 1, rows)
        icrf = torch.from_numpy(np.stack([x ** (1.8 + 0.2 * c) for c in range(C)], 1)).to(dev)
        diff = torch.from_numpy(np.stack([np.gradient(x ** (1.8 + 0.2 * c), 2 / top) for c in range(C)], 1)).to(dev)
        t = (0.001 * rng.uniform(1.3, 2.0) ** np.arange(n)).tolist()
        dn = [to_dev(rng.integers(0, top + 1, (H, W, C)).astype(np_dt)) for _ in range(n)]
        std = [torch.from_numpy(rng.uniform(0.002, 0.02, (H, W, C))).to(dev) for _ in range(n)]
        kw = {}
        if rng.uniform() < 0.7:
            hot_p = float(rng.choice([0.0, 0.001, 0.01, 0.1, 0.5]))
            darks = []
            for k in range(n):
                if rng.uniform() < 0.5:
                    d = rng.integers(0, 8, (H, W, C)).astype(np_dt)
                    d[rng.uniform(size=d.shape) < hot_p] = int(0.8 * top)
                    darks.append(to_dev(d))
                else:
                    darks.append(None)
            if any(d is not None for d in darks):
                kw.update(darks=darks, dark_threshold=0.05, median_kernel=int(rng.choice([3, 3, 5])))
        if rng.uniform() < 0.5:
            flat = to_dev(np.clip(np.rint(rng.normal(0.7 * top, 0.02 * top, (H, W, C))), 1, top).astype(np_dt))
            fstd = torch.from_numpy(rng.uniform(0.001, 0.01, (H, W, C))).to(dev)
            roi = (H // 4, 3 * H // 4, W // 4, 3 * W // 4)
            kw.update(flat=flat, flat_std=fstd, flat_means=ops.flat_roi_means(flat, fstd, roi, max_dn=float(top)))
        try:
            v2, s2 = ops.hdr_merge(dn, std, t, icrf, diff, algo=fast, **kw)
        except RuntimeError as exc:
            if "UNSUPPORTED" in str(exc):
                continue
            raise
        v1, s1 = ops.hdr_merge(dn, std, t, icrf, diff, algo=1, **kw)
        for a, b, what in ((v2, v1, "val"), (s2, s1, "std")):
            fin = torch.isfinite(b) & (b != 0)
            assert torch.equal(torch.isfinite(a), torch.isfinite(b)), (case, what)
            rel = float(((a[fin] - b[fin]).abs() / b[fin].abs()).max()) if bool(fin.any()) else 0.0
            worst = max(worst, rel)
            assert rel < 1e-13, (case, what, rel, (H, W, C, n), sorted(kw))
    print(f"{n_cases} cases ({'uint16, algo 3' if bits16 else 'uint8, algo 2'}) == generic kernel within {worst:.2e}")


if __name__ == "__main__":
    main()

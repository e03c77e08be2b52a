"""GPU diagnostic: staged vs generic HDR-merge kernels vs an independent torch float64 evaluation,
over a range of sizes (tiles per CTA) and exposure counts.  Prints where the worst element lives."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from camera_linearity_b200 import ops  # noqa: E402


def torch_reference(dn, std, t, icrf, diff):
    S = 0
    vals = []
    for d in dn:
        v = d.to(torch.float64) / 255
        vals.append(v)
        S = S + torch.exp(-30 * (v - 0.5) ** 2)
    hv = torch.zeros_like(vals[0])
    hs = torch.zeros_like(vals[0])
    ch = torch.arange(3, device=dn[0].device)
    for k, d in enumerate(dn):
        v = vals[k]
        w = torch.exp(-30 * (v - 0.5) ** 2)
        dw = -60 * (v - 0.5) * w
        g = icrf[d.long(), ch]
        dg = diff[d.long(), ch] * std[k]
        hv += (w * g) / (S * t[k])
        hs += (((dw * g + w * dg) / S - (dw * w * g) / S ** 2) * dg / t[k]) ** 2
    return hv, hs.sqrt()


def worst(a, b, name):
    rel = ((a - b).abs() / b.abs().clamp_min(1e-300))
    m = float(rel.max())
    idx = int(rel.argmax())
    px, c = divmod(idx, 3)
    tile, pos = divmod(px, 512)
    bad = int((rel > 1e-9).sum())
    print(f"   {name}: max rel {m:.3e} at sample {idx} (px {px}, ch {c}, tile {tile}, pos {pos}, "
          f"cta {tile % 148}, round {tile // 148}); elements > 1e-9: {bad}")
    if bad:
        bad_idx = torch.nonzero(rel.flatten() > 1e-9).flatten()
        tiles = torch.unique(bad_idx // (3 * 512))
        print(f"      bad tiles ({tiles.numel()}): {tiles[:20].tolist()} ... rounds {torch.unique(tiles // 148)[:10].tolist()}")
        first = int(bad_idx[0])
        print(f"      first bad sample {first}: got {float(a.flatten()[first])!r} want {float(b.flatten()[first])!r}")
    return m


def main():
    dev = torch.device("cuda")
    x = np.linspace(0, 1, 256)
    icrf = np.stack([x ** (2.0 + 0.1 * c) for c in range(3)], axis=1)
    diff = np.stack([np.gradient(icrf[:, c], 2 / 255) for c in range(3)], axis=1)
    icrf_t, diff_t = torch.from_numpy(icrf).to(dev), torch.from_numpy(diff).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for n in (5, 16):
        for tiles in (149, 1500, 16200, 16200):
            h, w = tiles, 512
            t = [0.001 * 1.6 ** k for k in range(n)]
            rad = torch.rand((h, w, 3), generator=g, device=dev, dtype=torch.float64) * 25
            dn = [torch.round(255 * torch.clamp(rad * tk, 0, 1) ** (1 / 2.2)).to(torch.uint8) for tk in t]
            std = [torch.rand((h, w, 3), generator=g, device=dev, dtype=torch.float64) * 0.018 + 0.002 for _ in t]
            rv, rs = torch_reference(dn, std, t, icrf_t, diff_t)
            print(f"N={n} tiles={tiles}")
            for algo in (1, 2):
                for rep in range(2):
                    v, s = ops.hdr_merge(dn, std, t, icrf_t, diff_t, algo=algo)
                    torch.cuda.synchronize()
                    worst(v, rv, f"algo{algo} rep{rep} val")
                    worst(s, rs, f"algo{algo} rep{rep} std")


if __name__ == "__main__":
    main()

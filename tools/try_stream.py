"""Development check of the single-pass merge kernel (algo 4) against the staged (2) and generic (1) kernels:
max relative differences and CUDA-event timings on cfg2 / cfg1, same process, same GPU."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def timed(fn, reps=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        fn()
        ev[r + 1].record()
    torch.cuda.synchronize()
    return statistics.median(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))


def rel(a, b):
    fin = torch.isfinite(b) & (b != 0)
    return float(((a[fin] - b[fin]).abs() / b[fin].abs()).max())


def main():
    dev = torch.device("cuda:0")
    for name in ("cfg2", "cfg1"):
        wl = bench.WORKLOADS[name]
        data = bench.make_stack_device(dict(wl, corrections=True), 1234, dev)
        icrf_np, diff_np = bench.icrf_tables(wl["C"])
        icrf, diff = torch.from_numpy(icrf_np).to(dev), torch.from_numpy(diff_np).to(dev)
        t = [float(x) for x in data["t"]]
        cl.GlobalSettings.configure(IM_SIZE_X=wl["H"], IM_SIZE_Y=wl["W"])
        means = ops.flat_roi_means(data["flat"], data["flat_std"], cl.measurand._flat_roi())
        kw = dict(darks=data["darks"], dark_threshold=bench.DARK_THRESHOLD, median_kernel=bench.KERNEL, flat=data["flat"],
                  flat_std=data["flat_std"], flat_means=means)
        for label, k in (("+corr", kw), (" plain", {})):
            v1, s1 = ops.hdr_merge(data["dn"], data["std"], t, icrf, diff, algo=1, **k)
            v4, s4 = ops.hdr_merge(data["dn"], data["std"], t, icrf, diff, algo=4, **k)
            v4b, s4b = ops.hdr_merge(data["dn"], data["std"], t, icrf, diff, algo=4, **k)
            same = torch.equal(v4, v4b) and torch.equal(s4, s4b)
            o = (torch.empty_like(v1), torch.empty_like(v1))
            t2 = timed(lambda: ops.hdr_merge(data["dn"], data["std"], t, icrf, diff, algo=2, out=o, **k))
            t4 = timed(lambda: ops.hdr_merge(data["dn"], data["std"], t, icrf, diff, algo=4, out=o, **k))
            print(f"{name}{label}: algo4 vs generic: val {rel(v4, v1):.2e} std {rel(s4, s1):.2e} repeat-identical {same} | "
                  f"staged {t2:.4f} ms, stream {t4:.4f} ms ({t2 / t4:.3f}x)", flush=True)
            del v1, s1, v4, s4, v4b, s4b, o
        del data


if __name__ == "__main__":
    main()

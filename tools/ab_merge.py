"""One process = one library build: median CUDA-event time of the fused merge on cfg2 (with / without corrections),
cfg1 and the cfg2 STD-table variant.  Used by tools/ab.sh to alternate two builds on ONE GPU.

    python tools/ab_merge.py [reps]
"""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def timed(fn, reps):
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


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    dev = torch.device("cuda:0")
    out = []
    for name in ("cfg2", "cfg1"):
        wl = bench.WORKLOADS[name]
        data = bench.make_stack_device(dict(wl, corrections=True), 1234, dev)
        icrf_np, diff_np = bench.icrf_tables(wl["C"])
        icrf, diff = torch.from_numpy(icrf_np).to(dev), torch.from_numpy(diff_np).to(dev)
        t = [float(x) for x in data["t"]]
        cl.GlobalSettings.configure(IM_SIZE_X=wl["H"], IM_SIZE_Y=wl["W"])
        roi = cl.measurand._flat_roi()
        means = ops.flat_roi_means(data["flat"], data["flat_std"], roi)
        o = (torch.empty((wl["H"], wl["W"], 3), dtype=torch.float64, device=dev),
             torch.empty((wl["H"], wl["W"], 3), dtype=torch.float64, device=dev))
        kw = dict(darks=data["darks"], dark_threshold=bench.DARK_THRESHOLD, median_kernel=bench.KERNEL, flat=data["flat"],
                  flat_std=data["flat_std"], flat_means=means)
        out.append(f"{name}+corr {timed(lambda: ops.hdr_merge(data['dn'], data['std'], t, icrf, diff, out=o, **kw), reps):.4f}")
        out.append(f"{name} plain {timed(lambda: ops.hdr_merge(data['dn'], data['std'], t, icrf, diff, out=o), reps):.4f}")
        if name == "cfg2":
            lut = torch.from_numpy(bench.std_table(3)).to(dev)
            try:
                out.append(f"{name}+corr STD-table "
                           f"{timed(lambda: ops.hdr_merge(data['dn'], None, t, icrf, diff, std_lut=lut, out=o, **kw), reps):.4f}")
            except Exception as exc:
                out.append(f"STD-table n/a ({type(exc).__name__})")
        del data, o
    print(" | ".join(out))


if __name__ == "__main__":
    main()

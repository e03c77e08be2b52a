"""Run ops.hdr_merge on the cfg2 bench stack a few times (profiling target for ncu).

    python tools/run_merge.py [dark_threshold] [reps] [darks:0|1] [flat:0|1] [lut|-] [algo]
    (lut: sigma from the STD table; algo: 0 auto, 1 generic, 2 staged two-pass, 4 single-pass)
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def main():
    thr = float(sys.argv[1]) if len(sys.argv) > 1 else bench.DARK_THRESHOLD
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    use_darks = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    use_flat = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
    use_lut = len(sys.argv) > 5 and sys.argv[5] == "lut"
    algo = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    dev = torch.device("cuda:0")
    wl = bench.WORKLOADS["cfg2"]
    data = bench.make_stack_device(wl, 1234, dev)
    icrf_np, diff_np = bench.icrf_tables(wl["C"])
    icrf, diff = torch.from_numpy(icrf_np).to(dev), torch.from_numpy(diff_np).to(dev)
    t = [float(x) for x in data["t"]]
    roi = cl.measurand._flat_roi()
    means = ops.flat_roi_means(data["flat"], data["flat_std"], roi)
    if not use_darks:
        data["darks"] = None
    if not use_flat:
        data["flat"] = data["flat_std"] = means = None
    std_lut = torch.from_numpy(bench.std_table(3)).to(dev) if use_lut else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        out = ops.hdr_merge(data["dn"], None if use_lut else data["std"], t, icrf, diff,
                            std_lut=std_lut,
                            darks=data["darks"], dark_threshold=thr, algo=algo,
                            median_kernel=bench.KERNEL, flat=data["flat"], flat_std=data["flat_std"],
                            flat_means=means)
        ev[r + 1].record()
    torch.cuda.synchronize()
    print("threshold", thr, "darks", use_darks, "flat", use_flat, "ms per call:", [round(ev[r].elapsed_time(ev[r + 1]), 4) for r in range(reps)])


if __name__ == "__main__":
    main()

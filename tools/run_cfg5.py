"""Run the 16-bit merge on ONE stack of BASELINE cfg5 as bench.py builds it (12 x 4320 x 7680 x 1 uint16); ncu target.

    python tools/run_cfg5.py [std_table:0|1] [reps] [algo]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402


def main():
    std_table = len(sys.argv) > 1 and sys.argv[1] == "1"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dev = torch.device("cuda:0")
    icrf, diff, stdlut = bench.cfg5_tables(dev)
    dn, std, t = bench.cfg5_stack_device(5000, dev, with_std=not std_table)
    shape = (bench.CFG5["H"], bench.CFG5["W"], 1)
    out = (torch.empty(shape, dtype=torch.float64, device=dev), torch.empty(shape, dtype=torch.float64, device=dev))
    kw = dict(std_lut=stdlut) if std_table else {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=algo, **kw)
        ev[r + 1].record()
    torch.cuda.synchronize()
    print("std_table", std_table, "algo", algo, "ms per call:", [round(ev[r].elapsed_time(ev[r + 1]), 3) for r in range(reps)])


if __name__ == "__main__":
    main()

"""Run the K4 objective on cfg3 a few times (profiling target for ncu).

    python tools/run_k4.py [std:0|1] [reps]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import camera_linearity_b200 as cl  # noqa: E402


def main():
    use_std = len(sys.argv) > 1 and sys.argv[1] == "1"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    mean, pca, tt, stack, std, params = bench.cfg3_problem()
    ev = cl.EnergyEvaluator(mean, pca, stack, std if use_std else None, 5, 250, True, tt, 64, shard=False)
    p_dev = torch.from_numpy(np.ascontiguousarray(params.T)).to(dev)
    for _ in range(3):
        ev.device_energies(p_dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ev.device_energies(p_dev)
    b.record()
    torch.cuda.synchronize()
    print("std", use_std, "ms per population:", a.elapsed_time(b) / reps)


if __name__ == "__main__":
    main()

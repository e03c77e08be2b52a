"""Timeline of the STD-table e2e step (events around H2D / kernels / D2H on the two alternating streams)."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import parallel  # noqa: E402

rank, world, local = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
wl = bench.WORKLOADS["cfg2"]
cl.GlobalSettings.configure(IM_SIZE_X=wl["H"], IM_SIZE_Y=wl["W"], DARK_THRESHOLD=0.05, MEDIAN_FILTER_KERNEL_SIZE=3,
                            FF_MID_PERCENTAGE=0.2)
data = bench.make_stack_device(wl, 1000 + rank, dev)
icrf_np, diff_np = bench.icrf_tables(3)
icrf, diff = torch.from_numpy(icrf_np).to(dev), torch.from_numpy(diff_np).to(dev)
t = [float(x) for x in data["t"]]
host_dn = [x.cpu().pin_memory() for x in data["dn"]]
feats = lambda tk, s: {"illumination": "bf", "magnification": "10x", "exposure": tk, "subject": s}
dark_sets = []
for k, d in enumerate(data["darks"]):
    if d is not None:
        ds = cl.ImageSet(features=feats(t[k], "dark")); ds.set_digital_numbers(d); dark_sets.append(ds)
fs = cl.ImageSet(features=feats(0.0, "flat"), measurand=cl.Measurand(None, data["flat_std"])); fs.set_digital_numbers(data["flat"])
lut = torch.from_numpy(bench.std_table(3)).to(dev)
data["dn"] = data["std"] = None
shape = (wl["H"], wl["W"], 3)
outs = [(torch.empty(shape, dtype=torch.float64).pin_memory(), torch.empty(shape, dtype=torch.float64).pin_memory()) for _ in range(2)]
streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
mode = sys.argv[1] if len(sys.argv) > 1 else "api"


def step(i, ev):
    st = streams[i % 2]
    with torch.cuda.stream(st):
        ev[0].record()
        sets = []
        for k in range(wl["N"]):
            s = cl.ImageSet(features=feats(t[k], "s"))
            s.set_digital_numbers(host_dn[k])
            sets.append(s)
        ev[1].record()
        series = cl.ExposureSeries(input_image_sets=sets)
        series.process_HDR_image(icrf, diff, dark_list=dark_sets, flat_list=[fs], STD_data=lut)
        ev[2].record()
        m = series.merged_image_set.measurand
        outs[i % 2][0].copy_(m.val, non_blocking=True)
        outs[i % 2][1].copy_(m.std, non_blocking=True)
        ev[3].record()


n = 8
events = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
for i in range(2):
    step(i, [torch.cuda.Event(enable_timing=True) for _ in range(4)])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
host_t = []
for i in range(n):
    step(i, events[i])
    host_t.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
base = events[0][0]
for i in range(n):
    print(f"rank {rank} step {i}: host enqueue done {host_t[i]:7.2f} | h2d {base.elapsed_time(events[i][0]):7.2f}-{base.elapsed_time(events[i][1]):7.2f}"
          f" | kernels -{base.elapsed_time(events[i][2]):7.2f} | d2h -{base.elapsed_time(events[i][3]):7.2f}", flush=True)
print(f"rank {rank}: {wall / n:.2f} ms per step", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

"""cfg5: does a merge cost more when consecutive calls read DIFFERENT stacks?  (bench's batch cycles over 8 stacks.)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
icrf, diff, stdlut = bench.cfg5_tables(dev)
n_stacks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
stacks = [bench.cfg5_stack_device(5000 + i, dev) for i in range(n_stacks)]
shape = (bench.CFG5["H"], bench.CFG5["W"], 1)
out = (torch.empty(shape, dtype=torch.float64, device=dev), torch.empty(shape, dtype=torch.float64, device=dev))


def run(order, label):
    for i in order[:2]:
        dn, std, t = stacks[i]
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(order) + 1)]
    ev[0].record()
    for r, i in enumerate(order):
        dn, std, t = stacks[i]
        ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = [round(ev[r].elapsed_time(ev[r + 1]), 3) for r in range(len(order))]
    print(label, ms, "total/len", round(ev[0].elapsed_time(ev[-1]) / len(order), 4))


run([0] * 8, "same stack      ")
run([i % n_stacks for i in range(8)], "cycling stacks  ")
run([1] * 8, "same stack (1)  ")
run([i % 2 for i in range(8)], "two alternating ")
for rep in range(3):
    run([0] * 8, f"rep{rep} same 0     ")
    run([i % 2 for i in range(8)], f"rep{rep} alternating")
    run([i % n_stacks for i in range(8)], f"rep{rep} cycling    ")
run([2, 3] * 4, "alternating 2,3 ")
run([0, 2] * 4, "alternating 0,2 ")

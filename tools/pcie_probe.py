"""Host<->device copy bandwidth from pinned memory, one process per GPU (torchrun), alone and concurrently."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 400 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=5):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


for kind in ("h2d", "d2h", "both"):
    run(kind, 2)
    bw = run(kind)
    print(f"rank {rank}/{world} {kind}: {bw:.1f} GB/s per direction", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""cfg5: per-merge time under sustained load, with NVML clocks / power / temperature sampled alongside.

    python tools/cfg5_sustained.py [merges] [std_table:0|1]
"""
import sys
import threading
import time
from pathlib import Path

import pynvml
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from camera_linearity_b200 import ops  # noqa: E402

merges = int(sys.argv[1]) if len(sys.argv) > 1 else 400
std_table = len(sys.argv) > 2 and sys.argv[2] == "1"
dev = torch.device("cuda:0")
icrf, diff, stdlut = bench.cfg5_tables(dev)
dn, std, t = bench.cfg5_stack_device(5000, dev, with_std=not std_table)
shape = (bench.CFG5["H"], bench.CFG5["W"], 1)
out = (torch.empty(shape, dtype=torch.float64, device=dev), torch.empty(shape, dtype=torch.float64, device=dev))
kw = dict(std_lut=stdlut) if std_table else {}
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        try:
            mem_t = pynvml.nvmlDeviceGetTemperature(h, 0)
        except Exception:
            mem_t = -1
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        mem_t, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)


for _ in range(3):
    ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3, **kw)
torch.cuda.synchronize()
time.sleep(1.0)
th = threading.Thread(target=sampler)
th.start()
time.sleep(0.1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(merges + 1)]
t0 = time.perf_counter()
ev[0].record()
host = []
for r in range(merges):
    a = time.perf_counter()
    ops.hdr_merge(dn, std, t, icrf, diff, out=out, algo=3, **kw)
    host.append(time.perf_counter() - a)
    ev[r + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
stop = True
th.join()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(merges)]
print("std_table", std_table, "host enqueue per call: median %.3f ms, max %.3f ms" % (sorted(host)[merges // 2] * 1e3, max(host) * 1e3))
for a in range(0, merges, merges // 10):
    chunk = ms[a:a + merges // 10]
    print("merges %4d-%4d: mean %.4f ms  min %.4f  max %.4f" % (a, a + len(chunk) - 1, sum(chunk) / len(chunk), min(chunk), max(chunk)))
print("NVML samples during the run (t [s], SM MHz, mem MHz, W, temp C, throttle reasons):")
sel = [s for s in samples if t0 - 0.1 <= s[0] <= t1 + 0.05]
for s in sel[::max(1, len(sel) // 16)]:
    print("  %.3f  %d  %d  %.0f  %d  0x%x" % (s[0] - t0, s[1], s[2], s[3], s[4], s[5]))

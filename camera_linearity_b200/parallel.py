"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the box,
gloo in the CPU tests).

K1-K3 shard by independent units (whole stacks, or row blocks of one stack) with no data-path
collective.  K4 (calibration loss) has one real exchange: the per-(candidate, exposure-pair)
numerator/denominator sums are all-reduced (SUM, float64, S x pairs x 2 values ~ 10 KB) once per
DE generation; the gates and the nanmean are applied after the reduction, identically on every
rank.  The energy itself is a mean of ratios and is NOT linear -- never all-reduce it.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


def rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*.
    Returns (rank, world_size, local_rank).  A single process needs no group."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not is_distributed():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rk, world_size=ws, **kwargs)
    return rk, ws, local


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin this process to the CPUs that are local to its GPU (NVML's ideal CPU affinity), so that pinned
    staging buffers are first-touched on the GPU's NUMA node and host->device copies do not cross sockets.
    Best effort: returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except Exception:
        return False


def shard_range(n_units: int, rank_: int | None = None, world: int | None = None) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of ``n_units`` independent units for this rank."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    base, extra = divmod(n_units, w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard_rows(height: int, halo: int = 0, rank_: int | None = None, world: int | None = None):
    """Row block [lo, hi) of an image for this rank plus the input rows [src_lo, src_hi) it must
    read: a ``halo``-row apron (K // 2 for the K x K bad-pixel median) clipped at the borders."""
    lo, hi = shard_range(height, rank_, world)
    return lo, hi, max(0, lo - halo), min(height, hi + halo)


def allreduce_pair_sums(pair_acc: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of the K4 pair sums (float64; counts are exact below 2^53)."""
    if world_size() > 1:
        dist.all_reduce(pair_acc, op=dist.ReduceOp.SUM)
    return pair_acc


class PeerExchange:
    """One peer-visible exchange buffer per rank for K4's fused tail kernel (``cl_icrf_energy_population``):
    every rank allocates its buffer (``cl_peer_alloc``: cudaMalloc + CUDA IPC handle), the 64-byte handles are
    all-gathered through the process group, and every rank maps the others' buffers on its own device
    (``cl_peer_open``).  The kernels then exchange the (S x pairs x 2) pair sums with plain stores over NVLink
    and a flag per rank -- torch.distributed only carried the handles."""

    def __init__(self, nbytes: int):
        import ctypes as C
        from . import _lib
        from ._lib import check
        self.lib = _lib.load()
        self.world, self.rank = world_size(), rank()
        if self.world > _lib.CL_MAX_PEERS:
            raise ValueError(f"at most {_lib.CL_MAX_PEERS} ranks")
        own = C.c_void_p()
        handle = _lib.IpcHandle()
        check(self.lib.cl_peer_alloc(int(nbytes), C.byref(own), C.byref(handle)), "cl_peer_alloc")
        self.own = own.value
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = self.own
        self._opened = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.bytes))
            for r, raw in enumerate(handles):
                if r == self.rank:
                    continue
                h = _lib.IpcHandle()
                C.memmove(C.byref(h), raw, 64)
                ptr = C.c_void_p()
                check(self.lib.cl_peer_open(C.byref(h), C.byref(ptr)), "cl_peer_open")
                self.ptrs[r] = ptr.value
                self._opened.append(ptr.value)
            dist.barrier()                 # every buffer is mapped everywhere before the first kernel pushes

    def peer_group(self):
        from . import _lib
        pg = _lib.PeerGroup()
        pg.world, pg.rank = self.world, self.rank
        for r, p in enumerate(self.ptrs):
            pg.buffers[r] = p
        return pg

    def close(self):
        if self.own is None:
            return
        if self.world > 1 and is_distributed():
            torch.cuda.synchronize()
            dist.barrier()                 # nobody unmaps while a peer's kernel may still push
        for p in self._opened:
            self.lib.cl_peer_close(p)
        self._opened = []
        self.lib.cl_peer_free(self.own)
        self.own = None

    def __del__(self):
        try:
            if self.own is not None:
                for p in self._opened:
                    self.lib.cl_peer_close(p)
                self.lib.cl_peer_free(self.own)
                self.own = None
        except Exception:
            pass

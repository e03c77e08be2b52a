"""Multi-GPU check of the K4 exchange step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_energy_check.py

Every rank holds a pixel shard, evaluates the whole population on it, the per-(candidate, pair) sums
are all-reduced over NCCL, and every rank finalises.  Rank 0 compares with the NumPy oracle on the
full data and with the single-GPU (unsharded) evaluation, and times a generation.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import parallel  # noqa: E402
from oracle import icrf_energy as oe  # noqa: E402


def main():
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cl.GlobalSettings.DEVICE = str(dev)
    rng = np.random.default_rng(3)                    # identical data on every rank
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    pca, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.005 * 2.0 ** np.arange(5)
    rad = rng.uniform(0, 1, (1000, 400, 1)) * 25
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** (1 / 2.2)).astype(np.uint8)
    std = rng.uniform(0.002, 0.02, dn.shape)
    params = rng.uniform(-0.05, 0.05, (5, 64))
    out = {"world": world}
    for name, sd in (("nostd", None), ("std", std)):
        sharded = cl.EnergyEvaluator(mean, pca, dn, sd, 5, 250, True, t, 64, shard=True)
        e = sharded(params)
        whole = cl.EnergyEvaluator(mean, pca, dn, sd, 5, 250, True, t, 64, shard=False)(params)
        torch.cuda.synchronize()
        torch.distributed.barrier() if world > 1 else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            sharded(params)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        if rank == 0:
            ref = oe.energy_population(params[:, :4], mean, pca, dn, sd, 5, 250, True, t)
            def rel(a_, b_):
                assert np.array_equal(np.isinf(a_), np.isinf(b_))          # identical gate decisions
                fin = np.isfinite(b_)
                return float(np.max(np.abs(a_[fin] - b_[fin]) / np.abs(b_[fin]))) if fin.any() else 0.0
            out[name] = {"max_rel_vs_single_gpu": rel(e, whole), "max_rel_vs_oracle_first4": rel(e[:4], ref),
                         "finite_candidates": int(np.isfinite(e).sum()),
                         "ms_per_generation_incl_host": ms, "evals_per_s": 64 / ms * 1e3}
    # device-resident DE over the sharded objective: one NCCL all-reduce per generation, no host sync;
    # every rank draws the same counter-based random numbers, so the populations stay identical
    from scipy.stats import qmc
    from camera_linearity_b200 import ops
    unit = torch.from_numpy(qmc.Sobol(d=5, seed=np.random.default_rng(7)).random(n=64)).to(dev)
    ev_sh = cl.EnergyEvaluator(mean, pca, dn, None, 5, 250, True, t, 64, shard=True)
    de = ops.DeviceDE(ev_sh.device_energies, [-0.5] * 5, [0.5] * 5, unit, seed=7, tol=0.0)
    for _ in range(5):
        de.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        de.step()
    b.record()
    torch.cuda.synchronize()
    pop = de.pop.clone()
    if world > 1:
        gathered = [torch.empty_like(pop) for _ in range(world)]
        torch.distributed.all_gather(gathered, pop)
        same = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    else:
        same = True
    ev_1 = cl.EnergyEvaluator(mean, pca, dn, None, 5, 250, True, t, 64, shard=False)
    de1 = ops.DeviceDE(ev_1.device_energies, [-0.5] * 5, [0.5] * 5, unit, seed=7, tol=0.0)
    for _ in range(45):
        de1.step()
    _, gens, best = de.poll()
    _, _, best1 = de1.poll()
    out["device_de"] = {"ms_per_generation": a.elapsed_time(b) / 40, "generations": gens,
                        "populations_identical_across_ranks": same, "best_energy_sharded": best,
                        "best_energy_single_gpu": best1}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))
        assert same
        assert out["nostd"]["max_rel_vs_oracle_first4"] < 1e-9 and out["std"]["max_rel_vs_oracle_first4"] < 1e-9


if __name__ == "__main__":
    main()

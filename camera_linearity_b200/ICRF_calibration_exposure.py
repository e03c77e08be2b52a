"""ICRF calibration by differential evolution over PCA-basis curves
(reference: ``modules/ICRF_calibration_exposure.py``).

``_energy_function`` keeps the reference's signature and returns a python float for one candidate;
``EnergyEvaluator`` scores a whole DE population in one launch (SciPy ``vectorized=True`` passes
``(n_params, S)`` and expects ``(S,)``).  With ``torch.distributed`` initialised, pixels are
sharded across ranks and the per-(candidate, pair) sums are all-reduced once per generation
(``parallel.allreduce_pair_sums``) before the identical finalize on every rank.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import general_functions as gf
from . import ops
from . import parallel
from .image_set import ImageSet
from .settings import GlobalSettings as gs


def _inverse_camera_response_function(mean_ICRF, PCA_array, PCA_params, use_mean_ICRF):
    """Host restatement used for the final curve only (ICRF_calibration_exposure.py:20-44)."""
    p = np.asarray(PCA_params, dtype=np.float64)
    pca = np.asarray(PCA_array, dtype=np.float64)
    if not use_mean_ICRF:
        return np.linspace(0, 1, gs.BITS) ** p[0] + np.matmul(pca, p[1:])
    return np.asarray(mean_ICRF, dtype=np.float64) + np.matmul(pca, p)


class EnergyEvaluator:
    """Population objective for one colour channel, resident on this rank's GPU."""

    def __init__(self, mean_ICRF, PCA_array, image_value_stack, image_std_stack, lower, upper, use_mean,
                 exposure_values, n_candidates: int, shard: bool = True, exchange: str = "auto"):
        """shard: split the pixels over the ranks of the process group.  exchange: how the ranks' pair sums meet --
        "peer" (pushed through peer memory inside the fused tail kernel), "nccl" (all-reduce between the partial and
        finalize launches) or "auto" (peer when every rank has its own GPU, else nccl)."""
        dev = gs.device()
        dn = torch.as_tensor(image_value_stack)
        if dn.dtype.is_floating_point:
            raise IndexError("arrays used as indices must be of integer (or boolean) type")
        if dn.ndim != 3:
            raise ValueError("image_stack must be a 3D CuPy array with shape (X, Y, N).")
        sd = None if image_std_stack is None else torch.as_tensor(image_std_stack)
        if shard and parallel.world_size() > 1:
            # pixel rows are independent samples: contiguous shard per rank, no halo
            n_exp = dn.shape[2]
            flat = dn.reshape(-1, n_exp)
            lo, hi = parallel.shard_range(flat.shape[0])
            dn = flat[lo:hi].reshape(-1, 1, n_exp)
            if sd is not None:
                sd = sd.reshape(-1, n_exp)[lo:hi].reshape(-1, 1, n_exp)
        self.plan = ops.IcrfEnergyPlan(dn.to(dev), None if sd is None else sd.to(dev), exposure_values,
                                       mean_ICRF, torch.as_tensor(np.asarray(PCA_array), device=dev),
                                       int(lower), int(upper), bool(use_mean), int(n_candidates))
        self.sharded = shard and parallel.world_size() > 1
        self.exchange = "none"
        if self.sharded:
            if exchange == "auto":
                exchange = "peer" if torch.cuda.device_count() >= parallel.world_size() else "nccl"
            if exchange not in ("peer", "nccl"):
                raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
            self.exchange = exchange
            if exchange == "peer":
                nbytes = self.plan.lib.cl_icrf_exchange_bytes(self.plan.prob, parallel.world_size())
                self._peers = parallel.PeerExchange(nbytes)
                self.plan.attach_peers(self._peers)

    def __call__(self, params):
        """params: (n_params,) or (n_params, S) NumPy array -> float or (S,) NumPy array."""
        p = np.asarray(params, dtype=np.float64)
        single = p.ndim == 1
        pop = p.reshape(p.shape[0], -1).T          # (S, n_params)
        plan = self.plan
        if pop.shape[0] != plan.n_real:
            raise ValueError(f"evaluator was built for {plan.n_real} candidates, got {pop.shape[0]}")
        e = self.device_energies(torch.from_numpy(np.ascontiguousarray(pop))).cpu().numpy()
        return float(e[0]) if single else e

    def device_energies(self, params: torch.Tensor) -> torch.Tensor:
        """params: (S, n_params) device tensor -> (S,) device tensor; nothing synchronises with the host
        (the objective of the device-resident DE driver)."""
        plan = self.plan
        plan.set_params(params)
        plan.curves_and_tables()
        if self.exchange == "nccl":
            parallel.allreduce_pair_sums(plan.partial())
            return plan.finalize()
        return plan.population()          # single rank, or peer-memory exchange inside the tail kernel


def _energy_function(PCA_params, mean_ICRF, PCA_array, image_value_stack, image_std_stack, lower, upper,
                     use_mean, exposure_values):
    """One candidate (ICRF_calibration_exposure.py:148-201).  Builds a throw-away evaluator; use
    ``EnergyEvaluator`` to amortise the upload over a whole optimisation."""
    if not isinstance(PCA_params, np.ndarray):
        raise TypeError("PCA_params must be a NumPy array")      # D17: the reference leaves `temp` unbound
    ev = EnergyEvaluator(mean_ICRF, PCA_array, image_value_stack, image_std_stack, lower, upper, use_mean,
                         np.asarray(exposure_values), 1, shard=False)
    return ev(PCA_params)


def interpolate_ICRF(ICRF_array):
    """ICRF_calibration_exposure.py:204-216."""
    if gs.BITS == gs.DATAPOINTS:
        return ICRF_array
    x_new = np.linspace(0, 1, num=gs.BITS)
    x_old = np.linspace(0, 1, num=gs.DATAPOINTS)
    out = np.zeros((gs.BITS, gs.NUM_OF_CHS), dtype=float)
    for c in range(gs.NUM_OF_CHS):
        out[:, c] = np.interp(x_new, x_old, ICRF_array[:, c])
    return out


def initialize_channel_image_stacks(image_path: Path, use_std: bool, data_spacing):
    """Per-channel (X, Y, N) stacks of strided pixel samples (ICRF_calibration_exposure.py:219-285)."""
    x_step, y_step = data_spacing if type(data_spacing) is tuple else (data_spacing, data_spacing)
    sets = ImageSet.multiple_from_path(image_path)
    sets.sort(key=lambda s: s.features["exposure"])
    values, stds, exposures = [], [], []
    for image_set in sets:
        exposures.append(image_set.features['exposure'])
        image_set.load_value_image(bit64=True)
        v = gf.choose_evenly_spaced_points(image_set.measurand.val, x_step, y_step)
        values.append(v)
        if use_std:
            image_set.load_std_image()
            stds.append(gf.choose_evenly_spaced_points(image_set.measurand.std, x_step, y_step))
        image_set.measurand.val = None
        image_set.measurand.std = None
    channels = values[0].shape[2]
    value_stacks = [torch.stack([v[:, :, c] for v in values], dim=2) for c in range(channels)]
    std_stacks = ([torch.stack([s[:, :, c] for s in stds], dim=2) for c in range(channels)]
                  if use_std else [None] * channels)
    return value_stacks, std_stacks, np.array(exposures)


def _solve_channel_device(mean_ICRF, PCA_array, image_value_stack, image_std_stack, exposure_values, limits, x0,
                          data_limits, use_mean_ICRF, seed, energy_limit, max_iterations, popsize, check_every=8):
    """The same solve with the DE generation on the GPU (``ops.DeviceDE``): trial vectors, the population
    objective (K4) and the selection are enqueued back to back; the host looks at the status block every
    ``check_every`` generations.  Initial population as scipy's ``init='sobol'`` (2**ceil(log2(popsize*P))
    members, ``x0`` in row 0); random draws are counter based, so the search is statistically -- not bitwise --
    scipy's."""
    from scipy.stats import qmc
    limits = np.asarray(limits, dtype=np.float64)
    lower, upper = limits[:, 0], limits[:, 1]
    n_params = limits.shape[0]
    members = int(2 ** np.ceil(np.log2(popsize * n_params)))
    unit = qmc.Sobol(d=n_params, seed=np.random.default_rng(seed)).random(n=members)
    arg1, arg2 = 0.5 * (lower + upper), np.fabs(upper - lower)
    unit[0] = (np.asarray(x0, dtype=np.float64) - arg1) / arg2 + 0.5
    evaluator = EnergyEvaluator(mean_ICRF, PCA_array, image_value_stack, image_std_stack, data_limits[0],
                                data_limits[1], use_mean_ICRF, exposure_values, members)
    # single rank / peer-memory exchange: the fused generation (4 launches) replayed from a CUDA graph;
    # NCCL exchange: the generic step with the all-reduce between the partial and finalize launches
    fused = evaluator.exchange != "nccl"
    if fused:
        de = ops.DeviceDE.for_plan(evaluator.plan, lower, upper, torch.from_numpy(unit).to(gs.device()), seed)
    else:
        de = ops.DeviceDE(evaluator.device_energies, lower, upper, torch.from_numpy(unit).to(gs.device()), seed)
    iterations = 0
    energy = float("inf")
    while iterations < max_iterations:
        burst = min(check_every, max_iterations - iterations)
        if fused and burst == check_every:
            de.run_graph(burst, per_graph=check_every)
        else:
            for _ in range(burst):
                de.step_fused() if fused else de.step()
        converged, iterations, energy = de.poll()
        if converged or energy < energy_limit:
            break
    result = de.x.cpu().numpy()
    return _inverse_camera_response_function(mean_ICRF, PCA_array, result, use_mean_ICRF), result, energy, iterations


def solve_channel(mean_ICRF, PCA_array, image_value_stack, image_std_stack, exposure_values, limits, x0,
                  data_limits=(5, 250), use_mean_ICRF=True, seed=7, energy_limit=0.0, max_iterations=1000,
                  popsize=12, driver="scipy"):
    """One channel's differential-evolution solve (ICRF_calibration_exposure.py:341-378).

    ``driver='device'`` keeps the whole generation on the GPU (``_solve_channel_device``).

    SciPy's solver runs on the host with the reference's settings (currenttobest1bin, tol 0.01,
    mutation (0, 1.95), recombination 0.4, Sobol init); ``vectorized=True, updating='deferred'``
    hands the whole population to one GPU launch per generation.  ``rng=`` replaces the removed
    ``seed=`` keyword (D15); the stop rule (converged / max_iterations / energy below limit) is the
    reference's, one generation per loop step.
    """
    if driver == "device":
        return _solve_channel_device(mean_ICRF, PCA_array, image_value_stack, image_std_stack, exposure_values,
                                     limits, x0, data_limits, use_mean_ICRF, seed, energy_limit, max_iterations,
                                     popsize)
    if driver != "scipy":
        raise ValueError("driver must be 'scipy' or 'device'")
    from scipy.optimize._differentialevolution import DifferentialEvolutionSolver
    evaluator = None

    def objective(pop):
        nonlocal evaluator
        pop = np.asarray(pop, dtype=np.float64)
        n = 1 if pop.ndim == 1 else pop.shape[1]
        if evaluator is None or evaluator.plan.n_real != n:
            evaluator = EnergyEvaluator(mean_ICRF, PCA_array, image_value_stack, image_std_stack, data_limits[0],
                                        data_limits[1], use_mean_ICRF, exposure_values, n)
        return evaluator(pop)

    iterations = 0
    with DifferentialEvolutionSolver(objective, limits, strategy='currenttobest1bin', tol=0.01, x0=x0,
                                     mutation=(0, 1.95), recombination=0.4, init='sobol', rng=seed,
                                     popsize=popsize, vectorized=True, updating='deferred', polish=False) as solver:
        for step in solver:
            iterations += 1
            if solver.converged() or iterations == max_iterations or step[1] < energy_limit:
                break
        result, energy = solver.x, float(np.min(solver.population_energies))
    return _inverse_camera_response_function(mean_ICRF, PCA_array, result, use_mean_ICRF), result, energy, iterations


def calibration(lower_PCA_limit: float, upper_PCA_limit: float, initial_function=None, data_spacing=150,
                data_limits=(5, 250), use_std: Optional[bool] = False, image_path: Optional[Path] = None,
                energy_limit: Optional[float] = 0, rng_seed: Optional[int] = 7, use_cupy: Optional[bool] = False,
                max_iterations: int = 1000, driver: str = "scipy"):
    """Driver with the reference's signature and return tuple (ICRF_calibration_exposure.py:288-402).
    Channels are solved one after another on the GPU (the reference forks one process per channel).
    ``driver='device'`` runs the differential evolution itself on the GPU (see ``solve_channel``)."""
    image_path = gs.DEFAULT_IMG_SRC_PATH if image_path is None else image_path
    use_mean_ICRF = initial_function is None
    limits, x0 = [], []
    if not use_mean_ICRF:
        limits.append([1, 8])
        x0.append(3)
    for _ in range(gs.NUM_OF_PCA_PARAMS):
        limits.append([lower_PCA_limit, upper_PCA_limit])
        x0.append(0)
    value_stacks, std_stacks, exposure_values = initialize_channel_image_stacks(image_path, use_std, data_spacing)

    ICRF = np.zeros((gs.DATAPOINTS, gs.NUM_OF_CHS), dtype=float)
    final_energy_array = np.zeros(gs.NUM_OF_CHS, dtype=float)
    initial_energy_array = np.zeros(gs.NUM_OF_CHS, dtype=float)
    for c in range(gs.NUM_OF_CHS):
        pca = gf.read_txt_to_array(gs.PCA_FILES[c])
        mean = gf.read_txt_to_array(gs.MEAN_ICRF_FILES[c]) if use_mean_ICRF else initial_function
        curve, _, energy, _ = solve_channel(mean, pca, value_stacks[c], std_stacks[c], exposure_values, limits, x0,
                                            data_limits, use_mean_ICRF, rng_seed + c, energy_limit, max_iterations,
                                            driver=driver)
        ICRF[:, c] = curve
        ICRF[:, c] += 1 - ICRF[-1, c]
        ICRF[0, c] = 0
        final_energy_array[c] = energy
    ICRF[ICRF < 0] = 0
    ICRF[ICRF > 1] = 1
    return interpolate_ICRF(ICRF), initial_energy_array, final_energy_array, 0

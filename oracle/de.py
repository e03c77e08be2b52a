"""Oracle for the device-resident differential-evolution step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Spec: ``scipy.optimize._differentialevolution.DifferentialEvolutionSolver`` as the reference configures it
(``modules/ICRF_calibration_exposure.py:357-361``), vectorized / ``updating='deferred'``:
``_mutate_many`` + ``_currenttobest1`` (trial vectors), ``_ensure_constraint`` (bound repair),
``_scale_parameters``, the deferred selection ``trial_energies < population_energies``,
``_promote_lowest_energy`` and ``converged``.  SciPy draws from a NumPy ``Generator``; the device path
and this restatement use the same counter-based splitmix64 draws instead (keyed by seed, generation,
member and slot), so they agree bit for bit with each other while the search is statistically -- not
bitwise -- the one SciPy would run.  "Parity unpinned" against SciPy's own random stream by construction;
``tests/test_gpu_de.py`` compares the optimum both drivers reach.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
SLOT_R0, SLOT_R1, SLOT_FILL, SLOT_CROSS = 0, 1, 2, 3
SCALE_CANDIDATE = 0xFFFFFFFF


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def draw(seed: int, gen: int, member, slot):
    """Uniform [0, 1) double number ``slot`` of ``member`` in generation ``gen`` (53 random bits)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) ^ (np.uint64(gen) * np.uint64(0xD1342543DE82EF95)))
        ctr = (np.asarray(member, dtype=np.uint64) << np.uint64(8)) | np.asarray(slot, dtype=np.uint64)
        z = _splitmix64(key + ctr)
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def trial_population(pop, seed, gen, dither, crossover, lower, upper):
    """One generation's trial vectors (unit cube) and their scaled parameters."""
    pop = np.asarray(pop, dtype=np.float64)
    S, P = pop.shape
    i = np.arange(S)
    scale = dither[0] + (dither[1] - dither[0]) * draw(seed, gen, SCALE_CANDIDATE, 0)
    r0 = (draw(seed, gen, i, SLOT_R0) * (S - 1)).astype(np.int64)
    r0 += r0 >= i
    r1 = (draw(seed, gen, i, SLOT_R1) * (S - 2)).astype(np.int64)
    a, b = np.minimum(i, r0), np.maximum(i, r0)
    r1 += r1 >= a
    r1 += r1 >= b
    fill = (draw(seed, gen, i, SLOT_FILL) * P).astype(np.int64)
    j = np.arange(P)
    cross = draw(seed, gen, i[:, None], SLOT_CROSS + j[None, :]) < crossover
    cross[i, fill] = True
    bprime = pop + scale * (((pop[0][None, :] - pop) + pop[r0]) - pop[r1])          # _currenttobest1
    trial = np.where(cross, bprime, pop)
    oob = (trial > 1) | (trial < 0)                                                # _ensure_constraint
    trial = np.where(oob, draw(seed, gen, i[:, None], SLOT_CROSS + P + j[None, :]), trial)
    lower, upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
    params = 0.5 * (lower + upper) + (trial - 0.5) * np.fabs(upper - lower)        # _scale_parameters
    return trial, params


def select(pop, energies, trial, trial_energies, tol=0.01, atol=0.0):
    """Deferred selection, promotion of the best member to row 0 and scipy's convergence test."""
    pop, energies = np.array(pop, dtype=np.float64), np.array(energies, dtype=np.float64)
    loc = np.asarray(trial_energies) <= energies          # scipy _accept_trial (energy_trial <= energy_orig)
    pop = np.where(loc[:, None], trial, pop)
    energies = np.where(loc, trial_energies, energies)
    best = int(np.argmin(energies))
    converged = False
    if not np.any(np.isinf(energies)):
        converged = bool(np.std(energies) <= atol + tol * np.abs(np.mean(energies)))
    pop[[0, best]] = pop[[best, 0]]
    energies[[0, best]] = energies[[best, 0]]
    return pop, energies, dict(converged=converged, replaced=int(loc.sum()), best_index=best)

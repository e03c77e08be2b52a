"""Device-resident differential evolution (csrc/de.cu) vs its NumPy restatement (oracle/de.py, same
counter-based draws -> bit-exact) and vs the scipy driver on a calibration problem with a known answer."""
import numpy as np
import pytest
import torch

from oracle import de as ode
from oracle import icrf_energy as oe
from gpu_util import host

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import ICRF_calibration_exposure as cal  # noqa: E402


def _sphere(target):
    t = torch.as_tensor(target, dtype=torch.float64, device="cuda")
    return lambda params: torch.sum((params - t) ** 2, dim=1)


@pytest.mark.parametrize("S,P", [(64, 5), (8, 1), (32, 6), (5, 3)])
def test_generation_matches_oracle_bitexact(S, P):
    rng = np.random.default_rng(S * 10 + P)
    lower, upper = -rng.uniform(0.5, 2, P), rng.uniform(0.5, 2, P)
    unit = rng.uniform(0, 1, (S, P))
    evaluate = _sphere(rng.uniform(-0.3, 0.3, P))
    de = ops.DeviceDE(evaluate, lower, upper, torch.from_numpy(unit).cuda(), seed=1234)
    for gen in range(6):
        pop0, e0 = host(de.pop).copy(), host(de.energies).copy()
        de.step()
        trial, params = ode.trial_population(pop0, 1234, gen, (0.0, 1.95), 0.4, lower, upper)
        assert np.array_equal(host(de.trial), trial)
        assert np.array_equal(host(de.params), params)
        assert trial.min() >= 0 and trial.max() <= 1
        te = host(evaluate(de.params))                       # the energies the device selection saw
        pop1, e1, st = ode.select(pop0, e0, trial, te)
        assert np.array_equal(host(de.pop), pop1) and np.array_equal(host(de.energies), e1)
        status = host(de.status)
        assert (bool(status[0]), int(status[1]), int(status[2]), int(status[3])) == (
            st["converged"], gen + 1, st["replaced"], st["best_index"])
        assert float(de.best[0].cpu()) == e1.min() == e1[0]


def test_sphere_converges():
    P = 5
    target = np.array([0.11, -0.07, 0.02, 0.3, -0.25])
    from scipy.stats import qmc
    unit = qmc.Sobol(d=P, seed=np.random.default_rng(3)).random(n=64)
    de = ops.DeviceDE(_sphere(target), [-1] * P, [1] * P, torch.from_numpy(unit).cuda(), seed=3, tol=1e-6)
    for _ in range(400):
        de.step()
    converged, gens, best = de.poll()
    assert gens == 400 and best < 1e-6
    assert np.allclose(host(de.x), target, atol=2e-3)


def _calibration_problem():
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    pca = np.stack([0.1 * np.sin((k + 1) * np.pi * x) for k in range(5)], axis=1)
    p_true = np.array([0.12, -0.05, 0.03, 0.0, 0.01])
    curve = mean + pca @ p_true
    curve = curve + (1 - curve[-1])
    curve[0] = 0
    assert np.all(np.diff(curve) > 0)
    rng = np.random.default_rng(11)
    t = np.array([.005, .01, .02, .04, .08])
    rad = rng.uniform(0.5, 14, (64, 48, 1))
    target = np.clip(rad * t[None, None, :], 0, 1)
    dn = np.abs(curve[None, None, None, :] - target[..., None]).argmin(axis=-1).astype(np.uint8)   # the camera's CRF
    return mean, pca, p_true, curve, dn, t


def test_device_driver_reaches_the_scipy_optimum():
    mean, pca, p_true, curve, dn, t = _calibration_problem()
    limits = [[-0.5, 0.5]] * 5
    x0 = [0.0] * 5
    e_true = oe.energy(p_true, mean, pca, dn, None, 5, 250, True, t)
    res = {}
    for driver in ("scipy", "device"):
        c, p, e, it = cal.solve_channel(mean, pca, dn, None, t, limits, x0, seed=5, max_iterations=400, driver=driver)
        c = c + (1 - c[-1])
        res[driver] = (c, p, e, it)
        assert e <= e_true * 1.05 + 1e-12, (driver, e, e_true)
        assert np.max(np.abs(c - curve)) < 5e-3, driver
    # same optimum from both drivers (the searches differ: scipy's Generator vs counter-based draws)
    assert abs(res["device"][2] - res["scipy"][2]) <= 0.05 * e_true + 1e-12
    assert res["device"][3] <= 400


def test_fused_generation_and_cuda_graph_match_the_unfused_steps():
    """cl_de_trial_curves + cl_icrf_energy_population + cl_de_select (4 launches, replayed from a CUDA graph) must
    walk exactly the path of the unfused generation (cl_de_trial, cl_icrf_curves, partial, finalize, select)."""
    from scipy.stats import qmc
    mean, pca, p_true, curve, dn, t = _calibration_problem()
    unit = torch.from_numpy(qmc.Sobol(d=5, seed=np.random.default_rng(9)).random(n=64)).cuda()
    ev_a = cl.EnergyEvaluator(mean, pca, dn, None, 5, 250, True, t, 64, shard=False)
    de_a = ops.DeviceDE(ev_a.device_energies, [-0.5] * 5, [0.5] * 5, unit, seed=9, tol=0.0)
    ev_b = cl.EnergyEvaluator(mean, pca, dn, None, 5, 250, True, t, 64, shard=False)
    de_b = ops.DeviceDE.for_plan(ev_b.plan, [-0.5] * 5, [0.5] * 5, unit, seed=9, tol=0.0)
    assert torch.equal(de_a.pop, de_b.pop) and torch.equal(de_a.energies, de_b.energies)
    for _ in range(24):
        de_a.step()
    de_b.run_graph(24, per_graph=8)
    torch.cuda.synchronize()
    assert torch.equal(de_a.pop, de_b.pop) and torch.equal(de_a.energies, de_b.energies)
    assert de_a.poll()[1] == de_b.poll()[1] == 24
    # and with uncertainty stacks, S not a multiple of 32 (padding candidates)
    rng = np.random.default_rng(4)
    sd = rng.uniform(0.002, 0.02, dn.shape)
    unit40 = unit[:40].clone()
    ev_c = cl.EnergyEvaluator(mean, pca, dn, sd, 5, 250, True, t, 40, shard=False)
    de_c = ops.DeviceDE(ev_c.device_energies, [-0.5] * 5, [0.5] * 5, unit40, seed=3, tol=0.0)
    ev_d = cl.EnergyEvaluator(mean, pca, dn, sd, 5, 250, True, t, 40, shard=False)
    de_d = ops.DeviceDE.for_plan(ev_d.plan, [-0.5] * 5, [0.5] * 5, unit40, seed=3, tol=0.0)
    for _ in range(9):
        de_c.step()
        de_d.step_fused()
    assert torch.equal(de_c.pop, de_d.pop) and torch.equal(de_c.energies, de_d.energies)

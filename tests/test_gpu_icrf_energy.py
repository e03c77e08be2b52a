"""K4 parity: calibration objective for a DE population vs the unmodified reference (goldens) and
the oracle.  Energies <= 1e-6 relative (actual ~1e-13); gate decisions (+inf) and valid-sample counts
exact."""
import numpy as np
import pytest
import torch

from oracle import icrf_energy as oe
from gpu_util import assert_rel, host

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402

TIGHT = 1e-11


def test_golden_reference_energies(golden_dir):
    g = np.load(golden_dir / "k4_energy.npz")
    s = g["params"].shape[1]
    ev = cl.EnergyEvaluator(g["mean"], g["pca"], g["dn"], None, 5, 250, True, g["t"], s)
    assert_rel(ev(g["params"]), g["e_nostd"], TIGHT)
    ev = cl.EnergyEvaluator(g["mean"], g["pca"], g["dn"], g["std"], 5, 250, True, g["t"], s)
    assert_rel(ev(g["params"]), g["e_std"], TIGHT)
    ev = cl.EnergyEvaluator(g["mean"], g["pca"], g["dn"], None, 5, 250, False, g["t"], g["params6"].shape[1])
    assert_rel(ev(g["params6"]), g["e6"], 1e-9)          # pow() on the device vs NumPy: a few ulp
    # single-candidate signature of the reference
    e = cl._energy_function(g["params"][:, 0].copy(), g["mean"], g["pca"], g["dn"], g["std"], 5, 250, True, g["t"])
    assert isinstance(e, float) and abs(e - g["e_std"][0]) <= 1e-11 * abs(g["e_std"][0])


def test_pair_sums_and_counts_exact(golden_dir):
    g = np.load(golden_dir / "k4_energy.npz")
    ev = cl.EnergyEvaluator(g["mean"], g["pca"], g["dn"], None, 5, 250, True, g["t"], g["params"].shape[1])
    ev(g["params"])
    acc = host(ev.plan.pair_acc)
    # candidate 0: compare per-pair mean with analyze_linearity and the valid count with a NumPy count
    curve = g["mean"] + g["pca"] @ g["params"][:, 0]
    curve += 1 - curve[-1]
    curve[0] = 0
    mapped = curve[g["dn"]]
    ok = ~((mapped < curve[5]) | (mapped > curve[250]))
    q = 0
    for i in range(5):
        for j in range(i + 1, 5):
            count = int(np.sum(ok[..., i] & ok[..., j]))
            assert acc[0, q, 1] == count                                  # integer: exact
            if count:
                assert abs(acc[0, q, 0] / acc[0, q, 1] - g["pairs_nostd"][q]) <= 1e-11 * g["pairs_nostd"][q]
            q += 1


def test_survey_kat_e():
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    pca = np.stack([0.1 * np.sin((k + 1) * np.pi * x) for k in range(5)], axis=1)
    xx, yy, nn = np.meshgrid(np.arange(64), np.arange(48), np.arange(5), indexing="ij")
    t = np.array([.005, .01, .02, .04, .08])
    rad = ((37 * xx + 101 * yy) % 997) / 997 * 20
    dn = np.clip(np.rint(255 * np.clip(rad * t[nn], 0, 1) ** (1 / 2.2)), 0, 255).astype(np.uint8)
    sd = 0.005 + 1e-5 * ((3 * xx + 5 * yy + 7 * nn) % 11)
    params = np.array([[0, 0, 0, 0, 0], [.1, 0, 0, 0, 0], [.05, -.03, .02, 0, .01], [0, 0, 0, 0, 2]], dtype=float).T
    e0 = np.array([0.010009522221360017, 0.09649365994071107, 0.04988969762746125, np.inf])
    e1 = np.array([0.007231459095189222, 0.07192212888277771, 0.017405883932145948, np.inf])
    assert_rel(cl.EnergyEvaluator(mean, pca, dn, None, 5, 250, True, t, 4)(params), e0, TIGHT)
    assert_rel(cl.EnergyEvaluator(mean, pca, dn, sd, 5, 250, True, t, 4)(params), e1, TIGHT)
    zero = cl.EnergyEvaluator(mean, pca, np.zeros_like(dn), None, 5, 250, True, t, 4)(params)
    assert np.isinf(zero).all()


@pytest.mark.parametrize("n_exp", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("use_std", [False, True])
def test_exposure_counts_and_population_sizes(n_exp, use_std):
    rng = np.random.default_rng(n_exp)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.0
    q, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.004 * 1.9 ** np.arange(n_exp)
    rad = rng.uniform(0, 1, (37, 23, 1)) * 30
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** 0.5).astype(np.uint8)
    sd = rng.uniform(0.002, 0.02, dn.shape) if use_std else None
    s = 40 + n_exp                                           # not a multiple of 32: padded internally
    params = rng.uniform(-0.04, 0.04, (5, s))
    params[:, 3] = [0, 0, 0, 0, 3.0]                          # gated
    e = cl.EnergyEvaluator(mean, q, dn, sd, 5, 250, True, t, s)(params)
    assert_rel(e, oe.energy_population(params, mean, q, dn, sd, 5, 250, True, t), TIGHT)


def test_zero_nan_and_denormal_variances_take_the_exact_branch():
    """The weighted pixel loop uses rsqrt's main path and revisits a pixel whose pair variance is not a normal
    positive number: sigma == 0 on both sides (pair skipped, general_functions.py:165-170), sigmas so small that
    the variance is denormal (library rsqrt), next to ordinary pixels."""
    rng = np.random.default_rng(11)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.0
    q, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.004 * 1.9 ** np.arange(5)
    rad = rng.uniform(0, 1, (64, 48, 1)) * 30
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** 0.5).astype(np.uint8)
    sd = rng.uniform(0.002, 0.02, dn.shape)
    sd[::7, ::5, :] = 0.0                        # every pair of the pixel has zero variance
    sd[1::7, ::5, 1:3] = 0.0                     # some pairs do
    sd[2::7, 1::5, :] = 2e-155                   # variances around 1e-309 .. 1e-306: denormal (~49 bits) and just normal
    sd[3::7, 2::5, 0] = 5e-155
    params = rng.uniform(-0.04, 0.04, (5, 37))
    with np.errstate(all="ignore"):
        want = oe.energy_population(params, mean, q, dn, sd, 5, 250, True, t)
    got = cl.EnergyEvaluator(mean, q, dn, sd, 5, 250, True, t, 37)(params)
    assert np.isfinite(want).any()
    assert_rel(got, want, 1e-9)                  # (the denormal weights are ~1e150: their sums lose a few digits)


def test_error_contract():
    x = np.linspace(0, 1, 256)
    pca = np.zeros((256, 5))
    t = np.array([0.01, 0.02])
    with pytest.raises(ValueError, match="image_stack must be a 3D"):
        cl.EnergyEvaluator(x, pca, np.zeros((4, 2), np.uint8), None, 5, 250, True, t, 1)
    with pytest.raises(ValueError, match="exposure_values must be a 1D"):
        cl.EnergyEvaluator(x, pca, np.zeros((4, 4, 3), np.uint8), None, 5, 250, True, t, 1)
    with pytest.raises(IndexError):
        cl.EnergyEvaluator(x, pca, np.zeros((4, 4, 2), np.float64), None, 5, 250, True, t, 1)


def test_full_size_cfg3_against_oracle_subset():
    """cfg3 size (S=64, 400k px x 5): GPU on all pixels == oracle recombined from its own pair sums is
    too slow on CPU; instead check linearity of the pair sums: whole = sum of two halves (exact
    counts, 1e-12 sums) and 3 candidates against the oracle on the full data."""
    rng = np.random.default_rng(3)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    q, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.005 * 2.0 ** np.arange(5)
    rad = rng.uniform(0, 1, (1000, 400, 1)) * 25
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** (1 / 2.2)).astype(np.uint8)
    params = rng.uniform(-0.05, 0.05, (5, 64))
    whole = cl.EnergyEvaluator(mean, q, dn, None, 5, 250, True, t, 64, shard=False)
    e = whole(params)
    acc = host(whole.plan.pair_acc).copy()
    halves = []
    for part in (dn[:500], dn[500:]):
        ev = cl.EnergyEvaluator(mean, q, part, None, 5, 250, True, t, 64, shard=False)
        ev(params)
        halves.append(host(ev.plan.pair_acc).copy())
    assert np.array_equal(acc[..., 1], halves[0][..., 1] + halves[1][..., 1])
    np.testing.assert_allclose(acc[..., 0], halves[0][..., 0] + halves[1][..., 0], rtol=1e-12)
    ref = oe.energy_population(params[:, :3], mean, q, dn, None, 5, 250, True, t)
    assert_rel(e[:3], ref, TIGHT)


def test_full_size_cfg3_with_std_against_oracle_subset():
    """cfg3 WITH uncertainty stacks (S=64, 400k px x 5, inverse-sigma weighted pair means): pair sums are
    linear in the pixels (whole = sum of two halves) and 3 candidates agree with the oracle on the full data."""
    rng = np.random.default_rng(33)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    q, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.005 * 2.0 ** np.arange(5)
    rad = rng.uniform(0, 1, (1000, 400, 1)) * 25
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** (1 / 2.2)).astype(np.uint8)
    sd = rng.uniform(0.002, 0.02, dn.shape)
    sd[::97, ::13, 2] = 0.0                       # sigma == 0 -> the pair is skipped (general_functions.py:165-170)
    params = rng.uniform(-0.05, 0.05, (5, 64))
    params[:, 9] = [0, 0, 0, 0, 2.0]              # gated
    whole = cl.EnergyEvaluator(mean, q, dn, sd, 5, 250, True, t, 64, shard=False)
    e = whole(params)
    assert np.isinf(e[9]) and np.isfinite(e).sum() >= 32
    acc = host(whole.plan.pair_acc).copy()
    halves = []
    for part, sp in ((dn[:500], sd[:500]), (dn[500:], sd[500:])):
        ev = cl.EnergyEvaluator(mean, q, part, sp, 5, 250, True, t, 64, shard=False)
        ev(params)
        halves.append(host(ev.plan.pair_acc).copy())
    live = np.isfinite(e)                         # (a gated candidate's tables may hold NaN)
    np.testing.assert_allclose(acc[live], (halves[0] + halves[1])[live], rtol=1e-12)
    ref = oe.energy_population(params[:, :3], mean, q, dn, sd, 5, 250, True, t)
    assert_rel(e[:3], ref, TIGHT)

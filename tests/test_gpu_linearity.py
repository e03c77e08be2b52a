"""Linearity-analysis chain (thresholds -> pair difference -> per-channel statistics) on the GPU vs
the unmodified reference (goldens) and the oracle.  float64 reductions: <= 1e-6 relative (north
star); asserted at 1e-11."""
import numpy as np
import pytest
import torch

from oracle import linearity as oli
from gpu_util import assert_rel, dev, host

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402

TIGHT = 1e-11


def _check(stats, a, r, use_std):
    s = host(stats)
    for which, st in ((0, a), (1, r)):
        assert_rel(s[which, 0], st["mean"], TIGHT)
        assert_rel(s[which, 1], st["std"], TIGHT)
        if use_std:
            assert_rel(s[which, 2], st["error"], TIGHT)
        else:
            assert np.isnan(s[which, 2]).all()


def test_golden_reference(golden_dir):
    g = np.load(golden_dir / "k5_linearity.npz")
    lower, upper = list(g["lower"]), list(g["upper"])
    for tag, use_std in (("std", True), ("nostd", False)):
        stats = ops.pair_statistics(dev(g["x_val"]), dev(g["x_std"]) if use_std else None, dev(g["y_val"]),
                                    dev(g["y_std"]) if use_std else None, float(g["multiplier"]), lower, upper)
        s = host(stats)
        for which, name in ((0, "abs"), (1, "rel")):
            assert_rel(s[which, 0], g[f"{tag}_{name}_mean"], TIGHT)
            assert_rel(s[which, 1], g[f"{tag}_{name}_std"], TIGHT)
            if use_std:
                assert_rel(s[which, 2], g[f"{tag}_{name}_error"], TIGHT)


@pytest.mark.parametrize("shape", [(64, 48, 3), (37, 29, 1), (20, 31, 4), (5, 3, 2), (300, 411, 3)])
@pytest.mark.parametrize("use_std", [True, False])
@pytest.mark.parametrize("with_thresholds", [True, False])
def test_random_pairs(shape, use_std, with_thresholds):
    rng = np.random.default_rng(shape[0] * 7 + shape[2])
    x = rng.random(shape) + 0.05
    y = rng.random(shape) + 0.05
    xs = rng.uniform(0.001, 0.02, shape) if use_std else None
    ys = rng.uniform(0.001, 0.02, shape) if use_std else None
    c = shape[-1]
    lower = [0.1 + 0.02 * i for i in range(c)] if with_thresholds else None
    upper = [0.9 - 0.03 * i if i % 2 == 0 else None for i in range(c)] if with_thresholds else None
    if not with_thresholds:
        x[0, 0, 0] = np.nan                      # already-thresholded input
        if use_std:
            xs[0, 0, 0] = np.nan
    a, r = oli.pair_statistics(x, xs, y, ys, 0.37, lower, upper)
    stats = ops.pair_statistics(dev(x), dev(xs), dev(y), dev(ys), 0.37, lower, upper)
    _check(stats, a, r, use_std)


def test_only_one_side_has_std():
    rng = np.random.default_rng(5)
    x, y = rng.random((16, 16, 3)) + 0.1, rng.random((16, 16, 3)) + 0.1
    xs = rng.uniform(0.001, 0.02, x.shape)
    a, r = oli.pair_statistics(x, xs, y, None, 0.5)
    _check(ops.pair_statistics(dev(x), dev(xs), dev(y), None, 0.5), a, r, True)


def test_exposure_series_process_linearity_uses_the_fused_path():
    from camera_linearity_b200 import GlobalSettings as gs
    rng = np.random.default_rng(6)
    gs.configure(NUM_OF_CHS=3)
    x = np.linspace(0, 1, 256)
    icrf = np.stack([x ** (2.0 + 0.1 * c) for c in range(3)], axis=1)
    t = [0.01, 0.02, 0.04]
    rad = rng.uniform(0, 1, (40, 32, 3)) * 30
    dn = [np.rint(255 * np.clip(rad * tk, 0, 1) ** (1 / 2.2)).astype(np.uint8) for tk in t]
    vals = [icrf[d, np.arange(3)] for d in dn]
    stds = [rng.uniform(0.002, 0.02, d.shape) for d in dn]
    feats = lambda tk: {"illumination": "bf", "magnification": "10x", "exposure": tk, "subject": "s"}
    sets = [cl.ImageSet(value=vals[k].copy(), std=stds[k].copy(), features=feats(t[k])) for k in range(3)]
    series = cl.ExposureSeries(input_image_sets=sets)
    series.initialize_exposure_pairs()
    series.process_linearity(icrf, linearity_limit=5, use_std=True)
    absolute, relative = series.collect_exposure_pair_stats()
    lower = [float(icrf[5, c]) for c in range(3)]
    upper = [float(icrf[250, c]) for c in range(3)]
    k = 0
    for i in range(3):
        for j in range(i + 1, 3):
            a, r = oli.pair_statistics(vals[i], stds[i], vals[j], stds[j], t[i] / t[j], lower, upper)
            assert_rel(relative["means"][k], r["mean"], TIGHT)
            assert_rel(relative["stds"][k], r["std"], TIGHT)
            assert_rel(absolute["errors"][k], a["error"], TIGHT)
            k += 1


@pytest.mark.parametrize("shape", [(701, 503, 3), (611, 517, 1), (400, 333, 4)])
def test_large_pairs_with_special_samples_against_the_oracle(shape):
    """Many grid-stride rounds per thread (the golden pair is small): NaN / zero-scale / thresholded samples at regular
    strides and in the ragged tail, repeat-identical, and the same data at an odd element offset (mono)."""
    rng = np.random.default_rng(shape[0])
    n = int(np.prod(shape))
    x, y = rng.random(shape) + 0.05, rng.random(shape) + 0.05
    xs, ys = rng.uniform(0.001, 0.02, shape), rng.uniform(0.001, 0.02, shape)
    fx, fy, fxs, fys = (a.reshape(-1) for a in (x, y, xs, ys))
    c = shape[-1]
    tile = (256 // c) * c * 2
    spots = np.concatenate([np.arange(0, n, tile)[:200], np.arange(tile - 1, n, tile)[:200], np.arange(n - 7, n)])
    fx[spots[0::5]] = np.nan
    fy[spots[1::5]] = np.nan
    fys[spots[2::5]] = np.nan
    fxs[spots[3::5]] = np.nan
    fy[spots[4::5]] = 0.0                        # scale == 0 -> inf / NaN relative difference
    lower, upper = [0.1] * c, [0.9 - 0.01 * i for i in range(c)]
    a, r = oli.pair_statistics(x, xs, y, ys, 0.41, lower, upper)
    first = ops.pair_statistics(dev(x), dev(xs), dev(y), dev(ys), 0.41, lower, upper)
    _check(first, a, r, True)
    assert torch.equal(first, ops.pair_statistics(dev(x), dev(xs), dev(y), dev(ys), 0.41, lower, upper))

    def odd(arr):                                # same values, pointer = 8 mod 16
        buf = torch.empty(n + c, dtype=torch.float64, device="cuda")
        view = buf[1:n + 1].view(shape) if c == 1 else None
        if view is None:                         # keep the channel phase: shift by one element only works for C == 1
            return None
        view.copy_(dev(arr))
        return view
    if c == 1:
        direct = ops.pair_statistics(odd(x), odd(xs), odd(y), odd(ys), 0.41, lower, upper)
        np.testing.assert_allclose(host(direct), host(first), rtol=1e-12, equal_nan=True)


def test_samples_that_are_neither_ordinary_nor_dropped():
    """No thresholds, so nothing is dropped wholesale: a zero scale (infinite ratio), zero variances (infinite weights),
    infinities and NaNs in only the value or only the uncertainty image must be counted term by term exactly as
    np.nansum does -- the lean per-sample routine hands them to the general one."""
    rng = np.random.default_rng(77)
    shape = (300, 257, 4)
    x, y = rng.random(shape) + 0.05, rng.random(shape) + 0.05
    xs, ys = rng.uniform(0.001, 0.02, shape), rng.uniform(0.001, 0.02, shape)
    x[5::41, 3::7, 0] = np.nan                   # value NaN, uncertainty kept
    xs[7::43, 1::5, 0] = np.nan                  # uncertainty NaN, value kept
    y[3::37, 2::9, 1] = 0.0                      # scale == 0: relative difference +-inf, its uncertainty inf
    xs[2::31, 4::11, 2] = 0.0
    ys[2::31, 4::11, 2] = 0.0                    # zero variance: infinite weight
    x[11::47, 5::13, 3] = np.inf
    ys[13::53, 6::17, 3] = 1e-170                # variance underflows to a denormal
    with np.errstate(all="ignore"):
        a, r = oli.pair_statistics(x, xs, y, ys, 0.8)
    s = host(ops.pair_statistics(dev(x), dev(xs), dev(y), dev(ys), 0.8))
    for which, st in ((0, a), (1, r)):
        for k, name in enumerate(("mean", "std", "error")):
            np.testing.assert_allclose(s[which, k], st[name], rtol=1e-10, equal_nan=True, err_msg=f"{which} {name}")
    assert np.isfinite(s[:, :, :2]).any()        # the case is not all-NaN

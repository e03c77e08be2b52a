"""The reference-facing Python API on the GPU: Measurand / ImageSet / ExposureSeries."""
import numpy as np
import pytest
import torch

from oracle import hdr_merge as om
from oracle import linearize as ol
from gpu_util import assert_rel, host, icrf_tables, synth_stack

pytestmark = pytest.mark.gpu
pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402
from camera_linearity_b200 import GlobalSettings as gs  # noqa: E402


def _features(exposure, subject="s"):
    return {"illumination": "bf", "magnification": "10x", "exposure": exposure, "subject": subject}


def test_measurand_linearize_reference_property():
    # tests/unit/test_measurand.py:447-467: each output channel only holds values of its LUT column
    rng = np.random.default_rng(0)
    for shape in [(6, 5, 3), (2, 3, 4, 2), (7, 1)]:
        c = shape[-1]
        icrf = np.stack([np.linspace(0, 1, 256) ** (i + 1) for i in range(c)], axis=1)
        diff = ol.default_icrf_diff(icrf)
        val = rng.random(shape)
        m = cl.Measurand(val, val * 0.1)
        out = m.linearize(icrf[:, 0], diff[:, 0]) if c == 1 else m.linearize(icrf, diff)
        assert isinstance(out.val, torch.Tensor) and out.val.is_cuda
        ev, es = ol.linearize(val, val * 0.1, icrf[:, 0] if c == 1 else icrf, diff[:, 0] if c == 1 else diff)
        assert np.array_equal(host(out.val), ev) and np.array_equal(host(out.std), es)
        for i in range(c):
            assert np.isin(host(out.val)[..., i], icrf[:, i]).all()


def test_measurand_hot_methods():
    rng = np.random.default_rng(1)
    val, std = rng.random((20, 24, 3)), rng.uniform(0.001, 0.02, (20, 24, 3))
    m = cl.Measurand(val, std)
    w, dw = m.apply_gaussian_weight()
    ew, edw = om.gaussian_weight(val)
    assert_rel(host(w), ew, 1e-14)
    dark = cl.Measurand(rng.random(val.shape) * 0.1)
    gs.configure(MEDIAN_FILTER_KERNEL_SIZE=3)
    f = m.filter_larger_than_by_map(dark, 0.06)
    ev, es = om.bad_pixel_filter(val, std, host(dark.val), 0.06, 3)
    assert np.array_equal(host(f.val), ev) and np.array_equal(host(f.std), es)
    gs.configure(IM_SIZE_X=20, IM_SIZE_Y=24, FF_MID_PERCENTAGE=0.2)
    flat = cl.Measurand(rng.uniform(0.5, 1, val.shape), rng.uniform(0.001, 0.01, val.shape))
    n = m.normalize_by_map(flat)
    ev, es = om.normalize_by_map(val, std, host(flat.val), host(flat.std), om.flat_roi_bounds(20, 24, 0.2))
    assert_rel(host(n.val), ev, 1e-13)
    assert_rel(host(n.std), es, 1e-13)


def test_exposure_series_process_hdr_image_in_memory():
    rng = np.random.default_rng(2)
    h, w = 48, 64
    gs.configure(IM_SIZE_X=h, IM_SIZE_Y=w, DARK_THRESHOLD=0.05, MEDIAN_FILTER_KERNEL_SIZE=3, FF_MID_PERCENTAGE=0.2)
    t = 0.005 * 2.0 ** np.arange(6)
    dn, std = synth_stack(rng, h, w, 3, t)
    icrf, diff = icrf_tables(3)
    dark_t = [0.02, 0.08, 0.32]
    dark_dn = []
    for _ in dark_t:
        d = rng.poisson(2.0, (h, w, 3)).astype(np.uint8)
        hot = rng.uniform(size=d.shape) < 0.02
        d[hot] = rng.integers(40, 200, int(hot.sum()))
        dark_dn.append(d)
    flat = np.clip(np.rint(rng.normal(180, 6, (h, w, 3))), 1, 255).astype(np.uint8)
    fstd = rng.uniform(0.001, 0.01, (h, w, 3))

    sets = [cl.ImageSet(value=dn[k], std=std[k], features=_features(float(t[k]))) for k in range(6)]
    rng.shuffle(sets)
    series = cl.ExposureSeries.from_multiple_image_sets(sets)[0]          # sorts by exposure
    darks = [cl.ImageSet(value=d, features=_features(e, "dark")) for d, e in zip(dark_dn, dark_t)]
    flats = [cl.ImageSet(value=flat, std=fstd, features=_features(0.0, "flat"))]
    series.process_HDR_image(icrf, diff, dark_list=darks, flat_list=flats)
    merged = series.merged_image_set
    assert merged.is_HDR and merged.measurand.val.is_cuda

    sel = [om.select_dark_field(float(tk), dark_t, 0.05) for tk in t]
    hd = [None if s is None else om.dark_value_image(dark_dn[s[0]], s[1]) for s in sel]
    ev, es = om.hdr_merge(dn, std, t, icrf, diff, darks=hd, dark_threshold=0.05, kernel=3, flat_val=flat / 255.0,
                          flat_std=fstd, roi=om.flat_roi_bounds(h, w, 0.2))
    assert_rel(host(merged.measurand.val), ev, 1e-11)
    assert_rel(host(merged.measurand.std), es, 1e-11)

    lin = series.linearize(icrf, diff)
    ev1, es1 = ol.linearize(dn[0], std[0], icrf, diff)
    assert np.array_equal(host(lin.input_image_sets[0].measurand.val), ev1)
    assert np.array_equal(host(lin.input_image_sets[0].measurand.std), es1)


def test_float_images_without_dn_are_requantised():
    rng = np.random.default_rng(3)
    t = [0.01, 0.02, 0.04]
    dn, std = synth_stack(rng, 32, 32, 3, np.array(t))
    icrf, diff = icrf_tables(3)
    sets = [cl.ImageSet(value=dn[k] / 255.0, std=std[k], features=_features(t[k])) for k in range(3)]
    series = cl.ExposureSeries(input_image_sets=sets)
    series.process_HDR_image(icrf, diff, dark_list=[], flat_list=[])
    ev, es = om.hdr_merge(dn, std, np.array(t), icrf, diff)
    assert_rel(host(series.merged_image_set.measurand.val), ev, 1e-11)


def test_operators_on_device_match_the_unmodified_reference(golden_dir):
    from gpu_util import dev
    from test_measurand_contract import _operator_results, check_operator_goldens
    g = np.load(golden_dir / "k8_operators.npz")
    check_operator_goldens(_operator_results(cl.Measurand, g, dev), g, 1e-13)


def test_linearize_follows_the_thresholded_value_image():
    """load -> apply_thresholds -> linearize must linearise the CURRENT measurand.val (thresholded
    NaN wraps to ICRF[0], measurand.py:502-505), not the integers decoded from the file."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (16, 20, 3), dtype=np.uint8)
    icrf, diff = icrf_tables(3)
    s = cl.ImageSet(features=_features(0.01))
    s.set_digital_numbers(img)
    s.load_value_image()
    assert s.dn is not None
    before = s.linearize(icrf, diff)
    ev, _ = ol.linearize(img, None, icrf, diff)
    assert np.array_equal(host(before.measurand.val), ev)
    s.measurand.apply_thresholds([0.2] * 3, [0.8] * 3)
    assert s.dn is None
    after = s.linearize(icrf, diff)
    val = img.astype(np.float64) / 255
    val[(val < 0.2) | (val > 0.8)] = np.nan
    with np.errstate(invalid="ignore"):
        ev2, _ = ol.linearize(val, None, icrf, diff)
    assert np.array_equal(host(after.measurand.val), ev2)
    assert not np.array_equal(ev, ev2)


def test_fused_operator_kernels_are_bit_identical_to_the_numpy_formulae():
    """csrc/measurand_ops.cu: + - * / with propagation follow measurand.py:106-241 operation for operation
    (round-to-nearest, no FMA contraction) -> bit-identical to NumPy; pow / log within libm's 1-2 ulp.  Same-shape,
    per-channel and scalar second operands, with and without uncertainties."""
    rng = np.random.default_rng(77)
    x, xs = rng.uniform(0.1, 2.0, (31, 17, 3)), rng.uniform(0.001, 0.05, (31, 17, 3))
    cases = [
        (rng.uniform(0.1, 2.0, (31, 17, 3)), rng.uniform(0.001, 0.05, (31, 17, 3))),       # same shape
        (rng.uniform(0.5, 1.5, (3,)), rng.uniform(0.001, 0.05, (3,))),                     # per channel
        (np.array([1.7]), np.array([0.03])),                                               # scalar-like
        (rng.uniform(0.1, 2.0, (31, 17, 3)), None),                                        # one-sided uncertainty
    ]
    for y, ys in cases:
        mx, my = cl.Measurand(x, xs), cl.Measurand(y, ys)
        s2 = np.zeros_like(y) if ys is None else ys
        with np.errstate(all="ignore"):
            expect = {
                "add": (x + y, np.sqrt(xs ** 2 + s2 ** 2)),
                "sub": (x - y, np.sqrt(xs ** 2 + s2 ** 2)),
                "mul": (x * y, np.sqrt((x * s2) ** 2 + (y * xs) ** 2)),
                "div": (x / y, np.sqrt((xs / y) ** 2 + ((x * s2) / (y ** 2)) ** 2)),
                "pow": (x ** y, np.sqrt(((y * x ** (y - 1)) * xs) ** 2 + ((np.log(x) * x ** y) * s2) ** 2)),
            }
        got = {"add": mx + my, "sub": mx - my, "mul": mx * my, "div": mx / my, "pow": mx ** my}
        for name in ("add", "sub", "mul", "div"):
            assert got[name].val.is_cuda
            assert np.array_equal(host(got[name].val), expect[name][0]), name
            assert np.array_equal(host(got[name].std), expect[name][1]), name
        assert_rel(host(got["pow"].val), expect["pow"][0], 1e-13)
        assert_rel(host(got["pow"].std), expect["pow"][1], 1e-12)
    # values only
    r = cl.Measurand(x) * cl.Measurand(cases[0][0])
    assert r.std is None and np.array_equal(host(r.val), x * cases[0][0])
    # scalar on the left (ImageSet.scale_to_exposure: scale * measurand)
    r = 0.25 * cl.Measurand(x, xs)
    assert np.array_equal(host(r.val), x * 0.25) and np.array_equal(host(r.std), np.sqrt((x * 0.0) ** 2 + (0.25 * xs) ** 2))
    # logarithms and the pair difference
    le, l10 = cl.Measurand(x, xs).log_e(), cl.Measurand(x, xs).log_10()
    assert_rel(host(le.val), np.log(x), 1e-14)
    assert_rel(host(le.std), xs / np.log(x), 1e-13)
    assert_rel(host(l10.val), np.log10(x), 1e-14)
    assert_rel(host(l10.std), xs / (x * (np.log(5) + np.log(2))), 1e-15)
    y, ys = cases[0]
    a, rel = cl.Measurand.compute_difference(cl.Measurand(x, xs), cl.Measurand(y, ys), 0.37)
    scale = 0.37 * y
    assert np.array_equal(host(a.val), x - scale) and np.array_equal(host(rel.val), (x - scale) / scale)
    assert np.array_equal(host(a.std), np.sqrt(xs ** 2 + (0.37 * ys) ** 2))
    assert np.array_equal(host(rel.std), np.sqrt((xs / (0.37 * y)) ** 2 + ((ys * x) / (0.37 * y ** 2)) ** 2))

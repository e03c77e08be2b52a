"""The C-ABI layer as PyTorch custom ops: every ``torch.ops.camera_linearity.*`` entry against the oracle."""
import numpy as np
import pytest
import torch

from oracle import egress as oeg
from oracle import hdr_merge as om
from oracle import icrf_energy as oe
from oracle import linearity as oli
from oracle import linearize as ol
from oracle import welford as ow
from gpu_util import assert_rel, dev, host, icrf_tables, synth_stack

pytestmark = pytest.mark.gpu
pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200  # noqa: E402,F401  (registers the ops)

K = torch.ops.camera_linearity
TIGHT = 1e-11


def test_registered_surface():
    for name in ("linearize", "hdr_merge", "hdr_merge_corrected", "flat_roi_means", "welford_stack", "welford_update",
                 "welford_finalize", "gaussian_weight", "icrf_energy_curves", "icrf_energy_partial",
                 "icrf_energy_finalize", "pair_statistics", "quantize_8bit", "noise_profiles", "measurand_binary"):
        assert hasattr(K, name), name


def test_hdr_merge_corrected_op_with_darks_flat_and_std_table():
    rng = np.random.default_rng(21)
    n, h, w = 6, 64, 80
    t = 0.004 * 1.9 ** np.arange(n)
    dn, std = synth_stack(rng, h, w, 3, t)
    icrf, diff = icrf_tables(3)
    thr = 0.05
    dark_t = [float(x) for x in t[t >= thr]]
    dark_dn = []
    for _ in dark_t:
        d = rng.poisson(2.0, (h, w, 3)).astype(np.uint8)
        hot = rng.uniform(size=d.shape) < 0.01
        d[hot] = rng.integers(13, 200, int(hot.sum()))
        dark_dn.append(d)
    sel = [om.select_dark_field(float(tk), dark_t, thr) for tk in t]
    host_darks = [None if s is None else om.dark_value_image(dark_dn[s[0]], s[1]) for s in sel]
    dark_index = [-1 if s is None else s[0] for s in sel]
    scales = [1.0 if s is None else s[1] for s in sel]
    flat = np.clip(np.rint(rng.normal(180, 6, (h, w, 3))), 1, 255).astype(np.uint8)
    fstd = rng.uniform(0.001, 0.01, (h, w, 3))
    roi = om.flat_roi_bounds(h, w, 0.2)
    means = K.flat_roi_means(dev(flat), dev(fstd), list(roi), 255.0)
    # (a) uncertainty images
    ev, es = om.hdr_merge(dn, std, t, icrf, diff, darks=host_darks, dark_threshold=thr, kernel=3, flat_val=flat / 255.0,
                          flat_std=fstd, roi=roi)
    v, s = K.hdr_merge_corrected([dev(d) for d in dn], [dev(x) for x in std], list(range(n)), None, [float(x) for x in t],
                                 dev(icrf), dev(diff), [dev(d) for d in dark_dn], dark_index, scales, thr, 3, dev(flat),
                                 dev(fstd), means, 0)
    assert_rel(host(v), ev, TIGHT)
    assert_rel(host(s), es, TIGHT)
    # (b) no uncertainty images: the camera's STD table
    std_lut = 0.002 + 0.02 * np.sqrt(np.linspace(0, 1, 256))[:, None] * np.array([1.0, 0.9, 1.1])
    std_t = [std_lut[d, np.arange(3)] for d in dn]
    ev, es = om.hdr_merge(dn, std_t, t, icrf, diff, darks=host_darks, dark_threshold=thr, kernel=3)
    v, s = K.hdr_merge_corrected([dev(d) for d in dn], [], [], dev(std_lut), [float(x) for x in t], dev(icrf), dev(diff),
                                 [dev(d) for d in dark_dn], dark_index, scales, thr, 3, None, None, None, 0)
    assert_rel(host(v), ev, TIGHT)
    assert_rel(host(s), es, TIGHT)


def test_welford_update_and_finalize_ops():
    rng = np.random.default_rng(22)
    frames = np.clip(rng.integers(20, 231, (1, 10, 14, 3)) + rng.integers(-3, 4, (17, 10, 14, 3)), 0, 255).astype(np.uint8)
    o = ow.welford(list(frames))
    mean = torch.zeros((10, 14, 3), dtype=torch.float64, device="cuda")
    m2 = torch.zeros_like(mean)
    K.welford_update(dev(frames[:9]), mean, m2, 0, None, 255.0)
    K.welford_update(dev(frames[9:]), mean, m2, 9, None, 255.0)
    assert np.array_equal(host(mean), o["mean"]) and np.array_equal(host(m2), o["m2"])
    sem, mean_u8 = K.welford_finalize(mean, m2, 17, 255.0)
    assert np.array_equal(host(sem), o["sem"]) and np.array_equal(host(mean_u8), o["mean_u8"])
    m, s_, u8 = K.welford_stack(dev(frames), None, 255.0)
    assert np.array_equal(host(u8), o["mean_u8"])


def test_icrf_energy_ops_chain():
    rng = np.random.default_rng(23)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    pca, _ = np.linalg.qr(np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1))
    t = 0.005 * 2.0 ** np.arange(4)
    rad = rng.uniform(0, 1, (57, 43, 1)) * 25
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** (1 / 2.2)).astype(np.uint8)
    sd = rng.uniform(0.002, 0.02, dn.shape)
    params = rng.uniform(-0.05, 0.05, (5, 32))
    params[:, 7] = [0, 0, 0, 0, 3.0]
    for use_std in (False, True):
        curves, valid, tables = K.icrf_energy_curves(dev(np.ascontiguousarray(params.T)), dev(mean), dev(pca), 5, 250, 4,
                                                     use_std)
        assert int(valid[7]) == 0
        # two pixel shards, summed like an all-reduce would
        flat_dn, flat_sd = dn.reshape(-1, 4), sd.reshape(-1, 4)
        acc = None
        for lo, hi in ((0, 1000), (1000, flat_dn.shape[0])):
            part = K.icrf_energy_partial(tables, dev(flat_dn[lo:hi]), dev(flat_sd[lo:hi]) if use_std else None,
                                         [float(v) for v in t], 32, 5, 256, True, 5, 250)
            acc = part if acc is None else acc + part
        e = host(K.icrf_energy_finalize(acc, valid, 4))
        expect = oe.energy_population(params, mean, pca, dn, sd if use_std else None, 5, 250, True, t)
        assert_rel(e, expect, TIGHT)


def test_pair_statistics_and_quantize_ops(golden_dir):
    rng = np.random.default_rng(24)
    xv, yv = rng.random((40, 50, 3)) + 0.05, rng.random((40, 50, 3)) + 0.05
    xs, ys = rng.uniform(0.001, 0.02, xv.shape), rng.uniform(0.001, 0.02, xv.shape)
    got = host(K.pair_statistics(dev(xv), dev(xs), dev(yv), dev(ys), 0.5, [0.1] * 3, [0.9] * 3))
    a, r = oli.pair_statistics(xv, xs, yv, ys, 0.5, [0.1] * 3, [0.9] * 3)
    for which, st in ((0, a), (1, r)):
        assert_rel(got[which, 0], st["mean"], TIGHT)
        assert_rel(got[which, 1], st["std"], TIGHT)
        assert_rel(got[which, 2], st["error"], TIGHT)
    val = rng.random((30, 20, 3)) * 1.7
    out, mx = K.quantize_8bit(dev(val), 255.0)
    assert np.array_equal(host(out), oeg.quantize_8bit(val))
    assert float(mx) == val.max()


def test_operators_follow_the_device_of_their_tensors():
    """tensors on a device that is not the current one must still run on THEIR device's stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(25)
    icrf, diff = icrf_tables(3)
    img = rng.integers(0, 256, (32, 40, 3), dtype=np.uint8)
    other = torch.device("cuda", 1)
    assert torch.cuda.current_device() == 0
    (v,) = K.linearize(torch.from_numpy(img).to(other), None, torch.from_numpy(icrf).to(other), None, 255.0)
    assert v.device == other
    ev, _ = ol.linearize(img, None, icrf, None)
    assert np.array_equal(v.cpu().numpy(), ev)
    assert torch.cuda.current_device() == 0


def test_noise_profiles_and_measurand_binary_ops():
    from oracle import noise_profiles as onp
    rng = np.random.default_rng(26)
    base = rng.integers(0, 256, (20, 24, 3))
    video = [np.clip(base + rng.integers(-4, 5, base.shape), 0, 255).astype(np.uint8) for _ in range(11)]
    exp, mean_frame = onp.noise_profiles([video])
    assert np.array_equal(host(K.noise_profiles(dev(np.stack(video)), dev(mean_frame))), exp)
    x, xs = rng.uniform(0.1, 2, (20, 24, 3)), rng.uniform(0.001, 0.05, (20, 24, 3))
    y, ys = rng.uniform(0.5, 1.5, (3,)), rng.uniform(0.001, 0.05, (3,))
    v, s = K.measurand_binary("div", dev(x), dev(xs), dev(y), dev(ys))
    assert np.array_equal(host(v), x / y)
    assert np.array_equal(host(s), np.sqrt((xs / y) ** 2 + ((x * ys) / (y ** 2)) ** 2))

"""The NumPy oracle against (a) the golden vectors generated from the real reference
(tests/golden/make_golden.py) and (b) the survey's RNG-free known-answer values
(SURVEY.md section 8c).  CPU only."""
import numpy as np
import pytest

from oracle import hdr_merge as om
from oracle import icrf_energy as oe
from oracle import linearize as ol
from oracle import welford as ow


def _load(golden_dir, name):
    return np.load(golden_dir / name)


# ------------------------------------------------------------------ K1
def test_linearize_matches_reference_bitexact(golden_dir):
    g = _load(golden_dir, "k1_linearize.npz")
    v, s = ol.linearize(g["val"], g["std"], g["icrf"], g["icrf_diff"])
    assert np.array_equal(v, g["exp_val"]) and np.array_equal(s, g["exp_std"])
    v, s = ol.linearize(g["dn"], g["dn_std"], g["icrf"], g["icrf_diff"])
    assert np.array_equal(v, g["exp_dn_val"]) and np.array_equal(s, g["exp_dn_std"])
    v, s = ol.linearize(g["mono"], g["mono_std"], g["icrf"][:, 1], g["icrf_diff"][:, 1])
    assert np.array_equal(v, g["exp_mono_val"]) and np.array_equal(s, g["exp_mono_std"])


def test_lut_index_wrap_and_half_even():
    # survey probe (x86 NumPy): [256, 257, -1, nan, 300.4] -> [0, 1, 255, 0, 44]
    x = np.array([256, 257, -1, np.nan, 300.4, 0.5, 1.5, 2.5]) / 255.0
    assert ol.lut_index(x).tolist() == [0, 1, 255, 0, 44, 0, 2, 2]


@pytest.mark.parametrize("max_dn", [255, 65535])
def test_dn_roundtrip_is_identity(max_dn):
    # u8/u16 DNs may stay integer on the device: rint((k/MAX)*MAX) == k for every k
    k = np.arange(max_dn + 1)
    assert np.array_equal(np.around((k / max_dn) * max_dn).astype(np.int64), k)


def test_reference_own_linearize_property():
    # tests/unit/test_measurand.py:447-467 (fails at HEAD, satisfied by R1)
    rng = np.random.default_rng(0)
    for shape in [(4, 5, 3), (2, 3, 4, 2), (6, 5)]:
        c = shape[-1]
        icrf = np.stack([np.linspace(0, 1, 256) ** (i + 1) for i in range(c)], axis=1)
        val = rng.random(shape)
        out, _ = ol.linearize(val, None, icrf)
        assert out.shape == val.shape
        for i in range(c):
            assert np.isin(out[..., i], icrf[..., i]).all()


def test_gaussian_weight_matches_reference(golden_dir):
    g = _load(golden_dir, "gaussian_weight.npz")
    w, dw = om.gaussian_weight(g["v"])
    assert np.array_equal(w, g["w"]) and np.array_equal(dw, g["dw"])


# ------------------------------------------------------------------ K2
def _darks_for(g, max_dn=255):
    darks = []
    thr = float(g["dark_threshold"])
    for tk in g["t"]:
        sel = om.select_dark_field(float(tk), [float(x) for x in g["dark_t"]], thr)
        darks.append(None if sel is None else om.dark_value_image(g["dark_dn"][sel[0]], sel[1], max_dn))
    return darks


def test_merge_katm_reference_and_survey_values(golden_dir):
    g = _load(golden_dir, "k2_merge_katm.npz")
    val, std = om.hdr_merge(list(g["dn"]), list(g["std"]), g["t"], g["icrf"], g["icrf_diff"])
    assert np.array_equal(val, g["exp_val"]) and np.array_equal(std, g["exp_std"])
    # SURVEY.md 8(c) KAT-M
    assert [int(d.sum()) for d in g["dn"]] == [4884, 9818, 14800, 19659, 23458]
    assert val[0, 0].tolist() == [0.0, 0.02700191237234568, 0.06691215637264117]
    assert std[0, 0].tolist() == [2.8811886741485363e-08, 5.31822796486957e-05, 0.00017582004682363065]
    assert val[5, 7].tolist() == [1.0782387936433966, 1.0914821970876645, 1.1461947064885876]
    assert std[5, 7].tolist() == [0.008937988533996381, 0.0075688841926915395, 0.006502517830747571]
    assert val.sum() == pytest.approx(1341.4495199747855, rel=1e-13)
    assert std.sum() == pytest.approx(9.06360629426101, rel=1e-13)


def test_merge_dark_and_flat_matches_reference(golden_dir):
    g = _load(golden_dir, "k2_merge_dark_flat.npz")
    darks = _darks_for(g)
    assert sum(d is not None for d in darks) == 2          # t = .08 (exact) and .16 (scaled)
    val, std = om.hdr_merge(list(g["dn"]), list(g["std"]), g["t"], g["icrf"], g["icrf_diff"],
                            darks=darks, dark_threshold=float(g["dark_threshold"]),
                            kernel=int(g["kernel"]))
    assert np.array_equal(val, g["exp_val_dark"]) and np.array_equal(std, g["exp_std_dark"])
    roi = om.flat_roi_bounds(int(g["im_size_x"]), int(g["im_size_y"]), float(g["ff_mid"]))
    val, std = om.hdr_merge(list(g["dn"]), list(g["std"]), g["t"], g["icrf"], g["icrf_diff"],
                            darks=darks, dark_threshold=float(g["dark_threshold"]),
                            kernel=int(g["kernel"]), flat_val=g["flat_dn"] / 255.0,
                            flat_std=g["flat_std"], roi=roi)
    assert np.array_equal(val, g["exp_val"]) and np.array_equal(std, g["exp_std"])


def test_merge_kernel5_matches_reference(golden_dir):
    g = _load(golden_dir, "k2_merge_k5.npz")
    val, std = om.hdr_merge(list(g["dn"]), list(g["std"]), g["t"], g["icrf"], g["icrf_diff"],
                            darks=_darks_for(g), dark_threshold=float(g["dark_threshold"]),
                            kernel=int(g["kernel"]))
    assert np.array_equal(val, g["exp_val"]) and np.array_equal(std, g["exp_std"])


def test_median_semantics_bruteforce():
    # scipy 'reflect' = half-sample symmetric (d c b a | a b c d); rank K*K//2; per channel
    rng = np.random.default_rng(3)
    img = rng.random((7, 6, 2))
    for k in (2, 3, 4, 5):
        out, _ = om.bad_pixel_filter(img, None, np.ones_like(img), 0.5, k)
        h, w, _ = img.shape
        lo = k // 2
        for y in range(h):
            for x in range(w):
                for c in range(2):
                    win = []
                    for dy in range(-lo, k - lo):
                        for dx in range(-lo, k - lo):
                            yy, xx = y + dy, x + dx
                            yy = -yy - 1 if yy < 0 else (2 * h - 1 - yy if yy >= h else yy)
                            xx = -xx - 1 if xx < 0 else (2 * w - 1 - xx if xx >= w else xx)
                            win.append(img[yy, xx, c])
                    assert out[y, x, c] == sorted(win)[k * k // 2]


def test_median_of_dn_equals_median_of_scaled():
    rng = np.random.default_rng(4)
    dn = rng.integers(0, 256, (9, 8, 3), dtype=np.uint8)
    a, _ = om.bad_pixel_filter(dn / 255.0, None, np.ones(dn.shape), 0.5, 3)
    b, _ = om.bad_pixel_filter(dn.astype(np.float64), None, np.ones(dn.shape), 0.5, 3)
    assert np.array_equal(a, b / 255.0)


# ------------------------------------------------------------------ K3
def test_welford_matches_reference_and_survey(golden_dir):
    g = _load(golden_dir, "k3_welford.npz")
    r = ow.welford(list(g["katw"]))
    assert np.array_equal(r["mean_u8"], g["katw_mean_u8"])
    assert np.array_equal(r["std_u8"], g["katw_std_u8"]) and not r["std_u8"].any()   # D13
    assert int(r["mean_u8"].sum()) == 19199
    assert r["mean_u8"][0, 0].tolist() == [39, 44, 49] and r["mean_u8"][5, 7].tolist() == [72, 77, 82]
    assert r["mean"][0, 0].tolist() == [0.15294117647058825, 0.17254901960784313, 0.19215686274509805]
    assert np.all(r["sem"][0, 0] == 0.05998846486579746)
    assert r["mean"].sum() == pytest.approx(75.32549019607842, rel=1e-14)
    assert r["sem"].sum() == pytest.approx(13.802924071111722, rel=1e-14)
    r = ow.welford(list(g["frames"]))
    assert np.array_equal(r["mean_u8"], g["mean_u8"])


# ------------------------------------------------------------------ K4
def test_energy_matches_reference_bitexact(golden_dir):
    g = _load(golden_dir, "k4_energy.npz")
    e = oe.energy_population(g["params"], g["mean"], g["pca"], g["dn"], None, 5, 250, True, g["t"])
    assert np.array_equal(e, g["e_nostd"])
    e = oe.energy_population(g["params"], g["mean"], g["pca"], g["dn"], g["std"], 5, 250, True, g["t"])
    assert np.array_equal(e, g["e_std"])
    e = oe.energy_population(g["params6"], g["mean"], g["pca"], g["dn"], None, 5, 250, False, g["t"])
    assert np.array_equal(e, g["e6"])
    assert np.isinf(g["e_nostd"]).any() and np.isfinite(g["e_nostd"]).any()


def test_energy_survey_kat():
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    pca = np.stack([0.1 * np.sin((k + 1) * np.pi * x) for k in range(5)], axis=1)
    xx, yy, nn = np.meshgrid(np.arange(64), np.arange(48), np.arange(5), indexing="ij")
    t = np.array([.005, .01, .02, .04, .08])
    rad = ((37 * xx + 101 * yy) % 997) / 997 * 20
    dn = np.clip(np.rint(255 * np.clip(rad * t[nn], 0, 1) ** (1 / 2.2)), 0, 255).astype(np.uint8)
    assert int(dn.sum()) == 1921067
    sd = 0.005 + 1e-5 * ((3 * xx + 5 * yy + 7 * nn) % 11)
    kat = [([0, 0, 0, 0, 0], 0.010009522221360017, 0.007231459095189222),
           ([.1, 0, 0, 0, 0], 0.09649365994071107, 0.07192212888277771),
           ([.05, -.03, .02, 0, .01], 0.04988969762746125, 0.017405883932145948),
           ([0, 0, 0, 0, 2], np.inf, np.inf)]
    for p, e0, e1 in kat:
        p = np.array(p, dtype=float)
        assert oe.energy(p, mean, pca, dn, None, 5, 250, True, t) == e0
        assert oe.energy(p, mean, pca, dn, sd, 5, 250, True, t) == e1
    assert oe.energy(np.zeros(5), mean, pca, np.zeros_like(dn), None, 5, 250, True, t) == np.inf


# ------------------------------------------------------------------ linearity chain (8f rank 1)
def test_linearity_chain_matches_reference_bitexact(golden_dir):
    from oracle import linearity as oli
    g = _load(golden_dir, "k5_linearity.npz")
    lower, upper = list(g["lower"]), list(g["upper"])
    for tag, use_std in (("std", True), ("nostd", False)):
        a, r = oli.pair_statistics(g["x_val"], g["x_std"] if use_std else None, g["y_val"],
                                   g["y_std"] if use_std else None, float(g["multiplier"]), lower, upper)
        for name, st in (("abs", a), ("rel", r)):
            assert np.array_equal(st["mean"], g[f"{tag}_{name}_mean"])
            assert np.array_equal(st["std"], g[f"{tag}_{name}_std"])
            if use_std:
                assert np.array_equal(st["error"], g[f"{tag}_{name}_error"])
            else:
                assert st["error"] is None


# ------------------------------------------------------------------ 8-bit export (8f rank 2)
def test_save_8bit_matches_reference_bytes(golden_dir):
    from oracle import egress
    g = _load(golden_dir, "k6_save_8bit.npz")
    for name in ("hdr", "unit", "ties", "negative"):
        assert np.array_equal(egress.quantize_8bit(g[f"{name}_val"]), g[f"{name}_val_u8"])
        assert np.array_equal(egress.quantize_8bit(g[f"{name}_std"]), g[f"{name}_std_u8"])


def test_reciprocal_quotient_is_exact():
    # csrc/hdr_merge_wide.cu evaluates dn / MAX_DN as fma(fma(-q0, MAX, dn), RN(1/MAX), q0), q0 = dn * RN(1/MAX);
    # that equals the IEEE quotient NumPy computes for every 8- and 16-bit dn (exact rational check)
    from fractions import Fraction as F
    for b in (255.0, 65535.0):
        r = 1.0 / b
        for d in range(int(b) + 1):
            q0 = d * r
            rem = float(F(d) - F(q0) * F(b))
            assert float(F(rem) * F(r) + F(q0)) == d / b


# ------------------------------------------------------------------ channel histogram (8f rank 4)
def test_channel_histogram_matches_reference(golden_dir):
    from oracle import histogram as oh
    g = _load(golden_dir, "k7_histogram.npz")
    for tag, bins, rng_, use_std in (("a", 8, (0.0, 1.0), False), ("b", 64, None, False), ("c", 50, (0.1, 0.7), True),
                                     ("d", 5000, None, True)):
        for c in range(3):
            h, e = oh.channel_histogram(g["val"], g["std"], c, bins, rng_, use_std)
            assert np.array_equal(h, g[f"{tag}_hist_{c}"]) and np.array_equal(e, g[f"{tag}_edges_{c}"])


def test_welford_with_icrf_matches_the_unmodified_reference(golden_dir):
    # the reference's `if ICRF:` branch, reached with an always-true ndarray subclass (make_golden.py)
    g = _load(golden_dir, "k3_welford_icrf.npz")
    from oracle import welford as ow
    r = ow.welford(list(g["frames"]), g["icrf"])
    assert np.array_equal(r["mean_u8"], g["mean_u8"]) and np.array_equal(r["std_u8"], g["std_u8"])
    hh, ww, cc = np.meshgrid(np.arange(6), np.arange(8), np.arange(3), indexing="ij")
    katw = [((31 * hh + 17 * ww + 5 * cc + 3 * f * f + f * hh) % 256).astype(np.uint8) for f in range(7)]
    r = ow.welford(katw, g["icrf"])
    assert np.array_equal(r["mean_u8"], g["katw_mean_u8"]) and np.array_equal(r["std_u8"], g["katw_std_u8"])


def test_noise_profiles_match_the_unmodified_reference(golden_dir):
    """compute_noise_profiles (video_processing.py:77-106) run unmodified in make_golden.py: joint histogram of
    (uint8 mean DN, frame DN) per channel over both videos -- integer work, bit-exact."""
    from oracle import noise_profiles as onp
    g = np.load(golden_dir / "k9_noise_profiles.npz")
    profiles, mean_frame = onp.noise_profiles([list(g["frames0"]), list(g["frames1"])])
    assert np.array_equal(mean_frame, g["mean_frame"])
    assert profiles.dtype == np.int64 and np.array_equal(profiles, g["profiles"])
    assert profiles.sum() == (len(g["frames0"]) + len(g["frames1"])) * g["mean_frame"].size

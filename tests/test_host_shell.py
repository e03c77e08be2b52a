"""Host-side logic of ImageSet / ExposureSeries / settings, following the reference's unit tests
(tests/unit/test_image_set.py:109-361, tests/unit/test_exposure_series.py:26-128)."""
from pathlib import Path
from unittest.mock import MagicMock, patch

import numpy as np
import pytest
import torch

from camera_linearity_b200 import ExposureSeries, GlobalSettings, ImageSet, Measurand
from camera_linearity_b200.image_set import _features_from_file_name
from oracle import hdr_merge as om



@pytest.fixture(autouse=True, scope="module")
def _cpu_tensors():
    """Host-logic tests run on CPU tensors; restore the default device afterwards."""
    previous = GlobalSettings.DEVICE
    GlobalSettings.DEVICE = "cpu"
    yield
    GlobalSettings.DEVICE = previous


def test_image_set_defaults():
    s = ImageSet()
    assert s.path is None and s.features is None and s.is_HDR is False
    assert s.measurand.val is None and s.measurand.std is None
    with pytest.raises(AttributeError):
        s.use_cupy = False


@pytest.mark.parametrize("name,expected", [
    ("5ms BF sample_1 50x.tif", {"illumination": "BF", "magnification": "50x", "exposure": 0.005, "subject": "sample_1"}),
    ("10x df 12.5ms flat.tif", {"illumination": "df", "magnification": "10x", "exposure": 0.0125, "subject": "flat"}),
    ("dark 100ms.tif", {"illumination": "", "magnification": "", "exposure": 0.1, "subject": "dark"}),
])
def test_features_from_file_name(name, expected):
    assert _features_from_file_name(Path(name)) == expected


def test_multiple_from_path_skips_std(tmp_path):
    for n in ("5ms BF a 10x.tif", "5ms BF a 10x STD.tif", "10ms BF a 10x.tif", "notes.txt"):
        (tmp_path / n).write_bytes(b"")
    sets = ImageSet.multiple_from_path(tmp_path)
    assert sorted(s.path.name for s in sets) == ["10ms BF a 10x.tif", "5ms BF a 10x.tif"]


def test_is_exposure_match():
    a = ImageSet(file_path="5ms BF a 10x.tif")
    assert a.is_exposure_match(ImageSet(file_path="50ms BF a 10x.tif"))
    assert not a.is_exposure_match(ImageSet(file_path="5ms DF a 10x.tif"))
    assert not a.is_exposure_match(ImageSet())


def test_load_value_image_scales_by_max_dn():
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    with patch("cv2.imread", return_value=img):
        s = ImageSet(file_path="5ms BF a 10x.tif")
        s.load_value_image()
        assert np.array_equal(s.measurand.val.numpy(), img.astype(np.float64) / 255)
        assert np.array_equal(s.dn.numpy(), img)
        s2 = ImageSet(file_path="5ms BF a 10x.tif")
        s2.load_value_image(bit64=True)
        assert np.array_equal(s2.measurand.val.numpy(), img)


def test_forwarding_to_the_measurand():
    m = MagicMock(spec=Measurand)
    m.val = None
    m.std = None
    other = MagicMock(spec=Measurand)
    m.extract.return_value = other
    s = ImageSet(measurand=m, features={"exposure": 1.0})
    out = s.extract([0, 1])
    m.extract.assert_called_once_with(dims=[0, 1], axis=-1)
    assert out.measurand is other
    m.linearize.return_value = other
    assert s.linearize("icrf", "diff").measurand is other
    m.linearize.assert_called_once_with("icrf", "diff")


def test_exposure_interpolation_errors():
    a = ImageSet(value=np.ones((2, 2, 3)), features={"exposure": 1.0})
    b = ImageSet(value=np.ones((2, 2, 3)) * 3, features={"exposure": 3.0})
    with pytest.raises(TypeError):
        ImageSet.exposure_interpolation(a, b, 2)
    with pytest.raises(ValueError):
        ImageSet.exposure_interpolation(a, b, 4.0)
    mid = ImageSet.exposure_interpolation(a, b, 2.0)
    assert np.allclose(mid.measurand.val.numpy(), 2.0)


def test_dark_selection_matches_oracle_restatement():
    GlobalSettings.configure(DARK_THRESHOLD=0.05)
    rng = np.random.default_rng(0)
    for _ in range(200):
        dark_t = [float(x) for x in rng.choice([0.01, 0.02, 0.05, 0.08, 0.1, 0.32, 1.0], size=rng.integers(1, 5), replace=False)]
        target = float(rng.choice([0.01, 0.05, 0.08, 0.09, 0.2, 2.0]))
        darks = [ImageSet(features={"illumination": "", "magnification": "", "exposure": e, "subject": "dark"}) for e in dark_t]
        s = ImageSet(features={"illumination": "bf", "magnification": "10x", "exposure": target, "subject": "s"})
        got = s.select_dark_field(darks)
        want = om.select_dark_field(target, dark_t, 0.05)
        if want is None:
            assert got is None
        else:
            assert got[0] is darks[want[0]] and got[1] == want[1]


def test_scale_to_exposure_does_not_alias_features():
    d = ImageSet(value=np.full((2, 2, 3), 0.5), features={"exposure": 0.4, "subject": "dark", "illumination": "", "magnification": ""})
    s = d.scale_to_exposure(0.1)
    assert d.features["exposure"] == 0.4 and s.features["exposure"] == 0.1
    assert np.allclose(s.measurand.val.numpy(), 0.125)


def test_exposure_series_construction():
    e = ExposureSeries()
    assert e.merged_image_set is None and e.input_image_sets == [] and e.exposure_pairs is None
    assert ExposureSeries(directory_path=Path("/a/b/c.tif")).directory_path == Path("/a/b")
    assert ExposureSeries(directory_path=Path("/a/b")).directory_path == Path("/a/b")
    with pytest.raises(AttributeError):
        e.use_cupy = True
    sets = [ImageSet(file_path=f"{t}ms BF a 10x.tif") for t in (40, 5, 20)] + [ImageSet(file_path="7ms DF b 10x.tif")]
    series = ExposureSeries.from_multiple_image_sets(sets)
    assert len(series) == 2
    assert [s.features["exposure"] for s in series[0].input_image_sets] == [0.005, 0.02, 0.04]
    series[0].initialize_exposure_pairs()
    ratios = sorted(p.exposure_ratio for p in series[0].exposure_pairs)
    assert ratios == sorted([0.005 / 0.02, 0.005 / 0.04, 0.02 / 0.04])


def test_settings_from_ini(tmp_path):
    ini = tmp_path / "config.ini"
    ini.write_text("[Integer data]\nimage size x = 640\nimage size y = 480\nbit depth = 8\nmedian filter kernel size = 5\n"
                   "[Float data]\ndark threshold = 0.07\ninitial guess = 0,0.5,0,0,0\n"
                   "[Paths]\ndark frames path = /tmp/darks\nchannel names = Blue,Green,Red\n")
    GlobalSettings.from_ini(ini)
    try:
        assert GlobalSettings.IM_SIZE_X == 640 and GlobalSettings.MEDIAN_FILTER_KERNEL_SIZE == 5
        assert GlobalSettings.DARK_THRESHOLD == 0.07 and GlobalSettings.IN_PCA_GUESS == [0, 0.5, 0, 0, 0]
        assert GlobalSettings.DEFAULT_DARK_PATH == Path("/tmp/darks") and GlobalSettings.BITS == 256
    finally:
        GlobalSettings.configure(IM_SIZE_X=2048, IM_SIZE_Y=1536, MEDIAN_FILTER_KERNEL_SIZE=3, DARK_THRESHOLD=0.05,
                                 IN_PCA_GUESS=[0.0] * 5, DEFAULT_DARK_PATH=Path("data/dark"))


def test_tiff_roundtrip_64bit(tmp_path):
    # tests/integration/test_integration_image_set.py:48-83 (the 8-bit half needs the GPU: test_gpu_egress.py)
    rng = np.random.default_rng(1)
    val = rng.random((8, 9, 3))
    s = ImageSet(file_path=tmp_path / "5ms BF a 10x.tif", value=val, std=val * 0.1)
    s.save_64bit(tmp_path / "64" / "5ms BF a 10x.tif")
    back = ImageSet(file_path=tmp_path / "64" / "5ms BF a 10x.tif")
    back.load_value_image(bit64=True)
    back.load_std_image()
    assert np.allclose(back.measurand.val.numpy(), val) and np.allclose(back.measurand.std.numpy(), val * 0.1)


def test_device_only_methods_fail_loudly_on_host_tensors(tmp_path):
    # no CPU fallback: the 8-bit export and the histogram are CUDA kernels
    rng = np.random.default_rng(2)
    val = rng.random((8, 9, 3))
    s = ImageSet(file_path=tmp_path / "5ms BF a 10x.tif", value=val, std=val * 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.save_8bit(tmp_path / "8" / "5ms BF a 10x.tif")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.measurand.compute_channel_histogram(16, (0.0, 1.0))


def test_de_oracle_step_properties():
    # oracle/de.py (the spec of csrc/de.cu): distinct members, crossover floor, bounds, promotion
    from oracle import de as ode
    rng = np.random.default_rng(0)
    S, P = 64, 5
    pop = rng.uniform(0, 1, (S, P))
    for gen in range(20):
        i = np.arange(S)
        r0 = (ode.draw(9, gen, i, ode.SLOT_R0) * (S - 1)).astype(np.int64)
        r0 += r0 >= i
        r1 = (ode.draw(9, gen, i, ode.SLOT_R1) * (S - 2)).astype(np.int64)
        a, b = np.minimum(i, r0), np.maximum(i, r0)
        r1 += r1 >= a
        r1 += r1 >= b
        assert np.all(r0 != i) and np.all(r1 != i) and np.all(r0 != r1)
        assert r0.min() >= 0 and r0.max() < S and r1.min() >= 0 and r1.max() < S
        trial, params = ode.trial_population(pop, 9, gen, (0.0, 1.95), 0.4, [-2] * P, [2] * P)
        assert trial.min() >= 0 and trial.max() <= 1
        assert np.all((trial != pop).any(axis=1))            # the fill point always takes the mutant
        assert np.allclose(params, (trial - 0.5) * 4)
        e = rng.uniform(0, 1, S)
        te = rng.uniform(0, 1, S)
        pop, e2, st = ode.select(pop, e, trial, te)
        assert e2[0] == e2.min() == min(e.min(), te.min())
    u = ode.draw(1, 2, np.arange(100000), 3)
    assert 0.49 < u.mean() < 0.51 and u.min() >= 0 and u.max() < 1


def test_decoded_integers_follow_the_current_value_image():
    """ImageSet keeps the decoded integers next to ``measurand.val`` as a cache for the fused kernels.  The
    reference always works from the CURRENT ``measurand.val`` (measurand.py:502-505), so the cache must die
    when the value image is replaced, released (``val = None``) or modified in place (apply_thresholds)."""
    img = np.arange(4 * 5 * 3, dtype=np.uint8).reshape(4, 5, 3)
    with patch("cv2.imread", return_value=img):
        s = ImageSet(file_path="5ms BF a 10x.tif")
        s.load_value_image()
    assert s.dn is not None and torch.equal(s.dn, torch.from_numpy(img))
    assert torch.equal(s.digital_numbers(), torch.from_numpy(img))
    # in-place modification through the Measurand API
    s.measurand.apply_thresholds([0.1] * 3, [0.9] * 3)
    assert s.dn is None
    # re-loading gives a fresh, valid cache; assigning a new value image drops it again
    with patch("cv2.imread", return_value=img):
        s2 = ImageSet(file_path="5ms BF a 10x.tif")
        s2.load_value_image()
    s2.measurand.val = s2.measurand.val * 0.5
    assert s2.dn is None
    # releasing the image the reference's way releases the integers too
    with patch("cv2.imread", return_value=img):
        s3 = ImageSet(file_path="5ms BF a 10x.tif")
        s3.load_value_image()
    s3.measurand.val = None
    assert s3.dn is None
    with pytest.raises(ValueError):
        s3.digital_numbers()
    # in-memory integers attached without a value image stay valid until a value image appears
    s4 = ImageSet(features={"illumination": "bf", "magnification": "10x", "exposure": 0.01, "subject": "s"})
    s4.set_digital_numbers(img)
    assert s4.dn is not None and s4.measurand.val is None
    s4.measurand.val = torch.zeros((4, 5, 3), dtype=torch.float64)
    assert s4.dn is None
    # integer value image given to the constructor
    s5 = ImageSet(value=img)
    assert s5.dn is not None
    s5.measurand = Measurand(img.astype(np.float64) / 255)
    assert s5.dn is None

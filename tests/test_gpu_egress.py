"""8-bit export (ImageSet.save_8bit, image_set.py:321-358) on the GPU: byte-exact against the files the
unmodified reference wrote (goldens) and against the oracle on larger / awkward inputs."""
import numpy as np
import pytest
import torch

from oracle import egress
from gpu_util import dev, host

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402


def test_golden_reference_bytes(golden_dir):
    g = np.load(golden_dir / "k6_save_8bit.npz")
    for name in ("hdr", "unit", "ties", "negative"):
        for kind in ("val", "std"):
            got = ops.quantize_8bit(dev(g[f"{name}_{kind}"]))
            assert np.array_equal(host(got), g[f"{name}_{kind}_u8"]), (name, kind)


@pytest.mark.parametrize("shape", [(1,), (3,), (5, 7, 3), (129, 67, 3), (1080, 1920, 3), (1000003,)])
@pytest.mark.parametrize("scale", [0.7, 1.0, 41.0])
def test_matches_oracle_bytes(shape, scale):
    rng = np.random.default_rng(abs(hash((shape, scale))) % 2**32)
    val = rng.uniform(0, scale, shape)
    got, mx = ops.quantize_8bit(dev(val), 255.0, return_max=True)
    assert float(mx.cpu()) == val.max()
    assert np.array_equal(host(got), egress.quantize_8bit(val))


def test_unaligned_view_and_ties():
    base = torch.arange(0, 4099, dtype=torch.float64, device="cuda") / 510.0 * 3.0
    view = base[1:]                                  # 8-byte aligned only
    assert np.array_equal(host(ops.quantize_8bit(view)), egress.quantize_8bit(host(view)))


def test_nan_disables_normalisation():
    val = np.random.default_rng(3).uniform(0, 5, (64, 33, 3))
    val[7, 3, 1] = np.nan
    got, mx = ops.quantize_8bit(dev(val), return_max=True)
    assert np.isnan(float(mx.cpu()))                 # np.amax propagates NaN -> `max > 1` is False
    ref = egress.quantize_8bit(val)
    ok = ~np.isnan(val) & (val * 255 < 2**31)        # the cast of NaN / huge values is platform-defined
    assert np.array_equal(host(got)[ok], ref[ok])


def test_empty_raises():
    with pytest.raises(ValueError):
        ops.quantize_8bit(torch.empty((0, 3), dtype=torch.float64, device="cuda"))


def test_image_set_save_8bit_from_device(tmp_path, golden_dir):
    import cv2 as cv
    g = np.load(golden_dir / "k6_save_8bit.npz")
    s = cl.ImageSet(file_path=tmp_path / "hdr 5ms.tif", value=dev(g["hdr_val"]), std=dev(g["hdr_std"]))
    out = tmp_path / "o" / "hdr 5ms.tif"
    s.save_8bit(out, force_8_bit=True)
    assert np.array_equal(cv.imread(str(out), -1), g["hdr_val_u8"])
    assert np.array_equal(cv.imread(str(out).removesuffix(".tif") + " STD.tif", -1), g["hdr_std_u8"])


def test_tiff_roundtrip_8bit(tmp_path):
    # tests/integration/test_integration_image_set.py:48-83, 8-bit half
    rng = np.random.default_rng(1)
    val = rng.random((8, 9, 3))
    s = cl.ImageSet(file_path=tmp_path / "5ms BF a 10x.tif", value=dev(val), std=dev(val * 0.1))
    s.save_8bit(tmp_path / "8" / "5ms BF a 10x.tif")
    back = cl.ImageSet(file_path=tmp_path / "8" / "5ms BF a 10x.tif")
    back.load_value_image()
    assert np.allclose(host(back.measurand.val), val, atol=0.5 / 255 + 1e-12)


def test_save_8bit_all_golden_cases_through_image_set(tmp_path, golden_dir):
    import cv2 as cv
    g = np.load(golden_dir / "k6_save_8bit.npz")
    for name in ("hdr", "unit", "ties", "negative"):
        s = cl.ImageSet(file_path=tmp_path / f"{name} 5ms.tif", value=dev(g[f"{name}_val"]), std=dev(g[f"{name}_std"]))
        out = tmp_path / "out" / f"{name} 5ms.tif"
        s.save_8bit(out, force_8_bit=True)
        assert np.array_equal(cv.imread(str(out), -1), g[f"{name}_val_u8"])
        assert np.array_equal(cv.imread(str(out).removesuffix(".tif") + " STD.tif", -1), g[f"{name}_std_u8"])

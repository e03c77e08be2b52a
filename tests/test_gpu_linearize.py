"""K1 parity: CUDA linearize (through the C ABI) vs the oracle and the reference goldens.
LUT bins, gathered values and the std product are bit-exact."""
import numpy as np
import pytest
import torch

from oracle import linearize as ol
from gpu_util import dev, host, icrf_tables

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")


def _check(val, std, icrf, diff, max_dn=255):
    ev, es = ol.linearize(val, std, icrf, diff, max_dn)
    v, s, bins = ops.linearize(dev(val), dev(std), dev(icrf), dev(diff), max_dn, return_bins=True)
    assert np.array_equal(host(v), ev)
    if es is None:
        assert s is None
    else:
        assert np.array_equal(host(s), es)
    eb = ol.lut_index(val, max_dn)
    gb = host(bins)
    if gb.dtype == np.int16:
        gb = gb.view(np.uint16)
    assert np.array_equal(gb.astype(np.int64), eb.astype(np.int64))


def test_golden_vectors(golden_dir):
    g = np.load(golden_dir / "k1_linearize.npz")
    v, s = ops.linearize(dev(g["val"]), dev(g["std"]), dev(g["icrf"]), dev(g["icrf_diff"]))
    assert np.array_equal(host(v), g["exp_val"]) and np.array_equal(host(s), g["exp_std"])
    v, s = ops.linearize(dev(g["dn"]), dev(g["dn_std"]), dev(g["icrf"]), dev(g["icrf_diff"]))
    assert np.array_equal(host(v), g["exp_dn_val"]) and np.array_equal(host(s), g["exp_dn_std"])
    v, s = ops.linearize(dev(g["mono"]), dev(g["mono_std"]), dev(g["icrf"][:, 1].copy()), dev(g["icrf_diff"][:, 1].copy()))
    assert np.array_equal(host(v), g["exp_mono_val"]) and np.array_equal(host(s), g["exp_mono_std"])


@pytest.mark.parametrize("shape", [(1000, 777, 3), (33, 17, 3), (5, 1, 3), (64, 64, 1), (31, 9, 4), (7, 5, 2), (3,)])
@pytest.mark.parametrize("kind", ["u8", "f64"])
@pytest.mark.parametrize("with_std", [True, False])
def test_random_images(shape, kind, with_std):
    rng = np.random.default_rng(hash((shape, kind)) % 2**32)
    c = shape[-1]
    icrf, diff = icrf_tables(c if c > 1 else 0)
    if c == 1:
        pass
    val = rng.integers(0, 256, shape, dtype=np.uint8) if kind == "u8" else rng.random(shape)
    std = rng.uniform(1e-3, 2e-2, shape) if with_std else None
    _check(val, std, icrf, diff if with_std else None)


def test_uint16_65536_row_lut():
    rng = np.random.default_rng(5)
    icrf, diff = icrf_tables(0, bits=65536)
    val = rng.integers(0, 65536, (300, 257, 1), dtype=np.uint16)
    std = rng.uniform(1e-3, 2e-2, val.shape)
    ev, es = ol.linearize(val, std, icrf, diff, 65535)
    v, s = ops.linearize(dev(val.view(np.int16)).view(torch.uint16), dev(std), dev(icrf), dev(diff), 65535.0)
    assert np.array_equal(host(v), ev) and np.array_equal(host(s), es)
    fval = rng.random((50, 40, 1))
    _check(fval, None, icrf, None, 65535)


def test_wrapping_cast_and_half_even():
    icrf, _ = icrf_tables(0)
    x = np.array([256, 257, -1, np.nan, 300.4, 0.5, 1.5, 2.5, 254.5, 255.0, 1e30, -1e30, np.inf]) / 255.0
    _check(x.reshape(-1, 1), None, icrf, None)
    exact_ties = ((np.arange(255) + 0.5) / 255.0).reshape(-1, 1)
    _check(exact_ties, None, icrf, None)


def test_empty_and_errors():
    icrf, diff = icrf_tables(3)
    v, s = ops.linearize(dev(np.zeros((0, 4, 3), np.uint8)), None, dev(icrf), None)
    assert v.shape == (0, 4, 3) and s is None
    with pytest.raises(ValueError):
        ops.linearize(dev(np.zeros((4, 4, 2), np.uint8)), None, dev(icrf), None)
    with pytest.raises(RuntimeError):
        ops.linearize(torch.zeros((4, 4, 3), dtype=torch.uint8), None, torch.from_numpy(icrf), None)


def test_custom_op_registered():
    icrf, diff = icrf_tables(3)
    rng = np.random.default_rng(1)
    dn = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    out = torch.ops.camera_linearity.linearize(dev(dn), None, dev(icrf), None, 255.0)
    assert np.array_equal(host(out[0]), ol.linearize(dn, None, icrf)[0])

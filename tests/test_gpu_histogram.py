"""Per-channel histogram (compute_channel_histogram, measurand.py:430-469) on the GPU vs the unmodified
reference (goldens) and np.histogram: bin assignment is integer work -> counts bit-exact; inverse-sigma
weighted sums <= 1e-12 relative (atomic summation order)."""
import numpy as np
import pytest

from gpu_util import dev

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402

CASES = (("a", 8, (0.0, 1.0), False), ("b", 64, None, False), ("c", 50, (0.1, 0.7), True), ("d", 5000, None, True))


def test_golden_reference(golden_dir):
    g = np.load(golden_dir / "k7_histogram.npz")
    m = cl.Measurand(dev(g["val"]), dev(g["std"]))
    for tag, bins, rng_, use_std in CASES:
        h = m.compute_channel_histogram(bins, rng_, None, use_std)
        for c in range(3):
            assert np.array_equal(h[c][1], g[f"{tag}_edges_{c}"])
            if use_std:
                np.testing.assert_allclose(h[c][0], g[f"{tag}_hist_{c}"], rtol=1e-12, atol=0)
            else:
                assert h[c][0].dtype == g[f"{tag}_hist_{c}"].dtype
                assert np.array_equal(h[c][0], g[f"{tag}_hist_{c}"])


@pytest.mark.parametrize("bins", [1, 7, 256, 4096, 10000])
def test_random_against_numpy(bins):
    rng = np.random.default_rng(bins)
    val = rng.uniform(-1, 2, (300, 211, 3))
    val[rng.uniform(size=val.shape) < 0.01] = np.nan
    std = rng.uniform(0.01, 1, val.shape)
    std[rng.uniform(size=val.shape) < 0.01] = 0
    for c, use_std, r in ((0, False, None), (1, True, (-0.5, 1.5)), (2, False, (0.0, 1.0))):
        v = val[..., c]
        mask = np.isfinite(v)
        w = None
        if use_std:
            mask &= std[..., c] != 0
            w = 1 / std[..., c][mask]
        eh, ee = np.histogram(v[mask], bins=bins, range=r, weights=w)
        h, e = ops.channel_histogram(dev(val), dev(std) if use_std else None, c, bins, r)
        assert np.array_equal(e, ee)
        if use_std:
            np.testing.assert_allclose(h, eh, rtol=1e-12, atol=0)
        else:
            assert np.array_equal(h, eh)


def test_constant_channel_and_empty():
    val = np.full((10, 12, 2), 0.5)
    h, e = ops.channel_histogram(dev(val), None, 0, 10, None)
    eh, ee = np.histogram(val[..., 0], bins=10)
    assert np.array_equal(h, eh) and np.array_equal(e, ee)
    val[...] = np.nan
    h, e = ops.channel_histogram(dev(val), None, 1, 4, None)
    eh, ee = np.histogram(np.array([]), bins=4)
    assert np.array_equal(h, eh) and np.array_equal(e, ee)

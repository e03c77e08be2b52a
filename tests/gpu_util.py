"""Helpers shared by the GPU parity tests (everything goes through the C ABI via ops)."""
import numpy as np
import torch

DEV = "cuda"


def dev(x):
    if x is None:
        return None
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def host(t):
    return None if t is None else t.detach().cpu().numpy()


def icrf_tables(channels, bits=256, base=2.0, step=0.1):
    x = np.linspace(0, 1, bits)
    if channels == 0:
        icrf = x ** base
        return icrf, np.gradient(icrf, 2 / (bits - 1))
    icrf = np.stack([x ** (base + step * c) for c in range(channels)], axis=1)
    diff = np.stack([np.gradient(icrf[:, c], 2 / (bits - 1)) for c in range(channels)], axis=1)
    return icrf, diff


def synth_stack(rng, h, w, c, t, max_dn=255, dtype=np.uint8):
    rad = rng.uniform(0, 1, (h, w, c)) * 25
    dn = [np.rint(max_dn * np.clip(rad * tk, 0, 1) ** (1 / 2.2)).astype(dtype) for tk in t]
    std = [rng.uniform(0.002, 0.02, (h, w, c)) for _ in t]
    return dn, std


def assert_rel(actual, expected, rtol=1e-6):
    """The north-star tolerance: <= 1e-6 relative on float64 results (NaN/inf patterns must match)."""
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape
    np.testing.assert_array_equal(np.isnan(actual), np.isnan(expected))
    np.testing.assert_array_equal(np.isinf(actual), np.isinf(expected))
    fin = np.isfinite(expected)
    np.testing.assert_allclose(actual[fin], expected[fin], rtol=rtol, atol=0)


def max_rel(actual, expected):
    fin = np.isfinite(expected) & (expected != 0)
    return float(np.max(np.abs(actual[fin] - expected[fin]) / np.abs(expected[fin]))) if fin.any() else 0.0

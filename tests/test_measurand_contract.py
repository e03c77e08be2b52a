"""Constructor / operator contract of Measurand, following the reference's own tests
(tests/unit/test_measurand.py:120-444, 470-522) with torch tensors in place of NumPy arrays.
Runs on CPU tensors (the operator surface is device-agnostic torch code)."""
from copy import deepcopy

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from camera_linearity_b200 import Measurand, AbstractMeasurand, GlobalSettings
from camera_linearity_b200 import general_functions as gf



@pytest.fixture(autouse=True, scope="module")
def _cpu_tensors():
    """Host-logic tests run on CPU tensors; restore the default device afterwards."""
    previous = GlobalSettings.DEVICE
    GlobalSettings.DEVICE = "cpu"
    yield
    GlobalSettings.DEVICE = previous


@st.composite
def broadcastable_arrays(draw, max_dims=4, max_side=6):
    nd1, nd2 = draw(st.integers(1, max_dims)), draw(st.integers(1, max_dims))
    s1 = draw(st.lists(st.integers(1, max_side), min_size=nd1, max_size=nd1))
    s2 = draw(st.lists(st.integers(1, max_side), min_size=nd2, max_size=nd2))
    n = max(nd1, nd2)
    p1, p2 = [1] * (n - nd1) + s1, [1] * (n - nd2) + s2
    for i in range(n):
        if p1[i] != p2[i] and p1[i] != 1 and p2[i] != 1:
            if p1[i] > p2[i]:
                p2[i] = 1
            else:
                p1[i] = 1
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    a, b = rng.random(p1) + 0.05, rng.random(p2) + 0.05
    return a / a.max(), b / b.max()


@st.composite
def measurand_pairs(draw):
    a, b = draw(broadcastable_arrays())
    sa = draw(st.one_of(st.none(), st.just(a * 0.1)))
    sb = draw(st.one_of(st.none(), st.just(b * 0.1)))
    return Measurand(a, sa), Measurand(b, sb)


def close(x, y):
    return torch.allclose(x, y, atol=1e-8, equal_nan=True)


class TestInitialization:
    def test_scalar_becomes_float64_array(self):
        m = Measurand(10.0)
        assert isinstance(m.val, torch.Tensor) and m.val.shape == (1,) and m.val.dtype == torch.float64
        assert m.val == torch.tensor(10.0, dtype=torch.float64) and m.std is None
        m = Measurand(10.0, 1.0)
        assert m.std == torch.tensor(1.0, dtype=torch.float64)

    def test_arrays(self):
        v, s = np.array([10.0, 20.0]), np.array([1.0, 2.0])
        m = Measurand(v, s, use_cupy=False)
        assert np.array_equal(m.val.numpy(), v) and np.array_equal(m.std.numpy(), s)
        m = Measurand(torch.from_numpy(v))
        assert m.std is None

    def test_invalid_types(self):
        with pytest.raises(TypeError, match="Invalid value type"):
            Measurand("invalid_val", 1.0)
        with pytest.raises(TypeError, match="Invalid std type"):
            Measurand(10.0, "invalid_std")

    def test_shape_mismatch(self):
        with pytest.raises(ValueError, match="Value and std shapes must match."):
            Measurand(np.zeros((2, 3)), np.zeros((3, 2)))

    def test_setters_and_read_only_channels(self):
        m = Measurand(np.zeros((4, 5, 3)))
        with pytest.raises(TypeError, match="val must be an array or None"):
            m.val = "x"
        with pytest.raises(TypeError, match="std must be an array or None"):
            m.std = 3
        with pytest.raises(AttributeError):
            m.channels = 3
        m.val = None
        assert m.val is None

    def test_copy_and_deepcopy(self):
        m = Measurand(np.ones((2, 2)), np.ones((2, 2)))
        d = deepcopy(m)
        d.val[0, 0] = 5
        assert m.val[0, 0] == 1
        z = m.zeros_like_measurand()
        assert not z.val.any() and not z.std.any()


class TestArithmetic:
    @settings(deadline=None, max_examples=40)
    @given(measurand_pairs())
    def test_addition(self, ms):
        a, b = ms
        r1, r2 = a + b, b + a
        assert close(r1.val, r2.val)
        if a.std is not None or b.std is not None:
            assert close(r1.std, r2.std)
        else:
            assert r1.std is None
        ident = a + 0
        assert close(ident.val, a.val) and (a.std is None or close(ident.std, a.std))

    @settings(deadline=None, max_examples=40)
    @given(measurand_pairs())
    def test_subtraction(self, ms):
        a, b = ms
        assert close((a - b).val, -1 * (b - a).val)
        z = a - a
        assert close(z.val, torch.zeros_like(z.val))
        if a.std is not None:
            assert torch.all(z.std >= a.std)

    @settings(deadline=None, max_examples=40)
    @given(measurand_pairs())
    def test_division(self, ms):
        a, b = ms
        assert close((a / b).val, 1 / (b / a).val)
        c = deepcopy(a)
        assert close(((a + b) / c).val, (a / c + b / c).val)
        assert close((a / 1).val, a.val)
        assert close((a / a).val, torch.ones_like(a.val))
        inf = a / 0
        assert torch.isinf(inf.val).all()
        assert (inf.std is None) == (a.std is None)

    @settings(deadline=None, max_examples=40)
    @given(measurand_pairs())
    def test_multiplication(self, ms):
        a, b = ms
        r1, r2 = a * b, b * a
        assert close(r1.val, r2.val)
        if a.std is not None or b.std is not None:
            assert close(r1.std, r2.std)
        c = deepcopy(a)
        assert close((a * (b + c)).val, (a * b + a * c).val)
        zero = a * 0
        assert close(zero.val, torch.zeros_like(a.val))
        if a.std is not None:
            assert close(zero.std, torch.zeros_like(a.std))
        scaled = 2.0 * a                                   # __rmul__, image_set.py:260
        assert close(scaled.val, 2 * a.val)

    def test_power_and_logs(self):
        a = Measurand(np.array([1.0, 2.0, 3.0]), np.array([0.1, 0.1, 0.1]))
        p = a ** 2
        assert close(p.val, a.val ** 2) and close(p.std, 2 * a.val * a.std)
        assert close(a.log_e().val, torch.log(a.val))
        assert close(a.log_10().std, a.std / (a.val * np.log(10)))


class TestNormalizeInput:
    def test_cases(self):
        m1, m2 = Measurand(10.0), Measurand(20.0)
        other, use_std = m1._normalize_input(m2)
        assert other is m2 and use_std is False
        other, use_std = Measurand(10.0, 1.0)._normalize_input(20.0)
        assert isinstance(other, AbstractMeasurand) and other.val == 20.0 and other.std is None and use_std is True
        other, use_std = Measurand(10.0, 1.0)._normalize_input(np.array([1, 2, 3]))
        assert np.array_equal(other.val.numpy(), [1, 2, 3]) and use_std is True
        with pytest.raises(TypeError, match="Invalid other type."):
            m1._normalize_input("invalid_string")

    def test_not_broadcastable(self):
        with pytest.raises(ValueError, match="Measurands are not broadcastable."):
            Measurand(np.zeros((2, 3))) + Measurand(np.zeros((4, 5)))


class TestThresholds:
    @settings(deadline=None, max_examples=40)
    @given(measurand_pairs(), st.floats(0.25, 0.75), st.data())
    def test_regression_against_per_channel_loop(self, ms, threshold, data):
        m, _ = ms
        n = m.val.shape[-1]
        lower = data.draw(st.lists(st.one_of(st.floats(0.0, threshold), st.none()), min_size=n, max_size=n))
        upper = data.draw(st.lists(st.one_of(st.floats(threshold, 1.0), st.none()), min_size=n, max_size=n))
        expect_v = m.val.clone()
        expect_s = None if m.std is None else m.std.clone()
        for c in range(n):
            lo = -np.inf if lower[c] is None else lower[c]
            hi = np.inf if upper[c] is None else upper[c]
            mask = (expect_v[..., c] < lo) | (expect_v[..., c] > hi)
            expect_v[..., c][mask] = float("nan")
            if expect_s is not None:
                expect_s[..., c][mask] = float("nan")
        m.apply_thresholds(lower, upper)
        assert close(m.val, expect_v)
        if expect_s is not None:
            assert close(m.std, expect_s)

    def test_length_check(self):
        with pytest.raises(ValueError, match="must match the size of the independent axis"):
            Measurand(np.zeros((2, 3))).apply_thresholds([0.1], [0.2])


class TestStatistics:
    def test_compute_difference_and_statistics(self):
        rng = np.random.default_rng(0)
        x, y = rng.random((5, 6, 3)) + 0.1, rng.random((5, 6, 3)) + 0.1
        sx, sy = x * 0.05, y * 0.05
        a, r = Measurand.compute_difference(Measurand(x, sx), Measurand(y, sy), 0.5)
        assert np.allclose(a.val.numpy(), x - 0.5 * y) and np.allclose(r.val.numpy(), (x - 0.5 * y) / (0.5 * y))
        assert np.allclose(a.std.numpy(), np.sqrt(sx ** 2 + (0.5 * sy) ** 2))
        x[0, 0, 0] = np.nan
        stats = Measurand(x).compute_dimension_statistics(axis=(0, 1))
        assert np.allclose(stats["mean"].numpy(), np.nanmean(x, axis=(0, 1)))
        assert np.allclose(stats["std"].numpy(), np.nanstd(x, axis=(0, 1)))
        w = 1 / sx
        stats = Measurand(x, sx).compute_dimension_statistics(axis=(0, 1))
        mean = np.nansum(x * w, axis=(0, 1)) / np.nansum(w, axis=(0, 1))
        assert np.allclose(stats["mean"].numpy(), mean)
        assert np.allclose(stats["error"].numpy(), np.nanmean(sx, axis=(0, 1)))

    def test_extract_and_interpolate(self):
        x = np.arange(24.0).reshape(2, 4, 3)
        e = Measurand(x, x * 0.1).extract([0, 2], axis=-1)
        assert np.array_equal(e.val.numpy(), np.take(x, [0, 2], axis=-1))
        i = Measurand.interpolate(Measurand(x), Measurand(2 * x), 1.0, 3.0, 2.0)
        assert np.allclose(i.val.numpy(), 1.5 * x)


class TestGeneralFunctions:
    @settings(deadline=None, max_examples=60)
    @given(st.lists(st.integers(1, 5), min_size=1, max_size=4), st.lists(st.integers(1, 5), min_size=1, max_size=4))
    def test_is_broadcastable_matches_numpy(self, s1, s2):
        try:
            np.broadcast_shapes(tuple(s1), tuple(s2))
            expected = True
        except ValueError:
            expected = False
        assert gf.is_broadcastable(tuple(s1), tuple(s2)) == expected

    def test_empty_shape(self):
        with pytest.raises(ValueError, match="Shapes cannot be empty"):
            gf.is_broadcastable((), (1,))

    def test_hot_methods_fail_loudly_without_cuda(self):
        m = Measurand(np.random.rand(4, 4, 3))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.linearize(np.zeros((256, 3)))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.apply_gaussian_weight()


def _operator_results(M, g, to):
    a = M(to(g["a_val"]), to(g["a_std"]))
    b = M(to(g["b_val"]), to(g["b_std"]))
    k = M(to(g["k_val"]), None)
    return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b, "pow": a ** b, "neg": -a,
            "add_nostd": a + k, "mul_nostd": a * k, "div_scalar": a / 2.5, "rmul_scalar": 0.75 * a,
            "pow_scalar": a ** 2.2, "log_e": a.log_e(), "log_10": a.log_10(),
            "interp": M.interpolate(a, b, 0.01, 0.04, 0.025)}


def check_operator_goldens(results, g, rtol):
    for name, m in results.items():
        val, std = m.numpy()
        np.testing.assert_allclose(val, g[f"{name}_val"], rtol=rtol, atol=0, err_msg=name)
        if f"{name}_std" in g.files:
            np.testing.assert_allclose(std, g[f"{name}_std"], rtol=rtol, atol=0, err_msg=name)
        else:
            assert std is None, name


def test_operators_match_the_unmodified_reference(golden_dir):
    # measurand.py:106-279, 658-681 run unmodified by tests/golden/make_golden.py (k8_operators.npz)
    g = np.load(golden_dir / "k8_operators.npz")
    check_operator_goldens(_operator_results(Measurand, g, lambda x: x), g, 1e-13)

"""K3 parity: Welford mean/SEM frames.  uint8 mean bit-exact in every mode (including rounding
ties); streaming update state bit-identical float64; stack-mode float64 within 1e-6 (actual ~1e-13)."""
import numpy as np
import pytest
import torch

from oracle import welford as ow
from gpu_util import assert_rel, dev, host, icrf_tables

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402


def _video(rng, f, shape, noise=3):
    base = rng.integers(20, 231, (1,) + shape)
    return np.clip(base + np.rint(rng.normal(0, noise, (f,) + shape)), 0, 255).astype(np.uint8)


def test_golden_reference_vectors(golden_dir):
    g = np.load(golden_dir / "k3_welford.npz")
    for frames, mean_key in ((g["katw"], "katw_mean_u8"), (g["frames"], "mean_u8")):
        r = cl.welford_stack(frames)
        assert np.array_equal(host(r["mean"]), g[mean_key])
        assert not host(r["std"]).any()                                   # D13: literal uint8 SEM is all zeros
        o = ow.welford(list(frames))
        assert_rel(host(r["mean_f64"]), o["mean"], 1e-12)
        assert_rel(host(r["sem"]), o["sem"], 1e-9)
        r2 = cl.welford_algorithm("mem", None, True, frame_source=lambda p, fr=frames: list(fr) + [None])
        assert np.array_equal(host(r2["mean"]), g[mean_key])
        assert np.array_equal(host(r2["mean_f64"]), o["mean"])           # sequential recurrence: bit-identical
        assert np.array_equal(host(r2["sem"]), o["sem"])


@pytest.mark.parametrize("f,shape", [(12, (16, 16, 3)), (600, (24, 32, 3)), (37, (9, 7, 3)), (2, (32, 16, 3)),
                                     (64, (40, 48, 1)), (1000, (8, 8, 3))])
def test_stack_u8_bit_exact_mean_including_ties(f, shape):
    rng = np.random.default_rng(f)
    frames = _video(rng, f, shape, noise=1)          # small noise + even F -> many exact .5 ties
    o = ow.welford(list(frames))
    mean, sem, mean_u8 = ops.welford_stack(dev(frames))
    assert np.array_equal(host(mean_u8), o["mean_u8"])
    assert_rel(host(mean), o["mean"], 1e-12)
    if f > 1:
        np.testing.assert_allclose(host(sem), o["sem"], rtol=1e-9, atol=1e-16)
    else:
        assert np.isnan(host(sem)).all()


def test_adversarial_all_ties():
    # half the frames d, half d+1 -> mean*255 is exactly k+0.5 for every sample
    rng = np.random.default_rng(3)
    base = rng.integers(0, 255, (20, 16, 3), dtype=np.uint8)
    frames = np.stack([base + (i % 2) for i in range(10)]).astype(np.uint8)
    o = ow.welford(list(frames))
    _, _, mean_u8 = ops.welford_stack(dev(frames))
    assert np.array_equal(host(mean_u8), o["mean_u8"])
    # 960 ties: the replay runs one LANE per tie (one warp per tie below 256).  The same through the ICRF path with
    # the identity table d / 255, whose near-tie samples are replayed from the table in shared memory.
    ident = np.repeat((np.arange(256) / 255.0)[:, None], 3, axis=1)
    oi = ow.welford(list(frames), icrf=ident)
    _, _, mean_u8_i = ops.welford_stack(dev(frames), dev(ident))
    assert np.array_equal(host(mean_u8_i), oi["mean_u8"])
    few = frames[:, :4, :5]                          # 60 ties: one warp per tie
    _, _, mean_u8_f = ops.welford_stack(dev(np.ascontiguousarray(few)))
    assert np.array_equal(host(mean_u8_f), ow.welford(list(few))["mean_u8"])


def test_streaming_update_is_bit_identical():
    rng = np.random.default_rng(4)
    frames = _video(rng, 45, (13, 11, 3))
    o = ow.welford(list(frames))
    mean = torch.zeros((13, 11, 3), dtype=torch.float64, device="cuda")
    m2 = torch.zeros_like(mean)
    count = 0
    for lo in range(0, 45, 7):                      # ragged chunks
        count = ops.welford_update(dev(frames[lo:lo + 7]), mean, m2, count)
    assert count == 45
    assert np.array_equal(host(mean), o["mean"]) and np.array_equal(host(m2), o["m2"])
    sem, mean_u8 = ops.welford_finalize(mean, m2, count)
    assert np.array_equal(host(sem), o["sem"]) and np.array_equal(host(mean_u8), o["mean_u8"])


def test_with_icrf_linearisation():
    rng = np.random.default_rng(5)
    icrf, _ = icrf_tables(3)
    frames = _video(rng, 40, (12, 20, 3))
    o = ow.welford(list(frames), icrf=icrf)
    mean, sem, mean_u8 = ops.welford_stack(dev(frames), dev(icrf))
    assert np.array_equal(host(mean_u8), o["mean_u8"])
    assert_rel(host(mean), o["mean"], 1e-12)
    np.testing.assert_allclose(host(sem), o["sem"], rtol=1e-7, atol=1e-16)
    r = cl.welford_algorithm("mem", icrf, True, frame_source=lambda p: list(frames) + [None])
    assert np.array_equal(host(r["mean_f64"]), o["mean"]) and np.array_equal(host(r["sem"]), o["sem"])
    static = np.repeat(frames[:1], 9, axis=0)       # identical frames: SEM must be exactly 0
    _, sem0, _ = ops.welford_stack(dev(static), dev(icrf))
    assert not host(sem0).any()


def test_full_size_cfg4_checksum():
    """cfg4 size (600 x 1080x1920x3): exact integer identity sum(mean_u8 candidates) and oracle on a crop."""
    gen = torch.Generator(device="cuda").manual_seed(4)
    base = torch.randint(20, 231, (1, 1080, 1920, 3), generator=gen, device="cuda", dtype=torch.int16)
    frames = torch.empty((600, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    for f0 in range(0, 600, 50):
        noise = torch.round(torch.randn((50, 1080, 1920, 3), generator=gen, device="cuda") * 3).to(torch.int16)
        frames[f0:f0 + 50] = torch.clamp(base + noise, 0, 255).to(torch.uint8)
    del noise
    mean, sem, mean_u8 = ops.welford_stack(frames)
    # size-independent property: exact integer sums vs torch's own integer reduction on a sample of rows
    rows = [0, 511, 1079]
    for r in rows:
        s = frames[:, r].to(torch.int64).sum(dim=0)
        exact = s.to(torch.float64) / (600 * 255.0)
        assert float((mean[r] - exact).abs().max()) < 1e-15
        q, rem = s // 600, s % 600
        expect = torch.where(2 * rem > 600, q + 1, q)
        not_tie = 2 * rem != 600
        assert torch.equal(mean_u8[r][not_tie].to(torch.int64), expect[not_tie])
        # the ties of this row, if any (the sum of 600 noise terms of sigma 3 rarely reaches +-300: a handful in the
        # image), against the oracle's sequential recurrence
        w_idx, c_idx = torch.nonzero(~not_tie, as_tuple=True)
        if len(w_idx):
            sub = host(frames[:, r][:, w_idx, c_idx])                      # (600, K)
            o = ow.welford([f.reshape(-1, 1, 1) for f in sub])
            assert np.array_equal(host(mean_u8[r][w_idx, c_idx]), o["mean_u8"].reshape(-1))
    crop = host(frames[:, 500:502, 100:164])
    o = ow.welford(list(crop))
    assert np.array_equal(host(mean_u8[500:502, 100:164]), o["mean_u8"])
    np.testing.assert_allclose(host(sem[500:502, 100:164]), o["sem"], rtol=1e-9)


def test_welford_algorithm_pinned_double_buffer_many_chunks():
    # 5 buffer swaps + a ragged last chunk, two "videos": bit-identical to the sequential NumPy recurrence
    rng = np.random.default_rng(99)
    a = _video(rng, 3 * 32 + 5, (20, 24, 3))
    b = _video(rng, 2 * 32 + 17, (20, 24, 3))
    vids = {"a": a, "b": b}
    r = cl.welford_algorithm(["a", "b"], None, True, frame_source=lambda p: list(vids[p]) + [None])
    o = ow.welford(list(a) + list(b))
    assert r["count"] == len(a) + len(b)
    assert np.array_equal(host(r["mean_f64"]), o["mean"])
    assert np.array_equal(host(r["sem"]), o["sem"])


def test_icrf_variant_matches_the_unmodified_reference(golden_dir):
    # golden from the reference's own `if ICRF:` branch (always-true ndarray subclass, make_golden.py)
    g = np.load(golden_dir / "k3_welford_icrf.npz")
    r = cl.welford_stack(g["frames"], g["icrf"])
    assert np.array_equal(host(r["mean"]), g["mean_u8"])
    r2 = cl.welford_algorithm("mem", g["icrf"], True, frame_source=lambda p: list(g["frames"]) + [None])
    assert np.array_equal(host(r2["mean"]), g["mean_u8"])
    assert np.array_equal(host(r2["std"]), g["std_u8"])

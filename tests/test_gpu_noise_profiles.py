"""Camera noise profiles (joint histogram of mean DN vs frame DN per channel) on the GPU vs the unmodified reference
(golden k9) and the oracle: integer counts, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import noise_profiles as onp
from gpu_util import dev, host

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("camera_linearity_b200.ops")
import camera_linearity_b200 as cl  # noqa: E402


def test_golden_from_the_unmodified_reference(golden_dir):
    g = np.load(golden_dir / "k9_noise_profiles.npz")
    videos = {"0": list(g["frames0"]), "1": list(g["frames1"])}
    profiles, mean_frame = cl.compute_noise_profiles(["0", "1"], frame_source=lambda p: videos[str(p)] + [None])
    assert np.array_equal(host(mean_frame), g["mean_frame"])
    assert profiles.dtype == torch.int64 and np.array_equal(host(profiles), g["profiles"])


@pytest.mark.parametrize("shape,frames", [((37, 41, 3), 19), ((64, 64, 1), 40), ((9, 7, 3), 5), ((30, 50, 2), 12)])
def test_kernel_matches_the_oracle(shape, frames):
    """aligned and ragged sample counts, 1 / 2 / 3 channels, values far outside the +-8 DN shared-memory window"""
    rng = np.random.default_rng(sum(shape) + frames)
    base = rng.integers(0, 256, shape)
    video = [np.clip(base + np.rint(rng.normal(0, 4.0, shape)).astype(int), 0, 255).astype(np.uint8) for _ in range(frames)]
    video[frames // 2][0, 0] = 255 - video[frames // 2][0, 0]            # wild outliers
    exp, mean_frame = onp.noise_profiles([video])
    got = ops.noise_profiles(dev(np.stack(video)), dev(mean_frame))
    assert np.array_equal(host(got), exp)
    # accumulation over chunks == one call
    half = frames // 2
    acc = ops.noise_profiles(dev(np.stack(video[:half])), dev(mean_frame))
    acc = ops.noise_profiles(dev(np.stack(video[half:])), dev(mean_frame), acc)
    assert np.array_equal(host(acc), exp)


def test_cfg4_sized_video_counts_every_sample_frame():
    """1080 x 1920 x 3, 64 frames: the total count and the per-row marginals are exact identities"""
    g = torch.Generator(device="cuda").manual_seed(4)
    base = torch.randint(20, 231, (1, 1080, 1920, 3), generator=g, device="cuda", dtype=torch.int16)
    noise = torch.round(torch.randn((64, 1080, 1920, 3), generator=g, device="cuda") * 3).to(torch.int16)
    frames = torch.clamp(base + noise, 0, 255).to(torch.uint8)
    del noise
    mean_u8 = ops.welford_stack(frames)[2]
    hist = ops.noise_profiles(frames, mean_u8)
    assert int(hist.sum()) == frames.numel()
    # marginal over the frame DN = how often each mean value occurs, times the frame count
    for c in range(3):
        rows = hist[:, :, c].sum(dim=1)
        expect = torch.bincount(mean_u8[..., c].reshape(-1).to(torch.int64), minlength=256) * 64
        assert torch.equal(rows, expect)
    # marginal over the mean DN = histogram of the frame values
    cols = hist.sum(dim=0)
    for c in range(3):
        assert torch.equal(cols[:, c], torch.bincount(frames[..., c].reshape(-1).to(torch.int64), minlength=256))

"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is "reference" here:

* K4 ``_energy_function`` / ``analyze_linearity``, K3 ``welford_algorithm(ICRF=None)``,
  ``apply_gaussian_weight``, single-channel ``linearize`` and the Measurand operators run
  UNMODIFIED (only file IO is replaced by in-memory arrays).
* The HDR merge, multi-channel linearize, bad-pixel filter and flat-field correction raise at
  reference HEAD.  For them this script executes the reference's own loops
  (``ExposureSeries._precalculate_sum_of_weights`` / ``_compute_HDR_image_set`` /
  ``AbstractMeasurand.normalize_by_map``) with the minimal repair set R1..R8 applied as
  monkeypatches -- each patch below is labelled with the repair it implements.  No formula of
  the merge itself is restated here; lines ``exposure_series.py:382-394`` run as written.

The outputs are small ``.npz`` files committed next to this script; the oracle and the CUDA
path are both tested against them.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
from scipy.ndimage import median_filter

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent


# ----------------------------------------------------------------------------- helpers
def icrf_tables(channels: int, bits: int = 256, base: float = 1.8, step: float = 0.2):
    x = np.linspace(0, 1, bits)
    icrf = np.stack([x ** (base + step * c) for c in range(channels)], axis=1)
    diff = np.stack([np.gradient(icrf[:, c], 2 / (bits - 1)) for c in range(channels)], axis=1)
    return icrf, diff


class _IntSliceArray(np.ndarray):
    """Repair R7 / D10 only: ``flat_field[x0:x1, y0:y1, :]`` with float bounds -> int()."""

    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = tuple(slice(None if s.start is None else int(s.start),
                              None if s.stop is None else int(s.stop), s.step)
                        if isinstance(s, slice) else s for s in key)
        return np.asarray(super().__getitem__(key))


def install_repairs(ref, store):
    """Monkeypatch the repair set into the imported reference modules.

    ``store`` maps id(ImageSet) -> dict(val=uint8 image, std=float64 image) and replaces
    ``cv.imread``-based loading (file IO is out of scope on both sides).
    """
    M = ref.measurand.AbstractMeasurand
    NM = ref.measurand.NumpyMeasurand
    IS = ref.image_set.ImageSet
    gs = ref.gs

    # R1 (D1): multi-channel gather is ICRF[DN, arange(C)] (video_processing.py:201)
    def _linearize_channel(self, ICRF, ICRF_diff=None):
        use_std = self.std is not None and ICRF_diff is not None
        if not np.issubdtype(self.val.dtype, np.integer):
            iv = np.around(self.val * gs.MAX_DN).astype(np.dtype('uint8'))
        else:
            iv = self.val.copy()
        ch = np.arange(self.val.shape[-1])
        result = ICRF[iv, ch]
        if not use_std:
            return self.__class__(result, None)
        return self.__class__(result, ICRF_diff[iv, ch] * self.std)
    M._linearize_channel = _linearize_channel

    # R6 (D7, D8): call the median filter as a plain function; where(mask, median, original)
    def filter_larger_than_by_map(self, map, threshold_value):
        k = gs.MEDIAN_FILTER_KERNEL_SIZE
        large = map.val > threshold_value
        val = np.where(large, median_filter(self.val, size=(k, k), axes=(0, 1), mode='reflect'),
                       self.val)
        std = None
        if self.std is not None:
            std = np.where(large, median_filter(self.std, size=(k, k), axes=(0, 1), mode='reflect'),
                           self.std)
        return self.__class__(val, std)
    M.filter_larger_than_by_map = filter_larger_than_by_map

    # R5 (D6): the filtered measurand replaces the exposure's measurand
    orig_bpf = IS.bad_pixel_filter

    def bad_pixel_filter(self, darkSet, threshold_value=gs.DARK_THRESHOLD):
        filtered = orig_bpf(self, darkSet, threshold_value)
        self._measurand = filtered.measurand
        return filtered
    IS.bad_pixel_filter = bad_pixel_filter

    # R8 (D11): scale = target / original, no aliasing of the feature dict
    def scale_to_exposure(self, target_exp):
        feats = dict(self.features)
        scale = target_exp / feats['exposure']
        feats['exposure'] = target_exp
        return IS(file_path=self.path, features=feats, measurand=scale * self.measurand)
    IS.scale_to_exposure = scale_to_exposure

    # file IO -> in-memory (image_set.py:214-243)
    def load_value_image(self, bit64=False):
        img = store[id(self)]['val']
        self.measurand.val = img.astype(np.float64) / gs.MAX_DN if not bit64 else img
    IS.load_value_image = load_value_image

    def load_std_image(self, STD_data=None, bit64=False):
        self.measurand.std = store[id(self)]['std']
    IS.load_std_image = load_std_image
    return NM, IS


def run_reference_merge(ref, dn, std, t, icrf, icrf_diff, dark_dn=None, dark_t=None,
                        flat_dn=None, flat_std=None, config_roi=True):
    """Drive the reference's own merge loops (with repairs) on in-memory images."""
    store = {}
    NM, IS = install_repairs(ref, store)
    ES = ref.exposure_series.ExposureSeries

    def make(img, sd, exposure, subject='s'):
        s = IS(features={'illumination': 'bf', 'magnification': '10x', 'exposure': exposure,
                         'subject': subject}, measurand=NM(None, None))
        s.path = Path(f'/tmp/{subject} {exposure}.tif')
        store[id(s)] = {'val': img, 'std': sd}
        return s

    sets = [make(dn[k], std[k], float(t[k])) for k in range(len(dn))]
    darks = []
    if dark_dn is not None:
        darks = [make(dark_dn[k], None, float(dark_t[k]), 'dark') for k in range(len(dark_dn))]

    series = ES(input_image_sets=sets)
    sw, sw2 = series._precalculate_sum_of_weights(darks)
    # R3 (D4): plain arrays.  R4 (D5): accumulators need allocated val/std on the first set.
    sets[0].load_value_image()
    sets[0].load_std_image()
    hdr = series._compute_HDR_image_set(darks, sw.val, sw2.val, icrf, icrf_diff)
    val, sd = hdr.measurand.val, hdr.measurand.std
    if flat_dn is not None:                       # R7 (D9, D10)
        fv = (flat_dn.astype(np.float64) / ref.gs.MAX_DN).view(_IntSliceArray)
        fs = flat_std.view(_IntSliceArray)
        out = hdr.measurand.normalize_by_map(NM(fv, fs))
        val, sd = np.asarray(out.val), np.asarray(out.std)
    return np.asarray(val), np.asarray(sd)


def synth_stack(rng, h, w, c, t, max_dn=255):
    rad = rng.uniform(0, 1, (h, w, c)) * 25
    dn = [np.rint(max_dn * np.clip(rad * tk, 0, 1) ** (1 / 2.2)).astype(np.uint8) for tk in t]
    std = [rng.uniform(0.002, 0.02, (h, w, c)) for _ in t]
    return dn, std


# ----------------------------------------------------------------------------- K1
def golden_linearize(ref):
    NM = ref.measurand.NumpyMeasurand
    rng = np.random.default_rng(11)
    icrf, diff = icrf_tables(3)
    val = rng.uniform(0, 1, (12, 10, 3))
    val[0, :5, 0] = (np.arange(5) * 2 + 0.5) / 255       # exact .5 ties -> half-even
    val[1, :4, 1] = [0.0, 1.0, 0.5, 254.5 / 255]
    std = rng.uniform(0.001, 0.02, val.shape)
    exp_val = np.empty_like(val)
    exp_std = np.empty_like(val)
    for c in range(3):   # unmodified _linearize_single, one channel at a time == R1
        m = NM(val[..., c:c + 1].copy(), std[..., c:c + 1].copy()).linearize(icrf[:, c], diff[:, c])
        exp_val[..., c], exp_std[..., c] = m.val[..., 0], m.std[..., 0]
    dn = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    dn_std = rng.uniform(0.001, 0.02, dn.shape)
    exp_dn_val = np.empty(dn.shape)
    exp_dn_std = np.empty(dn.shape)
    for c in range(3):
        m = NM(dn[..., c:c + 1].copy(), dn_std[..., c:c + 1].copy()).linearize(icrf[:, c], diff[:, c])
        exp_dn_val[..., c], exp_dn_std[..., c] = m.val[..., 0], m.std[..., 0]
    mono = rng.uniform(0, 1, (7, 5, 1))
    mono_std = rng.uniform(0.001, 0.02, mono.shape)
    m = NM(mono.copy(), mono_std.copy()).linearize(icrf[:, 1], diff[:, 1])
    np.savez_compressed(OUT / 'k1_linearize.npz', icrf=icrf, icrf_diff=diff, val=val, std=std,
                        exp_val=exp_val, exp_std=exp_std, dn=dn, dn_std=dn_std,
                        exp_dn_val=exp_dn_val, exp_dn_std=exp_dn_std, mono=mono,
                        mono_std=mono_std, exp_mono_val=m.val, exp_mono_std=m.std)
    # Gaussian weight, unmodified
    v = np.arange(256) / 255.0
    w, dw = NM(v.reshape(-1, 1)).apply_gaussian_weight()
    np.savez_compressed(OUT / 'gaussian_weight.npz', v=v, w=w[:, 0], dw=dw[:, 0])


# ----------------------------------------------------------------------------- K2
def golden_merge():
    cases = {}
    # (a) survey KAT-M: no corrections, RNG-free
    ref = ref_loader.load({'image size x': 8, 'image size y': 6})
    hh, ww, cc = np.meshgrid(np.arange(6), np.arange(8), np.arange(3), indexing='ij')
    icrf, diff = icrf_tables(3)
    t = np.array([.005, .01, .02, .04, .08])
    dn = [np.clip(((29 * hh + 13 * ww + 7 * cc) % 200) * (k + 1) // 3, 0, 255).astype(np.uint8)
          for k in range(5)]
    std = [0.004 + 0.0005 * k + 1e-5 * ((hh + ww + cc) % 7) for k in range(5)]
    val, sd = run_reference_merge(ref, dn, std, t, icrf, diff)
    cases['katm'] = dict(dn=np.stack(dn), std=np.stack(std), t=t, icrf=icrf, icrf_diff=diff,
                         exp_val=val, exp_std=sd)

    # (b) random stack, darks (exact-exposure match + scaled longer dark) + flat, K = 3
    h, w = 30, 40
    ref = ref_loader.load({'image size x': h, 'image size y': w, 'dark threshold': 0.05,
                           'median filter kernel size': 3})
    rng = np.random.default_rng(21)
    t = 0.005 * 2.0 ** np.arange(6)                      # .005 .. .16 ; >= .05 gets a dark
    dn, std = synth_stack(rng, h, w, 3, t)
    icrf, diff = icrf_tables(3, base=2.0, step=0.1)
    dark_t = np.array([0.02, 0.08, 0.32])
    dark_dn = []
    for _ in dark_t:
        d = rng.poisson(2.0, (h, w, 3)).astype(np.uint8)
        hot = rng.uniform(size=d.shape) < 0.02
        d[hot] = rng.integers(40, 200, hot.sum())
        dark_dn.append(d)
    flat_dn = np.clip(np.rint(rng.normal(180, 6, (h, w, 3))), 1, 255).astype(np.uint8)
    flat_std = rng.uniform(0.001, 0.01, (h, w, 3))
    val, sd = run_reference_merge(ref, dn, std, t, icrf, diff, dark_dn, dark_t)
    valf, sdf = run_reference_merge(ref, dn, std, t, icrf, diff, dark_dn, dark_t, flat_dn, flat_std)
    cases['dark_flat'] = dict(dn=np.stack(dn), std=np.stack(std), t=t, icrf=icrf, icrf_diff=diff,
                              dark_dn=np.stack(dark_dn), dark_t=dark_t, flat_dn=flat_dn,
                              flat_std=flat_std, exp_val_dark=val, exp_std_dark=sd,
                              exp_val=valf, exp_std=sdf, dark_threshold=0.05, kernel=3,
                              im_size_x=h, im_size_y=w, ff_mid=0.2)

    # (c) K = 5 median, ragged sizes (W*C not a multiple of anything nice), 2 exposures
    h, w = 17, 13
    ref = ref_loader.load({'image size x': h, 'image size y': w, 'dark threshold': 0.01,
                           'median filter kernel size': 5})
    rng = np.random.default_rng(22)
    t = np.array([0.02, 0.5])
    dn, std = synth_stack(rng, h, w, 3, t)
    dark_t = np.array([0.02, 0.5])
    dark_dn = [rng.integers(0, 12, (h, w, 3), dtype=np.uint8) for _ in dark_t]
    val, sd = run_reference_merge(ref, dn, std, t, icrf, diff, dark_dn, dark_t)
    cases['k5'] = dict(dn=np.stack(dn), std=np.stack(std), t=t, icrf=icrf, icrf_diff=diff,
                       dark_dn=np.stack(dark_dn), dark_t=dark_t, exp_val=val, exp_std=sd,
                       dark_threshold=0.01, kernel=5)
    for name, d in cases.items():
        np.savez_compressed(OUT / f'k2_merge_{name}.npz', **d)


# ----------------------------------------------------------------------------- K3
def golden_welford(ref):
    vp = ref.video_processing

    class _Cap:
        def __init__(self, shape):
            self.shape = shape

        def get(self, prop):
            import cv2
            return self.shape[1] if prop == cv2.CAP_PROP_FRAME_WIDTH else self.shape[0]

    def run(frames):
        def gen(_path):
            for f in frames:
                yield f
            yield None
        vp.gf.video_frame_generator = gen
        vp.cv.VideoCapture = lambda p: _Cap(frames[0].shape)
        return vp.welford_algorithm(Path('/tmp/x.avi'), icrf_arg, True)

    icrf_arg = None

    hh, ww, cc = np.meshgrid(np.arange(6), np.arange(8), np.arange(3), indexing='ij')
    katw = [((31 * hh + 17 * ww + 5 * cc + 3 * f * f + f * hh) % 256).astype(np.uint8)
            for f in range(7)]
    r1 = run(katw)
    rng = np.random.default_rng(31)
    base = rng.integers(20, 231, (10, 12, 3))
    # even frame count, small noise -> many exact .5 ties in mean*255 (the hard case)
    fr = [np.clip(base + rng.integers(-1, 2, base.shape), 0, 255).astype(np.uint8) for _ in range(12)]
    r2 = run(fr)
    np.savez_compressed(OUT / 'k3_welford.npz', katw=np.stack(katw), katw_mean_u8=r1['mean'],
                        katw_std_u8=r1['std'], frames=np.stack(fr), mean_u8=r2['mean'],
                        std_u8=r2['std'])

    # The ICRF branch (`if ICRF:`, video_processing.py:200) raises for a plain ndarray (defect D13).  An
    # ndarray SUBCLASS whose truth value is True makes the UNMODIFIED function take that branch: the
    # linearised-frame Welford is pinned by the reference's own code, not only by the oracle.
    class TruthyArray(np.ndarray):
        def __bool__(self):
            return True

    icrf, _ = icrf_tables(3, base=2.0, step=0.15)
    icrf_arg = icrf.view(TruthyArray)
    fr3 = [np.clip(base + rng.integers(-3, 4, base.shape), 0, 255).astype(np.uint8) for _ in range(15)]
    r3 = run(fr3)
    r4 = run(katw)
    np.savez_compressed(OUT / 'k3_welford_icrf.npz', icrf=icrf, frames=np.stack(fr3), mean_u8=r3['mean'],
                        std_u8=r3['std'], katw_mean_u8=r4['mean'], katw_std_u8=r4['std'])


def golden_noise_profiles(ref):
    """UNMODIFIED ``compute_noise_profiles`` (video_processing.py:77-106): joint histogram of (uint8 mean frame DN,
    frame DN) per channel over every frame of the videos; only the frame source is replaced."""
    vp = ref.video_processing

    class _Cap:
        def __init__(self, shape):
            self.shape = shape

        def get(self, prop):
            import cv2
            return self.shape[1] if prop == cv2.CAP_PROP_FRAME_WIDTH else self.shape[0]

    rng = np.random.default_rng(91)
    base = rng.integers(0, 256, (24, 20, 3))
    base[:3] = 0                                   # clipped ends: rows whose noise leaves [0, 255]
    base[3:6] = 255
    videos = [[np.clip(base + np.rint(rng.normal(0, 2.5, base.shape)).astype(int), 0, 255).astype(np.uint8)
               for _ in range(n)] for n in (9, 6)]
    # one pixel with a wild outlier far from its mean (outside any small window around the mean)
    videos[0][4][10, 10, 1] = 255 if base[10, 10, 1] < 128 else 0

    def gen(path):
        for f in videos[int(Path(path).stem)]:
            yield f
        yield None
    vp.gf.video_frame_generator = gen
    vp.cv.VideoCapture = lambda p: _Cap(base.shape)
    profiles, mean_frame = vp.compute_noise_profiles([Path('/tmp/0.avi'), Path('/tmp/1.avi')])
    np.savez_compressed(OUT / 'k9_noise_profiles.npz', frames0=np.stack(videos[0]), frames1=np.stack(videos[1]),
                        profiles=profiles, mean_frame=mean_frame)


# ----------------------------------------------------------------------------- K4
def golden_energy(ref):
    ice = ref.ICRF_calibration_exposure
    rng = np.random.default_rng(41)
    x = np.linspace(0, 1, 256)
    mean = x ** 2.2
    modes = np.stack([np.sin((k + 1) * np.pi * x) * x for k in range(5)], axis=1)
    q, _ = np.linalg.qr(modes)
    pca = q
    t = 0.005 * 2.0 ** np.arange(5)
    rad = rng.uniform(0, 1, (24, 20, 1)) * 25
    dn = np.rint(255 * np.clip(rad * t[None, None, :], 0, 1) ** (1 / 2.2)).astype(np.uint8)
    dn = np.clip(dn.astype(int) + rng.integers(-2, 3, dn.shape), 0, 255).astype(np.uint8)
    std = rng.uniform(0.002, 0.02, dn.shape)
    std[3, 4, 2] = 0.0                                   # exercises the sigma != 0 gate
    params = np.concatenate([rng.uniform(-0.05, 0.05, (5, 12)),
                             rng.uniform(-2, 2, (5, 4))], axis=1)       # last 4 mostly gated
    e_nostd = np.array([ice._energy_function(params[:, s].copy(), mean.copy(), pca, dn, None,
                                             5, 250, True, t) for s in range(params.shape[1])])
    e_std = np.array([ice._energy_function(params[:, s].copy(), mean.copy(), pca, dn, std,
                                           5, 250, True, t) for s in range(params.shape[1])])
    # use_mean_ICRF = False branch: extra exponent parameter
    params6 = np.concatenate([rng.uniform(1.5, 3.0, (1, 6)), rng.uniform(-0.03, 0.03, (5, 6))])
    e6 = np.array([ice._energy_function(params6[:, s].copy(), mean.copy(), pca, dn, None,
                                        5, 250, False, t) for s in range(params6.shape[1])])
    # pair-level results of analyze_linearity for candidate 0
    curve = mean + pca @ params[:, 0]
    curve += 1 - curve[-1]
    curve[0] = 0
    pairs_nostd = ice.analyze_linearity(curve[dn], None, curve[5], curve[250], True, t)
    pairs_std = ice.analyze_linearity(curve[dn], std.copy(), curve[5], curve[250], True, t)
    np.savez_compressed(OUT / 'k4_energy.npz', mean=mean, pca=pca, t=t, dn=dn, std=std,
                        params=params, e_nostd=e_nostd, e_std=e_std, params6=params6, e6=e6,
                        pairs_nostd=pairs_nostd, pairs_std=pairs_std)


# ----------------------------------------------------------------------------- linearity chain
def golden_linearity(ref):
    # apply_thresholds -> ImageSet.compute_difference -> compute_dimension_statistics(axis=(0,1)),
    # all UNMODIFIED reference code (exposure_series.py:421-446)
    NM = ref.measurand.NumpyMeasurand
    M = ref.measurand.AbstractMeasurand
    rng = np.random.default_rng(51)
    icrf, _ = icrf_tables(3, base=2.0, step=0.1)
    t = (0.01, 0.04)
    rad = rng.uniform(0, 1, (36, 44, 3)) * 30
    dn = [np.rint(255 * np.clip(rad * tk, 0, 1) ** (1 / 2.2)).astype(np.uint8) for tk in t]
    vals = [icrf[d, np.arange(3)] * (1 + rng.normal(0, 0.01, d.shape)) for d in dn]
    stds = [rng.uniform(0.002, 0.02, d.shape) for d in dn]
    lower = [float(icrf[5, c]) for c in range(3)]
    upper = [float(icrf[250, c]) for c in range(3)]
    out = dict(x_val=vals[0], y_val=vals[1], x_std=stds[0], y_std=stds[1], lower=np.array(lower),
               upper=np.array(upper), multiplier=t[0] / t[1])
    for tag, use_std in (("std", True), ("nostd", False)):
        x = NM(vals[0].copy(), stds[0].copy() if use_std else None)
        y = NM(vals[1].copy(), stds[1].copy() if use_std else None)
        x.apply_thresholds(lower, upper)
        y.apply_thresholds(lower, upper)
        a, r = M.compute_difference(x, y, t[0] / t[1])
        for name, m in (("abs", a), ("rel", r)):
            st = m.compute_dimension_statistics(axis=(0, 1))
            out[f"{tag}_{name}_mean"] = st["mean"]
            out[f"{tag}_{name}_std"] = st["std"]
            if st["error"] is not None:
                out[f"{tag}_{name}_error"] = st["error"]
    np.savez_compressed(OUT / 'k5_linearity.npz', **out)


def golden_save_8bit(ref):
    # ImageSet.save_8bit (image_set.py:321-358), UNMODIFIED, written through OpenCV and read back
    import tempfile
    import cv2 as cv
    NM = ref.measurand.NumpyMeasurand
    rng = np.random.default_rng(77)
    out = {}
    cases = {
        "hdr": rng.uniform(0, 37.5, (40, 52, 3)),                       # max > 1: normalised
        "unit": rng.uniform(0, 1, (40, 52, 3)),                         # max <= 1: scaled only
        "ties": (np.arange(40 * 52 * 3).reshape(40, 52, 3) % 511) / 510.0,   # x.5 ties -> half-even
        "negative": rng.uniform(-0.4, 2.0, (40, 52, 3)),                # the C cast wraps negatives
    }
    with tempfile.TemporaryDirectory() as tmp:
        for name, val in cases.items():
            std = rng.uniform(0, 3.0, val.shape)
            path = Path(tmp) / f"{name} 5ms.tif"
            iset = ref.image_set.ImageSet(file_path=path, measurand=NM(val.copy(), std.copy()))
            iset.save_8bit(save_path=Path(tmp) / "out" / path.name, force_8_bit=True)
            out[f"{name}_val"] = val
            out[f"{name}_std"] = std
            out[f"{name}_val_u8"] = cv.imread(str(Path(tmp) / "out" / path.name), -1)
            out[f"{name}_std_u8"] = cv.imread(str(Path(tmp) / "out" / path.name).removesuffix('.tif') + ' STD.tif', -1)
    np.savez_compressed(OUT / 'k6_save_8bit.npz', **out)


def golden_histogram(ref):
    # AbstractMeasurand.compute_channel_histogram (measurand.py:430-469), UNMODIFIED
    NM = ref.measurand.NumpyMeasurand
    rng = np.random.default_rng(88)
    val = rng.normal(0.4, 0.2, (61, 47, 3))
    val[3, 5, 1] = np.nan
    val[7, 9, 2] = np.inf
    val[::7, ::5, 0] = 0.25                       # exact bin edges for bins=8 on (0, 1)
    std = rng.uniform(0.001, 0.02, val.shape)
    std[10, 10, :] = 0.0
    m = NM(val.copy(), std.copy())
    out = dict(val=val, std=std)
    for tag, bins, rng_, use_std in (("a", 8, (0.0, 1.0), False), ("b", 64, None, False), ("c", 50, (0.1, 0.7), True),
                                     ("d", 5000, None, True)):
        h = m.compute_channel_histogram(bins, rng_, None, use_std)
        for c in range(3):
            out[f"{tag}_hist_{c}"] = h[c][0]
            out[f"{tag}_edges_{c}"] = h[c][1]
    np.savez_compressed(OUT / 'k7_histogram.npz', **out)


def golden_operators(ref):
    # AbstractMeasurand operators with first-order propagation (measurand.py:106-279, 658-681), UNMODIFIED
    NM = ref.measurand.NumpyMeasurand
    M = ref.measurand.AbstractMeasurand
    rng = np.random.default_rng(99)
    shape = (11, 13, 3)
    a = NM(rng.uniform(0.2, 2.0, shape), rng.uniform(0.001, 0.05, shape))
    b = NM(rng.uniform(0.5, 3.0, shape), rng.uniform(0.001, 0.05, shape))
    k = NM(rng.uniform(0.5, 3.0, shape), None)
    out = dict(a_val=a.val, a_std=a.std, b_val=b.val, b_std=b.std, k_val=k.val)
    results = {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b, "pow": a ** b, "neg": -a,
               "add_nostd": a + k, "mul_nostd": a * k, "div_scalar": a / 2.5, "rmul_scalar": 0.75 * a,
               "pow_scalar": a ** 2.2, "log_e": a.log_e(), "log_10": a.log_10(),
               "interp": M.interpolate(a, b, 0.01, 0.04, 0.025)}
    for name, m in results.items():
        out[f"{name}_val"] = m.val
        if m.std is not None:
            out[f"{name}_std"] = m.std
    np.savez_compressed(OUT / 'k8_operators.npz', **out)


if __name__ == '__main__':
    import warnings
    warnings.simplefilter('ignore')
    ref = ref_loader.load()
    golden_linearize(ref)
    golden_welford(ref)
    golden_noise_profiles(ref)
    golden_energy(ref)
    golden_linearity(ref)
    golden_save_8bit(ref)
    golden_histogram(ref)
    golden_operators(ref)
    golden_merge()
    for f in sorted(OUT.glob('*.npz')):
        print(f.name, f.stat().st_size)

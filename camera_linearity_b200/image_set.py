"""``ImageSet``: one acquired image, its uncertainty image and the features parsed from its file
name.  Host-side shell around ``Measurand`` (reference: ``modules/image_set.py:25-572``).

File-name grammar (image_set.py:1-9, 542-568): space separated descriptors in any order --
``<exposure>ms``, ``bf``/``df`` illumination, ``<magnification>x`` and a subject; uncertainty
images carry an extra ``STD`` descriptor; ``.tif`` only.  File IO uses OpenCV on the host; the
decoded integer image is kept on the device as digital numbers (``.dn``) so that the fused
kernels never materialise ``dn / MAX_DN`` in HBM.
"""
from __future__ import annotations

import copy
import re
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from . import general_functions
from . import ops
from .measurand import Measurand
from .settings import GlobalSettings as gs


def _quantize_8bit(x) -> np.ndarray:
    """image_set.py:343-350: max-normalise (only when amax > 1), scale by MAX_DN, round half-even, uint8 --
    by the CUDA kernel; the bytes cross PCIe, not the float64 image.  No CPU fallback: host tensors raise."""
    from . import ops
    return ops.quantize_8bit(torch.as_tensor(x), gs.MAX_DN).cpu().numpy()


def _imread(path: str, flags=None):
    import cv2 as cv
    return cv.imread(path) if flags is None else cv.imread(path, flags)


class ImageSet(object):

    def __init__(self, file_path=None, value=None, std=None, features: Optional[Dict] = None,
                 measurand: Optional[Measurand] = None, use_cupy: Optional[bool] = False):
        self.path = Path(file_path) if isinstance(file_path, str) else file_path
        self._dn_cache = None
        self._dn_val = None            # (measurand.val tensor derived from the DNs, its version) or None
        if measurand is not None:
            self._measurand = measurand
        else:
            self._measurand = Measurand(value, std)
            v = self._measurand.val
            if v is not None and not v.dtype.is_floating_point:
                self._set_dn(v, v)
        self._use_cupy = True          # single (GPU) backend; kept as a read-only attribute
        if features is not None:
            self.features = features
        elif file_path is not None:
            self.features = _features_from_file_name(self.path)
        else:
            self.features = None
        self.is_HDR = False

    # ---- attributes (image_set.py:55-100) ----
    @property
    def measurand(self):
        return self._measurand

    @measurand.setter
    def measurand(self, new_measurand: Measurand):
        if not isinstance(new_measurand, Measurand):
            raise ValueError(f'Expected type {Measurand.backend}, got {type(new_measurand)} instead.')
        self._measurand = new_measurand
        self._dn_cache = None
        self._dn_val = None

    @property
    def use_cupy(self):
        return self._use_cupy

    @use_cupy.setter
    def use_cupy(self, _):
        raise AttributeError("use_cupy is a read-only attribute, managing the state of the used array backend.")

    def to_numpy(self):
        """image_set.py:88-93.  One backend: kept for call compatibility, changes nothing (``measurand.numpy()``
        returns host copies)."""

    def to_cupy(self):
        """image_set.py:95-100.  One backend (torch on the device): kept for call compatibility."""

    # The decoded integers are a cache of ``measurand.val``: they stay valid only while ``measurand.val`` is
    # still the very tensor that was derived from them (same object, not written to since).  Assigning
    # ``measurand.val`` (also to None, the reference's way of releasing an image) or mutating it in place
    # (``apply_thresholds``) drops them, so ``linearize`` / ``digital_numbers`` always follow the CURRENT value
    # image as the reference does (measurand.py:502-505).
    def _set_dn(self, dn, derived_val):
        self._dn_cache = dn
        self._dn_val = None if derived_val is None else (derived_val, derived_val._version)

    @property
    def _dn(self):
        if self._dn_cache is None:
            return None
        current = self._measurand.val
        if self._dn_val is None:
            ok = current is None
        else:
            ok = current is self._dn_val[0] and current._version == self._dn_val[1]
        if not ok:
            self._dn_cache = None
            self._dn_val = None
        return self._dn_cache

    @property
    def dn(self):
        """Integer digital numbers on the device, or None if only float data is held (or the value image
        has been replaced / modified since it was decoded)."""
        return self._dn

    def digital_numbers(self):
        """DN tensor for the fused kernels: the decoded integers, or ``rint(val * MAX_DN)`` with the
        reference's wrapping cast when only float values exist (measurand.py:503)."""
        if self._dn is not None:
            return self._dn
        v = self.measurand.val
        if v is None:
            raise ValueError("ImageSet holds no value image")
        if not v.dtype.is_floating_point:
            return v
        _, _, bins = ops.linearize(v, None, torch.zeros((gs.BITS, v.shape[-1]), dtype=torch.float64, device=v.device),
                                   None, gs.MAX_DN, return_bins=True)
        return bins.to(torch.uint8) if gs.MAX_DN <= 255 else bins

    # ---- forwarding maths (image_set.py:102-115, 200-212, 387-480) ----
    def linearize(self, ICRF, ICRF_diff=None):
        src = Measurand(self._dn, self.measurand.std) if self._dn is not None else self.measurand
        new_measurand = src.linearize(ICRF, ICRF_diff)
        return ImageSet(file_path=self.path, features=self.features, measurand=new_measurand)

    def extract(self, channels=None):
        new_measurand = self.measurand.extract(dims=channels, axis=-1)
        return ImageSet(file_path=self.path, features=self.features, measurand=new_measurand)

    def bad_pixel_filter(self, darkSet: 'ImageSet', threshold_value: Optional[float] = None):
        if threshold_value is None:
            threshold_value = gs.DARK_THRESHOLD
        new_measurand = self.measurand.filter_larger_than_by_map(darkSet.measurand, threshold_value)
        return ImageSet(file_path=self.path, measurand=new_measurand)

    def flat_field_correction(self, flatSet: 'ImageSet'):
        if flatSet.measurand.val is None:
            flatSet.load_value_image()
        if flatSet.measurand.std is None:
            flatSet.load_std_image()
        new_measurand = self.measurand.normalize_by_map(flatSet.measurand)
        return ImageSet(file_path=self.path, measurand=new_measurand)

    @staticmethod
    def compute_difference(short_exposure_set: 'ImageSet', long_exposure_set: 'ImageSet'):
        ratio = short_exposure_set.features["exposure"] / long_exposure_set.features["exposure"]
        abs_m, rel_m = Measurand.compute_difference(short_exposure_set.measurand,
                                                    long_exposure_set.measurand, ratio)
        return (ImageSet(file_path=short_exposure_set.path, features=short_exposure_set.features, measurand=abs_m),
                ImageSet(file_path=short_exposure_set.path, features=short_exposure_set.features, measurand=rel_m))

    @staticmethod
    def exposure_interpolation(short_exposure_set: 'ImageSet', long_exposure_set: 'ImageSet', exp: float):
        if not isinstance(exp, float):
            raise TypeError('Interpolation point has unsupported type.')
        exp0 = short_exposure_set.features['exposure']
        exp1 = long_exposure_set.features['exposure']
        if exp > exp1 or exp < exp0:
            raise ValueError('Interpolation point is not between the reference values.')
        new_measurand = Measurand.interpolate(short_exposure_set.measurand, long_exposure_set.measurand,
                                              exp0, exp1, exp)
        return ImageSet(features=short_exposure_set.features, measurand=new_measurand)

    # ---- names / matching (image_set.py:117-155) ----
    def get_file_path_without_exposure(self):
        if self.path is not None:
            return self.path.parent.joinpath(
                f"{self.features['subject']} {self.features['illumination']} {self.features['magnification']}.tif")
        return None

    def is_exposure_match(self, other: 'ImageSet'):
        if self.features is None or other.features is None:
            return False
        for key in self.features.keys():
            if key == "exposure":
                continue
            if self.features[key] != other.features[key]:
                return False
        return True

    def get_flat_field(self, list_of_flat_fields: Optional[List['ImageSet']] = None):
        if list_of_flat_fields is None:
            list_of_flat_fields = ImageSet.multiple_from_path(gs.DEFAULT_FLAT_PATH)
        for flat_set in list_of_flat_fields:
            if (self.features['illumination'] == flat_set.features['illumination']
                    and self.features['magnification'] == flat_set.features['magnification']):
                return flat_set
        return None

    def select_dark_field(self, list_of_dark_fields: List['ImageSet']):
        """Choice made by ``get_dark_field`` (image_set.py:171-198) without touching pixels:
        None, or (dark ImageSet, scale) with scale = target / original exposure (repair R8).
        Scan order, "last longer dark seen" and the early return are the reference's behaviour."""
        target = self.features['exposure']
        if not target >= gs.DARK_THRESHOLD:
            return None
        lesser = greater = False
        greater_index = 0
        for i, dark in enumerate(list_of_dark_fields):
            e = dark.features['exposure']
            if e < target:
                lesser = True
            if e > target:
                greater = True
                greater_index = i
            if target == e:
                return dark, 1.0
            if lesser and greater:
                chosen = list_of_dark_fields[greater_index]
                return chosen, target / chosen.features['exposure']
        return None

    def get_dark_field(self, list_of_dark_fields: Optional[List['ImageSet']] = None):
        if list_of_dark_fields is None:
            list_of_dark_fields = ImageSet.multiple_from_path(gs.DEFAULT_DARK_PATH)
        chosen = self.select_dark_field(list_of_dark_fields)
        if chosen is None:
            return None
        dark, scale = chosen
        if dark.measurand.val is None:
            dark.load_value_image()
        if scale == 1.0:
            return dark
        return dark.scale_to_exposure(self.features['exposure'])

    def scale_to_exposure(self, target_exp: float):
        """image_set.py:245-262 with repair R8 (no aliasing of the feature dict)."""
        new_features = dict(self.features)
        scale = target_exp / new_features['exposure']
        new_features['exposure'] = target_exp
        return ImageSet(file_path=self.path, features=new_features, measurand=scale * self.measurand)

    # ---- IO (image_set.py:214-243, 264-363) ----
    def set_digital_numbers(self, image):
        """Attach an in-memory integer image (what ``cv.imread`` would have returned)."""
        t = torch.as_tensor(np.ascontiguousarray(image) if isinstance(image, np.ndarray) else image)
        if t.dtype.is_floating_point:
            raise TypeError("digital numbers must be an integer image (uint8 / uint16)")
        self._measurand.val = None
        self._set_dn(t.to(gs.device(), non_blocking=True), None)

    def load_value_image(self, bit64: Optional[bool] = False):
        dn = self._dn
        if dn is None:
            if self.path is None:
                raise ValueError("ImageSet has neither a file path nor in-memory digital numbers")
            img = _imread(str(self.path)) if not bit64 else _imread(str(self.path), -1)
            if img is None:
                raise FileNotFoundError(str(self.path))
            dn = torch.from_numpy(np.ascontiguousarray(img)).to(gs.device())
        if not bit64 and not dn.dtype.is_floating_point:
            self._measurand.val = dn.to(torch.float64) / gs.MAX_DN        # image_set.py:223
        else:
            self._measurand.val = dn                                       # :225
        if dn.dtype.is_floating_point:
            self._dn_cache = None
            self._dn_val = None
        else:
            self._set_dn(dn, self._measurand.val)

    def std_file_exists(self) -> bool:
        """True when the '<name> STD.tif' uncertainty image of image_set.py:236 exists."""
        return self.path is not None and Path(str(self.path).removesuffix('.tif') + ' STD.tif').exists()

    def load_std_image(self, STD_data=None, bit64: Optional[bool] = False):
        std_array = None
        if self.path is not None:
            std_path = str(self.path).removesuffix('.tif') + ' STD.tif'
            if Path(std_path).exists():
                img = _imread(std_path, -1)
                if img is not None:
                    std_array = torch.from_numpy(np.ascontiguousarray(img)).to(gs.device())
        if std_array is None:
            std_array = self.calculate_numerical_STD(STD_data)
        if std_array is None:
            return
        self._measurand.std = std_array

    def calculate_numerical_STD(self, STD_data=None):
        """Uncertainty from the camera's STD table: ``STD_data[DN, c]`` (image_set.py:365-385)."""
        if STD_data is None:
            try:
                STD_data = general_functions.read_txt_to_array(gs.STD_FILE_NAME)
            except (FileNotFoundError, OSError):
                print('Could not load STD data for numerical estimation.')
                return None
        src = Measurand(self._dn, None) if self._dn is not None else self.measurand
        return src.linearize(ICRF=STD_data).val

    def save_64bit(self, save_path: Optional[Path] = None, is_HDR: Optional[bool] = False,
                   separate_channels: Optional[bool] = False):
        import cv2 as cv
        file_path = self.path.parent.joinpath('64bit', self.path.name) if save_path is None else save_path
        file_path.parent.mkdir(parents=True, exist_ok=True)
        file_path = str(file_path)
        acq_suffix, std_suffix = (' HDR.tif', ' HDR STD.tif') if is_HDR else ('.tif', ' STD.tif')
        val, std = self.measurand.numpy()
        stem = file_path.removesuffix('.tif')
        if not separate_channels:
            cv.imwrite(stem + acq_suffix, val.astype(np.float64))
            if std is not None:
                cv.imwrite(stem + std_suffix, std.astype(np.float64))
        else:
            for c in range(gs.NUM_OF_CHS):
                cv.imwrite(stem + acq_suffix.replace('.tif', f' {gs.CH_NAMES[c]}.tif'), val[:, :, c])
                if std is not None:
                    cv.imwrite(stem + std_suffix.replace('.tif', f' {gs.CH_NAMES[c]}.tif'), std[:, :, c])

    def save_8bit(self, save_path: Optional[Path] = None, force_8_bit: Optional[bool] = False):
        import cv2 as cv
        file_path = self.path.parent.joinpath('8bit', self.path.name) if save_path is None else save_path
        file_path.parent.mkdir(parents=True, exist_ok=True)
        file_path = str(file_path)
        cv.imwrite(file_path, _quantize_8bit(self.measurand.val))
        if self.measurand.std is not None:
            if force_8_bit:
                std = _quantize_8bit(self.measurand.std)
            else:
                std = self.measurand.numpy()[1]
            cv.imwrite(file_path.removesuffix('.tif') + ' STD.tif', std)

    @classmethod
    def multiple_from_path(cls, path: Path):
        """ImageSets for every ``*.tif`` under path whose name has no 'STD' (image_set.py:482-501)."""
        return [cls(file_path=f) for f in Path(path).glob("*.tif") if "STD" not in f.name]


def _features_from_file_name(file_path: Path):
    """image_set.py:542-568."""
    features = {"illumination": "", "magnification": "", "exposure": 0.0, "subject": ""}
    for element in file_path.name.removesuffix('.tif').split():
        if element.casefold() in ('bf', 'df'):
            features["illumination"] = element
        elif re.match("^[0-9]+.*[xX]$", element):
            features["magnification"] = element
        elif re.match("^[0-9]+.*ms$", element):
            features["exposure"] = float(element.removesuffix('ms')) / 1000
        else:
            features["subject"] = element
    return features


def calibrate_flats():
    """Bias subtraction of raw flat fields (image_set.py:504-521)."""
    darks = ImageSet.multiple_from_path(gs.DEFAULT_DARK_PATH)
    darks.sort(key=lambda s: s.features['exposure'])
    bias = darks[0]
    bias.load_value_image()
    bias.load_std_image()
    for flat in ImageSet.multiple_from_path(gs.UNCALIBRATED_FLAT_PATH):
        flat.load_value_image()
        flat.load_std_image()
        flat.measurand = flat.measurand - bias.measurand
        flat.save_8bit(gs.DEFAULT_FLAT_PATH)


def calibrate_dark_frames():
    """Bias subtraction of raw dark frames (image_set.py:524-539)."""
    darks = ImageSet.multiple_from_path(gs.UNCALIBRATED_DARK_PATH)
    darks.sort(key=lambda s: s.features['exposure'])
    bias = darks[0]
    bias.load_value_image()
    bias.load_std_image()
    for dark in darks:
        dark.load_value_image()
        dark.load_std_image()
        dark.measurand = dark.measurand - bias.measurand
        dark.save_8bit(gs.DEFAULT_DARK_PATH)

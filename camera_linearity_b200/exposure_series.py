"""``ExposureSeries`` / ``ExposurePair``: exposure stacks, HDR merging and linearity analysis.

Host-side shell (reference: ``modules/exposure_series.py:18-499``).  ``process_HDR_image`` replaces
the reference's two streaming NumPy passes (``_precalculate_sum_of_weights`` :317-345 and
``_compute_HDR_image_set`` :347-397, ~25 full-image float64 temporaries per exposure) with ONE
fused kernel launch over the whole stack (``ops.hdr_merge``), including the bad-pixel repair
(repairs R5/R6) and the flat-field epilogue (R7).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from . import general_functions as gf
from . import ops
from .image_set import ImageSet
from .measurand import Measurand, _flat_roi
from .settings import GlobalSettings as gs


class ExposurePair(object):
    """Two ImageSets of one series and their difference statistics (exposure_series.py:18-76)."""

    def __init__(self, short_exposure: ImageSet, long_exposure: ImageSet):
        self.short_exposure = short_exposure
        self.long_exposure = long_exposure
        self.exposure_ratio = short_exposure.features["exposure"] / long_exposure.features["exposure"]
        self.absolute_difference = None
        self.relative_difference = None
        self.absolute_stats = None
        self.relative_stats = None

    def compute_difference(self):
        self.absolute_difference, self.relative_difference = (
            ImageSet.compute_difference(self.short_exposure, self.long_exposure))

    def compute_stats(self, axis=None, release_memory_after: Optional[bool] = True):
        self.absolute_stats = self.absolute_difference.measurand.compute_dimension_statistics(axis=axis)
        self.relative_stats = self.relative_difference.measurand.compute_dimension_statistics(axis=axis)
        if release_memory_after:
            self.absolute_difference = None
            self.relative_difference = None

    def compute_pair_statistics(self, lower=None, upper=None):
        """Fused GPU path of ``compute_difference`` + ``compute_stats(axis=(0, 1))`` (optionally with the
        thresholds of ``apply_thresholds``): one kernel sequence, no difference images in HBM."""
        x, y = self.short_exposure.measurand, self.long_exposure.measurand
        stats = ops.pair_statistics(x.val, x.std, y.val, y.std, self.exposure_ratio, lower, upper)
        use_std = x.std is not None or y.std is not None
        self.absolute_stats = {"mean": stats[0, 0], "std": stats[0, 1], "error": stats[0, 2] if use_std else None}
        self.relative_stats = {"mean": stats[1, 0], "std": stats[1, 1], "error": stats[1, 2] if use_std else None}

    def process_linearity_distribution(self, bins: int, included_range=None, channels=None, use_std=False):
        return (self.absolute_difference.measurand.compute_channel_histogram(bins, included_range, channels, use_std),
                self.relative_difference.measurand.compute_channel_histogram(bins, included_range, channels, use_std))


class ExposureSeries(object):
    """ImageSets that differ only in exposure time (exposure_series.py:79-476)."""

    def __init__(self, merged_image_set: Optional[ImageSet] = None, directory_path: Optional[Path] = None,
                 input_image_sets: Optional[List[ImageSet]] = None, use_cupy: Optional[bool] = True):
        self.merged_image_set = merged_image_set
        self.input_image_sets = input_image_sets if input_image_sets is not None else []
        if isinstance(directory_path, Path) and directory_path.suffix != "":
            self.directory_path = directory_path.parent
        else:
            self.directory_path = directory_path
        self.exposure_pairs = None
        self._use_cupy = use_cupy if not input_image_sets else input_image_sets[0].use_cupy

    @property
    def use_cupy(self):
        return self._use_cupy

    @use_cupy.setter
    def use_cupy(self, _):
        raise AttributeError("use_cupy is a read-only attribute, managing the state of the used array backend.")

    # ---- construction (exposure_series.py:117-203) ----
    @classmethod
    def from_image_set(cls, reference_image_set: ImageSet, directory_path: Optional[Path] = None):
        search_path = reference_image_set.path.parent if directory_path is None else directory_path
        matches = [s for s in ImageSet.multiple_from_path(search_path) if reference_image_set.is_exposure_match(s)]
        matches.sort(key=lambda s: s.features["exposure"])
        return cls(directory_path=search_path, input_image_sets=matches)

    @classmethod
    def from_dir_path(cls, directory_path: Path):
        return ExposureSeries.from_multiple_image_sets(ImageSet.multiple_from_path(directory_path))

    @classmethod
    def from_multiple_image_sets(cls, list_of_image_sets: List[ImageSet]):
        groups: List[List[ImageSet]] = []
        for image_set in list_of_image_sets:
            for group in groups:
                if group[0].is_exposure_match(image_set):
                    group.append(image_set)
                    break
            else:
                groups.append([image_set])
        series = []
        for group in groups:
            group.sort(key=lambda s: s.features['exposure'])
            series.append(cls(input_image_sets=group))
        return series

    def load_value_images(self, bit_64: Optional[bool] = False):
        for image_set in self.input_image_sets:
            image_set.load_value_image(bit64=bit_64)

    def load_std_images(self, bit_64: Optional[bool] = False):
        for image_set in self.input_image_sets:
            image_set.load_std_image(bit64=bit_64)

    # ---- per-image operations (exposure_series.py:226-281) ----
    def linearize(self, ICRF, ICRF_diff=None, release_memory: Optional[bool] = False):
        new_sets = []
        for image_set in self.input_image_sets or []:
            new_sets.append(image_set.linearize(ICRF, ICRF_diff))
            if release_memory:
                image_set.measurand.val = None
                image_set.measurand.std = None
        return ExposureSeries(merged_image_set=self.merged_image_set, directory_path=self.directory_path,
                              input_image_sets=new_sets)

    def extract(self, channels=None, release_memory: Optional[bool] = False):
        merged = self.merged_image_set.extract(channels) if self.merged_image_set is not None else None
        new_sets = []
        for image_set in self.input_image_sets or []:
            new_sets.append(image_set.extract(channels))
            if release_memory:
                image_set.measurand.val = None
                image_set.measurand.std = None
        return ExposureSeries(merged_image_set=merged, directory_path=self.directory_path, input_image_sets=new_sets)

    def initialize_exposure_pairs(self):
        pairs = []
        for i, x in enumerate(self.input_image_sets):
            for j, y in enumerate(self.input_image_sets):
                if i >= j:
                    continue
                if x.features["exposure"] / y.features["exposure"] < 0.1:
                    continue
                pairs.append(ExposurePair(x, y))
        self.exposure_pairs = pairs

    # ---- HDR merge (exposure_series.py:317-419) ----
    def process_HDR_image(self, ICRF=None, ICRF_diff=None, dark_list: Optional[List[ImageSet]] = None,
                          flat_list: Optional[List[ImageSet]] = None, STD_data=None, algo: int = 0):
        """Merge the input exposures into ``self.merged_image_set``.

        ICRF / ICRF_diff: (BITS, C) tables; by default the calibrated ICRF file is read and the
        derivative is ``gradient(ICRF, 2/(BITS-1))`` (repair R2).  dark_list / flat_list default to
        the images under ``gs.DEFAULT_DARK_PATH`` / ``gs.DEFAULT_FLAT_PATH`` as in the reference.
        """
        if not self.input_image_sets:
            raise ValueError("ExposureSeries has no input images")
        if ICRF is None:
            ICRF = gf.read_txt_to_array(gs.ICRF_CALIBRATED_FILE)
        if ICRF_diff is None:
            ICRF_diff = gf.icrf_derivative(ICRF)
        if dark_list is None:
            dark_list = ImageSet.multiple_from_path(gs.DEFAULT_DARK_PATH) if Path(gs.DEFAULT_DARK_PATH).is_dir() else []

        sets = self.input_image_sets
        dn, std, exposures, darks, scales = [], [], [], [], []
        std_lut = None
        for image_set in sets:
            if image_set.dn is None and image_set.measurand.val is None:
                image_set.load_value_image()
            if image_set.measurand.std is None and image_set.std_file_exists():
                image_set.load_std_image(STD_data)          # a '... STD.tif' next to the image wins (image_set.py:228-243)
            # otherwise the uncertainty is STD_data[DN, c] (image_set.py:365-385): the merge kernel gathers it
            # from the table itself, so no float64 uncertainty image is built, uploaded or streamed
            d = image_set.digital_numbers()
            dn.append(d)
            std.append(image_set.measurand.std)
            exposures.append(float(image_set.features['exposure']))
            chosen = image_set.select_dark_field(dark_list) if dark_list else None
            if chosen is None:
                darks.append(None)
                scales.append(1.0)
            else:
                dark_set, scale = chosen
                if dark_set.dn is None and dark_set.measurand.val is None:
                    dark_set.load_value_image()
                darks.append(dark_set.digital_numbers())
                scales.append(scale)
        dev = dn[0].device
        if any(s is None for s in std):
            if STD_data is None:
                try:
                    STD_data = gf.read_txt_to_array(gs.STD_FILE_NAME)      # image_set.py:377
                except (FileNotFoundError, OSError, TypeError):
                    raise ValueError("an exposure has no uncertainty image and no STD_data table was given") from None
            std_lut = torch.as_tensor(STD_data, dtype=torch.float64, device=dev)

        reference_set = sets[0]
        if flat_list is None:
            flat_list = ImageSet.multiple_from_path(gs.DEFAULT_FLAT_PATH) if Path(gs.DEFAULT_FLAT_PATH).is_dir() else []
        flat_set = reference_set.get_flat_field(flat_list) if reference_set.features is not None and flat_list else None
        flat = flat_std = flat_means = None
        if flat_set is not None:
            if flat_set.dn is None and flat_set.measurand.val is None:
                flat_set.load_value_image()
            if flat_set.measurand.std is None:
                flat_set.load_std_image(STD_data)
            flat = flat_set.dn if flat_set.dn is not None else flat_set.measurand.val
            flat_std = flat_set.measurand.std
            flat_means = ops.flat_roi_means(flat, flat_std, _flat_roi(), gs.MAX_DN)

        val, sd = ops.hdr_merge(dn, std, exposures, torch.as_tensor(ICRF, device=dev),
                                torch.as_tensor(ICRF_diff, device=dev), std_lut=std_lut, darks=darks,
                                dark_scales=scales, dark_threshold=gs.DARK_THRESHOLD,
                                median_kernel=gs.MEDIAN_FILTER_KERNEL_SIZE, flat=flat, flat_std=flat_std,
                                flat_means=flat_means, algo=algo)
        hdr_path = reference_set.get_file_path_without_exposure() if reference_set.features is not None else None
        hdr = ImageSet(file_path=hdr_path, features=reference_set.features, measurand=Measurand(val, sd))
        hdr.is_HDR = True
        self.merged_image_set = hdr

    # ---- linearity analysis (exposure_series.py:421-476) ----
    def process_linearity(self, ICRF, linearity_limit: Optional[int] = None, use_std: Optional[bool] = False):
        lower, upper = gf.map_linearity_limits(linearity_limit, linearity_limit, ICRF)
        for image_set in self.input_image_sets:
            if image_set.measurand.val is None:
                image_set.load_value_image()
            if image_set.measurand.std is None and use_std:
                image_set.load_std_image()
            image_set.measurand.apply_thresholds(lower, upper)
        fused = all(s.measurand.val.is_cuda and s.measurand.val.ndim == 3 for s in self.input_image_sets)
        for pair in self.exposure_pairs:
            if fused:
                pair.compute_pair_statistics()         # the inputs are already thresholded (NaN-marked)
            else:
                pair.compute_difference()
                pair.compute_stats(axis=(0, 1), release_memory_after=True)

    def collect_exposure_pair_stats(self, return_cupy: Optional[bool] = False):
        keys = ('ratios', 'means', 'stds', 'errors')
        relative = {k: [] for k in keys}
        absolute = {k: [] for k in keys}
        for pair in self.exposure_pairs:
            for store, stats in ((absolute, pair.absolute_stats), (relative, pair.relative_stats)):
                store['ratios'].append(pair.exposure_ratio)
                store['means'].append(stats['mean'])
                store['stds'].append(stats['std'])
                store['errors'].append(stats['error'])
        return _to_2d_array(absolute), _to_2d_array(relative)


def _to_2d_array(dictionary: Dict):
    def host(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x
    return {k: np.array([host(x) for x in v]) for k, v in dictionary.items()}

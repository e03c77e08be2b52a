"""camera_linearity_b200 -- B200-native hot path of samivout/camera_linearity.

ICRF linearisation (K1), fused weighted HDR merge (K2), Welford mean/std frame stacking (K3) and
the ICRF calibration objective for a whole DE population (K4) as hand-written sm_100a CUDA kernels
behind a C ABI (``include/camera_linearity.h``), driven through the reference's own
``Measurand`` / ``ImageSet`` / ``ExposureSeries`` API.  No NumPy/CuPy dispatch, no CPU fallback.
"""
from .settings import GlobalSettings
from .measurand import Measurand as _MeasurandClass
from .measurand import AbstractMeasurand, NumpyMeasurand, MeasurandFactory, measurand_to_numpy, measurand_to_cupy
from .image_set import ImageSet
from .exposure_series import ExposureSeries, ExposurePair
from . import general_functions, ops, parallel, video_processing, ICRF_calibration_exposure
from .video_processing import compute_noise_profiles, welford_algorithm, welford_stack
from .ICRF_calibration_exposure import _energy_function, EnergyEvaluator, calibration

# `Measurand(val, std, use_cupy=...)` is both the reference's factory call and the class
Measurand = _MeasurandClass

__all__ = ["GlobalSettings", "Measurand", "AbstractMeasurand", "NumpyMeasurand", "MeasurandFactory", "measurand_to_numpy", "measurand_to_cupy",
           "ImageSet", "ExposureSeries", "ExposurePair", "welford_algorithm", "welford_stack", "compute_noise_profiles",
           "_energy_function", "EnergyEvaluator", "calibration", "ops", "parallel"]

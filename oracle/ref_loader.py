"""Load the *actual* reference (samivout/camera_linearity) modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  This file is used (a) by ``tests/golden/make_golden.py`` to
generate the committed golden vectors and (b) by the optional ``reference``-marked tests
that pin the numpy oracle to the reference bit-for-bit.  It only works in the build
container, where ``/root/reference`` is mounted read-only; on the GPU box it reports
``available() == False`` and nothing depends on it.

The reference cannot be imported as shipped: ``modules/global_settings.py:16-18`` evaluates
``read_config`` at import time and ``data/config.ini`` is git-ignored.  Instead of copying
sources we inject a stub ``read_config`` module into ``sys.modules`` (the reference does
``import read_config as rd``, ``global_settings.py:5``) that serves the keys from a dict.
No reference source is copied or modified.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")
REFERENCE_MODULES = REFERENCE_ROOT / "modules"

# Keys follow modules/global_settings.py:9-81.  Values we had to choose (the reference's own
# config.ini is not shipped) are the defaults documented in DESIGN.md.
DEFAULT_CONFIG = {
    "image size x": 2048,
    "image size y": 1536,
    "channels": 3,
    "bit depth": 8,
    "final datapoints": 256,
    "datapoint multiplier": 1,
    "original DoRF datapoints": 1024,
    "number of principal components": 5,
    "median filter kernel size": 3,
    "lower linearity limit": 5,
    "upper linearity limit": 250,
    "initial guess": [0.0, 0.0, 0.0, 0.0, 0.0],
    "dark threshold": 0.05,
    "flat field middle zone percentage": 0.2,
    "hot pixel threshold": 0.02,
    "acquired images path": "/tmp/camlin/acq",
    "flat fields path": "/tmp/camlin/flat",
    "dark frames path": "/tmp/camlin/dark",
    "original flat fields path": "/tmp/camlin/oflat",
    "original dark frames path": "/tmp/camlin/odark",
    "calibrated ICRFs": "ICRF_calibrated.txt",
    "channel names": ["Blue", "Green", "Red"],
    "STD data": "STD_data.txt",
    "camera mean data": ["mean_b.txt", "mean_g.txt", "mean_r.txt"],
    "camera base data": ["base_b.txt", "base_g.txt", "base_r.txt"],
    "source DoRF data": "dorfCurves.txt",
    "ICRFs": ["ICRF_b.txt", "ICRF_g.txt", "ICRF_r.txt"],
    "mean ICRFs": ["mean_ICRF_b.txt", "mean_ICRF_g.txt", "mean_ICRF_r.txt"],
    "principal components": ["PC_b.txt", "PC_g.txt", "PC_r.txt"],
}

_REF_MODULE_NAMES = (
    "read_config", "global_settings", "array_wrapper", "general_functions", "measurand",
    "cupy_measurand", "measurand_factory", "image_set", "exposure_series",
    "video_processing", "ICRF_calibration_exposure",
)


def available() -> bool:
    return (REFERENCE_MODULES / "measurand.py").is_file()


def load(config: dict | None = None) -> types.SimpleNamespace:
    """Import the reference modules and return them in a namespace.

    Re-importing with a different ``config`` drops the cached modules first because
    ``GlobalSettings`` freezes the values as class attributes at import time.
    """
    if not available():
        raise RuntimeError("reference tree /root/reference is not mounted on this machine")
    cfg = dict(DEFAULT_CONFIG)
    if config:
        cfg.update(config)

    for name in _REF_MODULE_NAMES:
        sys.modules.pop(name, None)

    stub = types.ModuleType("read_config")
    stub.current_directory = REFERENCE_MODULES
    stub.root_directory = REFERENCE_ROOT
    stub.data_directory = Path("/tmp/camlin/data")
    stub.read_config_single = lambda key: cfg.get(key, "")
    stub.read_config_list = lambda key: list(cfg.get(key, []))
    sys.modules["read_config"] = stub

    if str(REFERENCE_MODULES) not in sys.path:
        sys.path.insert(0, str(REFERENCE_MODULES))

    ns = types.SimpleNamespace()
    for name in _REF_MODULE_NAMES[1:]:
        if name == "cupy_measurand":
            continue
        setattr(ns, name, importlib.import_module(name))
    ns.gs = ns.global_settings.GlobalSettings
    return ns

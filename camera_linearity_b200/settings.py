"""Explicit settings object with the reference's constant NAMES (modules/global_settings.py:9-81).

The reference parses ``data/config.ini`` at import time, once per key; here the same attribute
names live on a plain class with defaults, can be overridden with ``configure(**kwargs)`` or
loaded from an ini file with the reference's section/key grammar (``read_config.py:12-67``).
The reference's own values for thresholds and kernel size are not shipped (``data/`` is
git-ignored), so the defaults below are ours and are listed in DESIGN.md.
"""
from __future__ import annotations

import configparser
from pathlib import Path

import torch


class GlobalSettings:
    DATA_PATH = Path("data")
    OUTPUT_PATH = Path("output")

    IM_SIZE_X = 2048
    IM_SIZE_Y = 1536

    DEFAULT_IMG_SRC_PATH = Path("data/acquired")
    DEFAULT_FLAT_PATH = Path("data/flat")
    DEFAULT_DARK_PATH = Path("data/dark")
    UNCALIBRATED_FLAT_PATH = Path("data/flat_raw")
    UNCALIBRATED_DARK_PATH = Path("data/dark_raw")
    ICRF_CALIBRATED_FILE = "ICRF_calibrated.txt"

    NUM_OF_CHS = 3
    CH_NAMES = ["Blue", "Green", "Red"]
    CH_CHARS = ["B", "G", "R"]
    CH_STR = {0: "Blue", 1: "Green", 2: "Red"}

    BIT_DEPTH = 8
    BITS = 256
    MAX_DN = 255
    MIN_DN = 0

    DATAPOINTS = 256
    DATAPOINT_MULTIPLIER = 1
    STD_FILE_NAME = "STD_data.txt"
    MEAN_DATA_FILES = []
    BASE_DATA_FILES = []
    DORF_FILE = "dorfCurves.txt"
    DORF_DATAPOINTS = 1024
    ICRF_FILES = []
    MEAN_ICRF_FILES = []
    NUM_OF_PCA_PARAMS = 5
    PCA_FILES = []
    IN_PCA_GUESS = [0.0] * 5

    DARK_THRESHOLD = 0.05
    FF_MID_PERCENTAGE = 0.2
    HOT_PIXEL_THRESHOLD = 0.02
    MEDIAN_FILTER_KERNEL_SIZE = 3

    LOWER_LIN_LIM = 5
    UPPER_LIN_LIM = 250

    # device the Measurands live on; there is exactly one backend (torch CUDA tensors)
    DEVICE = None

    @classmethod
    def device(cls) -> torch.device:
        if cls.DEVICE is not None:
            return torch.device(cls.DEVICE)
        return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")

    @classmethod
    def configure(cls, **kwargs) -> None:
        for key, value in kwargs.items():
            if not hasattr(cls, key):
                raise AttributeError(f"unknown setting {key}")
            setattr(cls, key, value)
        if "BIT_DEPTH" in kwargs:
            cls.BITS = 2 ** cls.BIT_DEPTH
            cls.MAX_DN = cls.BITS - 1

    _INI_KEYS = {
        "image size x": "IM_SIZE_X", "image size y": "IM_SIZE_Y", "channels": "NUM_OF_CHS",
        "bit depth": "BIT_DEPTH", "final datapoints": "DATAPOINTS",
        "datapoint multiplier": "DATAPOINT_MULTIPLIER", "original DoRF datapoints": "DORF_DATAPOINTS",
        "number of principal components": "NUM_OF_PCA_PARAMS",
        "median filter kernel size": "MEDIAN_FILTER_KERNEL_SIZE",
        "lower linearity limit": "LOWER_LIN_LIM", "upper linearity limit": "UPPER_LIN_LIM",
        "initial guess": "IN_PCA_GUESS", "dark threshold": "DARK_THRESHOLD",
        "flat field middle zone percentage": "FF_MID_PERCENTAGE",
        "hot pixel threshold": "HOT_PIXEL_THRESHOLD", "acquired images path": "DEFAULT_IMG_SRC_PATH",
        "flat fields path": "DEFAULT_FLAT_PATH", "dark frames path": "DEFAULT_DARK_PATH",
        "original flat fields path": "UNCALIBRATED_FLAT_PATH",
        "original dark frames path": "UNCALIBRATED_DARK_PATH", "calibrated ICRFs": "ICRF_CALIBRATED_FILE",
        "channel names": "CH_NAMES", "STD data": "STD_FILE_NAME", "camera mean data": "MEAN_DATA_FILES",
        "camera base data": "BASE_DATA_FILES", "source DoRF data": "DORF_FILE", "ICRFs": "ICRF_FILES",
        "mean ICRFs": "MEAN_ICRF_FILES", "principal components": "PCA_FILES",
    }
    _LIST_KEYS = {"IN_PCA_GUESS", "CH_NAMES", "MEAN_DATA_FILES", "BASE_DATA_FILES", "ICRF_FILES",
                  "MEAN_ICRF_FILES", "PCA_FILES"}
    _PATH_KEYS = {"DEFAULT_IMG_SRC_PATH", "DEFAULT_FLAT_PATH", "DEFAULT_DARK_PATH",
                  "UNCALIBRATED_FLAT_PATH", "UNCALIBRATED_DARK_PATH"}

    @classmethod
    def from_ini(cls, path) -> None:
        """Load a reference-style config.ini: values in section 'Float data' / 'Integer data' are
        cast, every other section stays string; list values are comma separated."""
        parser = configparser.ConfigParser()
        parser.optionxform = str
        parser.read(path)
        lookup = {k.lower(): v for k, v in cls._INI_KEYS.items()}
        updates = {}
        for section in parser.sections():
            cast = float if section == "Float data" else int if section == "Integer data" else str
            for key, raw in parser[section].items():
                name = lookup.get(key.lower())
                if name is None:
                    continue
                if name in cls._LIST_KEYS:
                    updates[name] = [cast(x) for x in raw.split(",")]
                elif name in cls._PATH_KEYS:
                    updates[name] = Path(raw)
                else:
                    updates[name] = cast(raw)
        cls.configure(**updates)
        cls.DATA_PATH = Path(path).resolve().parent
        cls.CH_CHARS = [n[0] for n in cls.CH_NAMES]

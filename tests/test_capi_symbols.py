"""The C-ABI library loads on a CPU-only machine and exports every symbol the header declares
(no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "camera_linearity.h"


@pytest.fixture(scope="module")
def lib():
    from camera_linearity_b200 import _lib, build
    build.build()
    return _lib.load(build_if_missing=False)


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(cl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for required in ("cl_linearize_dn", "cl_linearize_f64", "cl_hdr_merge", "cl_flat_roi_means",
                     "cl_welford_update", "cl_welford_finalize", "cl_welford_stack", "cl_icrf_curves",
                     "cl_icrf_energy_partial", "cl_icrf_energy_finalize"):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    from camera_linearity_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_host_only_entry_points(lib):
    assert lib.cl_abi_version() == 1
    assert lib.cl_status_string(0) == b"ok"
    assert b"unsupported" in lib.cl_status_string(-2)
    assert lib.cl_welford_stack_workspace_bytes(600, 1000) >= 4000
    assert lib.cl_flat_roi_means_workspace_bytes(10, 10, 3) > 0
    assert lib.cl_launch_count() == 0


def test_struct_layout_matches_header(lib):
    from camera_linearity_b200._lib import HdrMergeArgs, IcrfProblem
    # 6 int32, 3 pointers, 3 pointers, 2 pointers, double, 2 int32, 3 pointers, 2 pointers, 2 int32
    assert ctypes.sizeof(HdrMergeArgs) == 24 + 8 * 8 + 8 + 8 + 5 * 8 + 8
    assert ctypes.sizeof(IcrfProblem) == 32


def test_argument_validation_without_a_gpu(lib):
    from camera_linearity_b200._lib import HdrMergeArgs
    args = HdrMergeArgs()
    assert lib.cl_hdr_merge(ctypes.byref(args), None, 0, None) == -1        # n_exposures = 0
    assert lib.cl_linearize_dn(None, 1, None, None, None, None, None, 10, 3, 256, None) == -1
    assert lib.cl_linearize_dn(None, 1, None, None, None, None, None, 0, 3, 256, None) == 0   # empty input is a no-op


def test_ops_refuse_cpu_tensors():
    import torch
    from camera_linearity_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gaussian_weight(torch.zeros(4, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.welford_stack(torch.zeros((2, 4, 4, 3), dtype=torch.uint8))

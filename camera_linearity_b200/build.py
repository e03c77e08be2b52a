"""Build the C-ABI CUDA library ``libcamlin_b200.so`` in-tree with nvcc for sm_100a.

    python -m camera_linearity_b200.build [--force]

The library has no PyTorch dependency (static cudart); it is loaded with ctypes by
``camera_linearity_b200._lib``.  Objects are compiled in parallel, one per ``.cu`` file.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"
LIB = PKG / "libcamlin_b200.so"
OBJ_DIR = PKG / "build"
STAMP = OBJ_DIR / "sources.sha256"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    f"-I{INCLUDE}", f"-I{CSRC}",
]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(path).exists():
        raise RuntimeError("nvcc not found: cannot build libcamlin_b200.so")
    return path


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    tmp = OBJ_DIR / (LIB.name + ".tmp")             # link next to the objects, then move into place atomically
    r = subprocess.run([nvcc, "-shared", "-o", str(tmp), *map(str, objs),
                        "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    STAMP.write_text(_fingerprint())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""In-tree build of libmockingbird_b200.so (nvcc, sm_100a only).

The shared library is the product's compute path; there is no CPU fallback.  The built .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libmockingbird_b200.so"
STAMP = PKG_DIR / ".libmockingbird_b200.hash"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh"))
                    + [PKG_DIR.parent / "include" / "mockingbird_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB_PATH.is_file() and STAMP.is_file() and STAMP.read_text().strip() == _digest()


def find_nvcc() -> str | None:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.isfile(cand) else None


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source into one shared library.  Raises on failure."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libmockingbird_b200.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *[str(s) for s in sources()]]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    STAMP.write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

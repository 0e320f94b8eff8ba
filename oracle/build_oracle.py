"""Build the oracle's C restatement (TEST INFRASTRUCTURE): oracle/libwavernn_twin.so.

gcc -O2 -ffp-contract=off so that every a*b+c stays two rounded operations unless written as
fmaf() - the arithmetic contract shared with the CUDA kernel (include/mb_wavernn_math.h).
The reference is pure Python: there is no C/C++ reference to compile into oracle/_ref/.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "wavernn_twin.c"
LIB = HERE / "libwavernn_twin.so"


def build(force: bool = False) -> Path:
    hdr = HERE.parent / "include" / "mb_wavernn_math.h"
    if not force and LIB.is_file() and LIB.stat().st_mtime >= max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-fopenmp", "-fPIC", "-shared",
           "-o", str(LIB), str(SRC), "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

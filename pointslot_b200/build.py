"""Build libpslot_orb.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libpslot_orb.so"
SOURCES = ["ingest.cu", "pyramid.cu", "fast.cu", "quadtree.cu", "describe.cu", "stereo.cu", "grid.cu", "partition.cu", "cvorb.cu", "bow.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # parity rule 3: no fused multiply-add anywhere on the float islands
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared", "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libpslot_orb.so cannot be built")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "pslot_internal.h", CSRC / "tma.cuh", CSRC / "orb_common.cuh", PKG.parent / "include" / "pslot_orb.h",
                                          PKG.parent / "include" / "pslot_orb_pattern.inc"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB), *SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))

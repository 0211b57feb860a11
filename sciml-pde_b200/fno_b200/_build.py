"""Builds libfno_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each translation unit is compiled to an object under ``sciml-pde_b200/build/`` (in parallel,
only when its sources changed) and the objects are linked into ``fno_b200/libfno_sm100.so``."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR.parent / "csrc"
OBJ_DIR = PKG_DIR.parent / "build"
LIB_PATH = PKG_DIR / "libfno_sm100.so"
SOURCES = ["api.cu", "transform2d.cu", "mix.cu", "pointwise.cu", "axis3d.cu", "headlift.cu", "steptail.cu", "head_tc.cu", "transform2d_tc.cu", "head_bwd_tc.cu", "dataset.cu", "layer2d_tc.cu", "metrics.cu", "spectral1d.cu", "pointwise_tc.cu", "head_wide_tc.cu"]
HEADERS = [CSRC / "common.cuh", CSRC / "tc_common.cuh", PKG_DIR.parent.parent / "include" / "fno_sm100.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    mtime = target.stat().st_mtime
    return any(d.stat().st_mtime > mtime for d in deps)


def needs_build() -> bool:
    return _stale(LIB_PATH, [CSRC / s for s in SOURCES] + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJ_DIR.mkdir(exist_ok=True)

    def compile_one(src: str):
        obj = OBJ_DIR / (src[:-3] + ".o")
        if not force and not _stale(obj, [CSRC / src] + HEADERS):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        print("[fno_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH), *[str(o) for o in objs]]
    print("[fno_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)

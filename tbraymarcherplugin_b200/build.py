"""Build libtbrm.so (hand-written sm_100a CUDA + the C ABI of include/tbrm.h) in-tree with nvcc.

    python -m tbraymarcherplugin_b200.build [--force] [--verbose]

--fmad=false and -ffp-contract=off are part of the arithmetic contract: fused multiply-adds exist
only where the sources spell __fmaf_rn / fmaf, which is what makes GPU-vs-oracle parity bit-exact.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libtbrm.so"
SOURCES = ["api.cu", "sweep.cu", "raymarch.cu", "mandelbulb.cu", "synth.cu", "ingest.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC)")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG_DIR.parent / "include" / "tbrm.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    extra = os.environ.get("TBRM_EXTRA_NVCC_FLAGS", "").split()  # diagnostic builds (e.g. -DTBRM_CHAIN_TIMERS)
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src} ---\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lz"]
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

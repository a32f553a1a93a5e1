"""In-tree build of the CUDA engine for sm_100a (explicit nvcc, no JIT cache).

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    sequencealigning_b200/_lib/libsa_engine.so   the C-ABI shared library (include/sa_engine.h)
    sequencealigning_b200/_lib/int_peak          integer-pipe microbenchmark (roofline denominator)
    sequencealigning_b200/_lib/sa_align          CLI with the reference's flags (parse.rs:8-34)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsa_engine.so")
INT_PEAK_PATH = os.path.join(LIB_DIR, "int_peak")
CLI_PATH = os.path.join(LIB_DIR, "sa_align")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")


def _newer(target: str, sources: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs: str) -> List[str]:
    out = []
    for d in dirs:
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    return out


def build_all(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    deps = _sources(CSRC, os.path.join(_ROOT, "include"))
    extra = ["-Xptxas", "-v"] if verbose else []
    if force or _newer(LIB_PATH, deps):
        # one object per translation unit, compiled concurrently, then one link
        srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
        obj_dir = os.path.join(LIB_DIR, "obj")
        os.makedirs(obj_dir, exist_ok=True)
        headers = [d for d in deps if not d.endswith(".cu")]
        jobs = []
        objs = []
        for src in srcs:
            obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if force or _newer(obj, [src] + headers):
                cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", obj, src, "-I", os.path.join(_ROOT, "include")]
                jobs.append((src, subprocess.Popen(cmd)))
        for src, proc in jobs:
            if proc.wait() != 0:
                raise RuntimeError(f"nvcc failed on {src}")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs], check=True)
    mb = os.path.join(CSRC, "microbench", "int_peak.cu")
    if force or _newer(INT_PEAK_PATH, [mb]):
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", INT_PEAK_PATH, mb], check=True)
    cli = os.path.join(CSRC, "cli", "sa_align.cpp")
    if os.path.exists(cli) and (force or _newer(CLI_PATH, deps + [LIB_PATH])):
        subprocess.run(
            ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", CLI_PATH, cli, "-I", os.path.join(_ROOT, "include"),
             "-L", LIB_DIR, "-lsa_engine", "-Wl,-rpath,$ORIGIN"],
            check=True,
        )
    return LIB_PATH


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_PATH)

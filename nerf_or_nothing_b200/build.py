"""Builds nerf_or_nothing_b200/libnerfb200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m nerf_or_nothing_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_obj"
LIB = HERE / "libnerfb200.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                     "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("NERF_NVCC_DEFS", "").split()  # A/B builds: -DNAME=0


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (Path(c).exists() or c == "nvcc"):
            return c
    return "nvcc"


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "nerfb200.h"]
    srcs = sorted(CSRC.glob("*.cu"))
    logs = {}

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and not _stale(obj, [src] + headers):
            return obj, None
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    for (o, log), src in zip(results, srcs):
        if log is not None:
            logs[src.name] = log
            (OBJ / (src.stem + ".ptxas.log")).write_text(log)
    if force or _stale(LIB, objs):
        # --no-undefined: a declaration / definition mismatch between two .cu files must fail the build, not the first call
        cmd = [_nvcc()] + ARCH + ["-shared", "-Xlinker", "--no-undefined", "-o", str(LIB)] + [str(o) for o in objs] + ["-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for k, v in logs.items():
            print(f"==== {k}\n{v}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))

"""Builds libshared_tree_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree.

    python genome-compression_b200/_build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to genome-compression_b200/build/
(git-ignored), the library next to this file (git-ignored, but it travels to the GPU
box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "build"
LIB = HERE / "libshared_tree_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]
HOST_BINARIES = {"compress_b200": ["compress.cpp"]}


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _deps():
    return list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "shared_tree_b200.h", HERE.parent / "include" / "shared_tree_b200_dist.h"]


def _stale(target: Path, inputs) -> bool:
    if not target.exists():
        return True
    mt = target.stat().st_mtime
    return any(Path(i).stat().st_mtime > mt for i in inputs)


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    deps = _deps()
    jobs = []
    for src in _sources():
        obj = BUILD / (src.stem + ".o")
        if force or _stale(obj, [src] + deps):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *ARCH, *FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        (BUILD / (src.stem + ".ptxas.log")).write_text(res.stderr)
        if verbose:
            print(res.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [BUILD / (s.stem + ".o") for s in _sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    build_host(force)
    return LIB


def build_host(force: bool = False) -> None:
    """C++ host side (the reference-shaped shim + drivers) on top of the C ABI."""
    host = HERE / "host"
    inc = HERE.parent / "include"
    for name, srcs in HOST_BINARIES.items():
        paths = [host / s for s in srcs]
        if not all(p.exists() for p in paths):
            continue
        out = host / name
        deps = paths + list((host / "include").glob("*.h")) + [LIB, inc / "shared_tree_b200.h"]
        if force or _stale(out, deps):
            cmd = ["g++", "-std=c++17", "-O2", "-Wall", f"-I{inc}", f"-I{host / 'include'}", *map(str, paths), "-o", str(out),
                   f"-L{HERE}", "-lshared_tree_b200", "-Wl,-rpath,$ORIGIN/..", "-pthread"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"g++ failed for {name}:\n{res.stdout}\n{res.stderr}")


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)

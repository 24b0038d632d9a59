"""Builds libpfe_b200.so in-tree with nvcc for sm_100a (the only target).

`python -m paintfe_b200.build [--force] [--verbose]`
The library has no torch or Python dependency; it links cudart statically.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpfe_b200.so")
SOURCES = ["ctx.cu", "flatten.cu", "gaussian.cu", "filters.cu", "adjust.cu", "warp.cu", "brush.cu", "hostops.cu", "effects2.cu", "effects3.cu", "geometry.cu", "tiles.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "fx_common.cuh"), os.path.join(CSRC, "hsl.cuh"), os.path.join(CSRC, "blend.cuh"), os.path.join(os.path.dirname(HERE), "include", "pfe_b200.h")]

# -fmad=false: the reference (Rust) never contracts a*b+c; bit-exactness depends on it.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Werror", "cross-execution-space-call",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    env = dict(os.environ)
    # nvcc's host compiler: the plain system gcc (a wrapper earlier on PATH lacks libgomp specs)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def one(src: str) -> str:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS + [os.path.abspath(__file__)]):
            cmd = [cc] + ccbin + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [cc] + ccbin + ["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
